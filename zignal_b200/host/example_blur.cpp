// example_blur.cpp -- compiles the C++ host mirror against the C ABI (link check; runs only on a GPU box).
// Mirrors the reference's examples/src/blur_box_vs_gaussian.zig call pattern: boxBlur vs gaussianBlur on one image.
#include <cstdio>
#include <vector>

#include "zignal.hpp"

int main() {
    using namespace zignal;
    const uint32_t rows = 64, cols = 96;
    std::vector<Rgba8> in(rows * cols), box(rows * cols), gauss(rows * cols);
    for (uint32_t i = 0; i < rows * cols; ++i) in[i] = Rgba8{(uint8_t)(i * 7), (uint8_t)(i * 13), (uint8_t)(i * 29), 255};
    Image<Rgba8> src{rows, cols, in.data(), cols}, b{rows, cols, box.data(), cols}, g{rows, cols, gauss.data(), cols};
    try {
        src.boxBlur(b, 2);
        src.gaussianBlur(g, 1.4f);
    } catch (const Error& e) {
        std::fprintf(stderr, "%s (%s)\n", e.what(), zb_last_error());
        return 1;
    }
    std::printf("box[0]=%u gauss[0]=%u\n", box[0].r, gauss[0].r);
    return 0;
}
