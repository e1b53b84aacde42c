import sys
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import zignal_b200 as zb
rng = np.random.default_rng(0)
shape = (int(sys.argv[1]), int(sys.argv[2]))
img = rng.integers(0, 256, shape + (4,), dtype=np.uint8)
dev = zb.Image.from_numpy(img)
a = np.float32(0.3)
out = dev.rotate(a)
torch.cuda.synchronize()
print("ok", zb.lib().zb_last_kernel().decode(), out.rows, out.cols)
