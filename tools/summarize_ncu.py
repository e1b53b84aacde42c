"""Condense `ncu --page raw --csv` exports into the small JSON summaries kept under profiles/.
usage: python tools/summarize_ncu.py gpurun_out/<tag>_ncu_<cfg>.csv ... -o profiles/<tag>_ncu_summary.json"""
import csv
import json
import sys

KEYS = {
    "gpu__time_duration.sum": "time",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed": "l1tex_pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "smsp__inst_executed.sum": "warp_instructions",
    "launch__registers_per_thread": "registers",
    "launch__waves_per_multiprocessor": "waves_per_sm",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active": "pipe_alu_pct",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active": "pipe_fma_pct",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active": "pipe_lsu_pct",
    "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active": "pipe_tensor_pct",
    "sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active": "tensor_hmma_cycles_pct",
    "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active": "pipe_uniform_pct",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "smem_bank_conflicts",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum": "smem_wavefronts",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio": "stall_long_scoreboard",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio": "stall_short_scoreboard",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio": "stall_barrier",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio": "stall_wait",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio": "stall_math_throttle",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio": "stall_mio_throttle",
}


def main():
    args = sys.argv[1:]
    out = args[args.index("-o") + 1]
    files = [a for a in args if a.endswith(".csv")]
    res = {}
    for f in files:
        rows = list(csv.reader(open(f)))
        hdr, units = rows[0], rows[1]
        ix = {h: i for i, h in enumerate(hdr)}
        ks = []
        for r in rows[2:]:
            k = {"kernel": r[ix["Kernel Name"]], "grid": r[ix["Grid Size"]], "block": r[ix["Block Size"]]}
            for m, name in KEYS.items():
                if m in ix and r[ix[m]] not in ("", "n/a"):
                    try:
                        k[name] = float(r[ix[m]])
                    except ValueError:
                        k[name] = r[ix[m]]
                    k[name + "_unit"] = units[ix[m]]
            for h in hdr:   # anything that mentions the tensor pipe / tcgen05
                if ("tensor" in h or "tmem" in h or "utc" in h) and "pct" in h and "avg" in h and r[ix[h]] not in ("", "n/a", "0"):
                    k.setdefault("tensor_metrics", {})[h] = r[ix[h]]
            ks.append(k)
        res[f.split("/")[-1]] = ks
    json.dump(res, open(out, "w"), indent=1)
    print("wrote", out, {k: len(v) for k, v in res.items()})


if __name__ == "__main__":
    main()
