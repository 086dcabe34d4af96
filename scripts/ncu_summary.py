"""ncu report -> small JSON summary for profiles/ (run here, no GPU needed).
usage: python scripts/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01_xxx_summary.json [algorithmic_bytes]"""
import csv, json, subprocess, sys

KEEP = {
    "gpu__time_duration.sum": "duration", "dram__bytes_read.sum": "dram_read", "dram__bytes_write.sum": "dram_write",
    "launch__registers_per_thread": "registers", "launch__grid_size": "grid", "launch__block_size": "block",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active": "pipe_xu_pct",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active": "pipe_fma_pct",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active": "pipe_alu_pct",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active": "pipe_lsu_pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "pipe_tensor_pct",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum": "smem_wavefronts",
    "smsp__inst_executed.sum": "warp_instructions", "sm__cycles_elapsed.max": "sm_cycles",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio": "stall_short_scoreboard",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio": "stall_long_scoreboard",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio": "stall_mio_throttle",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio": "stall_math_pipe_throttle",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio": "stall_wait",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio": "stall_barrier",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio": "stall_not_selected",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio": "stall_lg_throttle",
}


def main():
    rep, out = sys.argv[1], sys.argv[2]
    algo = float(sys.argv[3]) if len(sys.argv) > 3 else None
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    kernels = []
    for vals in rows[2:]:
        d = {"kernel": vals[hdr.index("Kernel Name")]}
        for h, u, v in zip(hdr, units, vals):
            if h in KEEP:
                try:
                    d[KEEP[h]] = float(v.replace(",", ""))
                    d[KEEP[h] + "_unit"] = u
                except ValueError:
                    pass
        if "dram_read" in d and "dram_write" in d:
            scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
            d["dram_bytes_per_launch"] = d["dram_read"] * scale.get(d.get("dram_read_unit"), 1) + \
                d["dram_write"] * scale.get(d.get("dram_write_unit"), 1)
            if algo:
                d["algorithmic_bytes"] = algo
                d["traffic_over_algorithmic"] = d["dram_bytes_per_launch"] / algo
        kernels.append(d)
    top = dict(kernels[0]) if kernels else {}
    top["source"] = rep
    top["all_launches"] = kernels
    json.dump(top, open(out, "w"), indent=1)
    print(json.dumps({k: v for k, v in top.items() if k != "all_launches"}, indent=1))


if __name__ == "__main__":
    main()
