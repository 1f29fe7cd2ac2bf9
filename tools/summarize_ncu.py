"""Condenses `ncu --page raw --csv` exports (tools/ncu_capture.sh) into one JSON per capture + a markdown table.
    python tools/summarize_ncu.py <dir-with-ncu_*_raw.csv> [--dominant <tag>]"""
import csv
import json
import os
import sys

KEYS = {
    "duration_us": "gpu__time_duration.sum",
    "dram_read_bytes": "dram__bytes_read.sum",
    "dram_write_bytes": "dram__bytes_write.sum",
    "dram_pct_of_peak": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm_pct_of_peak": "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex_pct_of_peak": "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l2_pct_of_peak": "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "issue_active_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
    "registers_per_thread": "launch__registers_per_thread",
    "grid": "launch__grid_size",
    "block": "launch__block_size",
    "dyn_smem_bytes": "launch__shared_mem_per_block_dynamic",
    "smem_config_bytes": "launch__shared_mem_config_size",
    "l1_hit_pct": "l1tex__t_sector_hit_rate.pct",
    "l2_hit_pct": "lts__t_sector_hit_rate.pct",
    "warp_inst": "smsp__inst_executed.sum",
    "global_ld_requests": "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "global_ld_sectors": "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "shared_wavefronts": "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "shared_bank_conflicts": "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "stall_long_scoreboard": "smsp__pcsamp_warps_issue_stalled_long_scoreboard",
    "stall_lg_throttle": "smsp__pcsamp_warps_issue_stalled_lg_throttle",
    "stall_barrier": "smsp__pcsamp_warps_issue_stalled_barrier",
    "stall_math_pipe": "smsp__pcsamp_warps_issue_stalled_math_pipe_throttle",
    "stall_wait": "smsp__pcsamp_warps_issue_stalled_wait",
    "stall_short_scoreboard": "smsp__pcsamp_warps_issue_stalled_short_scoreboard",
    "stall_mio_throttle": "smsp__pcsamp_warps_issue_stalled_mio_throttle",
}
UNIT_SCALE = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1, "ms": 1e3, "us": 1, "ns": 1e-3, "Kbyte/block": 1e3, "byte/block": 1}


def load(path):
    rows = list(csv.reader(open(path)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    out = {"kernel": vals[hdr.index("Kernel Name")]}
    for name, key in KEYS.items():
        if key in hdr:
            i = hdr.index(key)
            try:
                v = float(vals[i].replace(",", ""))
            except ValueError:
                continue
            out[name] = v * UNIT_SCALE.get(units[i], 1)
    return out


def main():
    d = sys.argv[1]
    dominant = sys.argv[sys.argv.index("--dominant") + 1] if "--dominant" in sys.argv else None
    lines = ["| capture | kernel | µs | DRAM rd+wr MB | L1TEX % | SM % | issue % | warps % | regs | L1 hit % | long_sb / lg_throttle / barrier |", "|---|---|---|---|---|---|---|---|---|---|---|"]
    for f in sorted(os.listdir(d)):
        if not (f.startswith("ncu_") and f.endswith("_raw.csv")):
            continue
        tag = f[4:-8]
        s = load(os.path.join(d, f))
        with open(os.path.join(d, "ncu_%s_summary.json" % tag), "w") as fh:
            json.dump(s, fh, indent=1)
        lines.append("| %s | `%s` | %.1f | %.1f | %.1f | %.1f | %.1f | %.1f | %d | %.1f | %d / %d / %d |" % (
            tag, s["kernel"][:60], s.get("duration_us", 0), (s.get("dram_read_bytes", 0) + s.get("dram_write_bytes", 0)) / 1e6,
            s.get("l1tex_pct_of_peak", 0), s.get("sm_pct_of_peak", 0), s.get("issue_active_pct", 0), s.get("warps_active_pct", 0),
            s.get("registers_per_thread", 0), s.get("l1_hit_pct", 0), s.get("stall_long_scoreboard", 0), s.get("stall_lg_throttle", 0), s.get("stall_barrier", 0)))
        if dominant == tag:
            with open(os.path.join(os.path.dirname(os.path.abspath(d)), "dominant_kernel.json"), "w") as fh:
                json.dump({"capture": tag, "kernel": s["kernel"], "dram_bytes_per_launch": s.get("dram_read_bytes", 0) + s.get("dram_write_bytes", 0),
                           "note": "ncu --set full capture of one launch over 16 x 1080p fp16 frames; bench.py scales it to its 64-frame launch", "frames_in_capture": 16, **s}, fh, indent=1)
    print("\n".join(lines))


if __name__ == "__main__":
    main()
