"""Summarise one kernel of an `ncu --set full` report into profiles/<name>.json (+ .txt):
duration, issue-slot / pipe utilisation, occupancy, DRAM bytes, instruction count, top stalls.

  python tools/ncu_summary.py gpurun_out/x.ncu-rep profiles/r2_pair_kernel_ncu "command that was profiled"
"""
import csv
import io
import json
import subprocess
import sys

KEYS = {
    "kernel_ms": ("gpu__time_duration.sum", "ms"),
    "dram_read_bytes": ("dram__bytes_read.sum", "bytes"),
    "dram_write_bytes": ("dram__bytes_write.sum", "bytes"),
    "dram_throughput_pct": ("dram__throughput.avg.pct_of_peak_sustained_elapsed", None),
    "warp_instructions": ("smsp__inst_executed.sum", None),
    "issue_slots_busy_pct": ("smsp__issue_active.avg.pct_of_peak_sustained_active", None),
    "ipc_sum_over_sms": ("sm__inst_executed.sum.per_cycle_active", None),
    "eligible_warps_per_scheduler": ("smsp__warps_eligible.avg.per_cycle_active", None),
    "achieved_occupancy_pct": ("sm__warps_active.avg.pct_of_peak_sustained_active", None),
    "registers_per_thread": ("launch__registers_per_thread", None),
    "dynamic_smem_per_block_bytes": ("launch__shared_mem_per_block_dynamic", "bytes"),
    "pipe_alu_pct": ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", None),
    "pipe_fma_pct": ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", None),
    "pipe_fp64_pct": ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", None),
    "pipe_xu_pct": ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", None),
    "pipe_lsu_pct": ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", None),
    "l1_hit_rate_pct": ("l1tex__t_sector_hit_rate.pct", None),
    "l2_hit_rate_pct": ("lts__t_sector_hit_rate.pct", None),
}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3,
        "second": 1e3, "msecond": 1.0, "usecond": 1e-3, "nsecond": 1e-6}


def main():
    rep, out, cmd = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    names, units, vals = rows[0], rows[1], rows[2]
    col = {n: i for i, n in enumerate(names)}
    res = {"report": rep.split("/")[-1], "command": cmd, "kernel": vals[col["Kernel Name"]]}
    for key, (metric, kind) in KEYS.items():
        if metric not in col:
            continue
        v = vals[col[metric]].replace(",", "")
        if not v:
            continue
        x = float(v)
        u = units[col[metric]]
        u = u.split("/")[0]
        if kind in ("bytes", "ms") and u in UNIT:
            x *= UNIT[u]
        res[key] = x
    if "dram_read_bytes" in res and "dram_write_bytes" in res:
        res["dram_bytes"] = res["dram_read_bytes"] + res["dram_write_bytes"]
    stalls = {}
    for n, i in col.items():
        if n.startswith("smsp__average_warps_issue_stalled_") and n.endswith("_per_issue_active.ratio") and vals[i]:
            stalls[n[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]] = float(vals[i])
    res["stall_warps_per_issue"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:6])
    json.dump(res, open(out + ".json", "w"), indent=1)
    with open(out + ".txt", "w") as f:
        f.write(f"# {res['kernel']}\n# {cmd}\n")
        for k, v in res.items():
            if k not in ("kernel", "command"):
                f.write(f"{k:32s} {v}\n")
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
