"""Summarise ncu outputs into small text files for profiles/ (run where ncu is installed; no GPU needed).

    python tools/ncu_summary.py launches gpurun_out/launches.csv profiles/r1_launches.txt
    python tools/ncu_summary.py full gpurun_out/prof.ncu-rep profiles/r1_full.txt
"""
import collections
import csv
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum",
    "smsp__inst_executed.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum", "launch__waves_per_multiprocessor",
]


def launches(src, dst):
    rows = [r for r in csv.reader(open(src)) if len(r) > 5]
    start = next(i for i, r in enumerate(rows) if r[0] == "ID")
    hdr = rows[start]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    data = [(r[ki], float(r[vi].replace(",", ""))) for r in rows[start + 1:]]
    marks = [i for i, (k, _) in enumerate(data) if "adam_kernel" in k]
    step = data[marks[-2] + 1:marks[-1] + 1] if len(marks) >= 2 else data
    tot = sum(v for _, v in step)
    agg = collections.OrderedDict()
    for k, v in step:
        e = agg.setdefault(k[:100], [0, 0.0])
        e[0] += 1
        e[1] += v
    with open(dst, "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none; one train step (last of the run)\n")
        f.write(f"# {len(step)} launches, {tot / 1000:.1f} us total (cold-cache, serialised: compare SHARES)\n")
        for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{v / 1000:9.1f} us  x{c:<3d} {100 * v / tot:5.1f}%  {k}\n")
    print(f"wrote {dst}")


def full(src, dst):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(dst, "w") as f:
        f.write(f"# ncu --set full --clock-control none, source {src}\n")
        seen = set()
        for r in rows[2:]:
            if r[idx['Kernel Name']] in seen:  # one section per kernel (the first captured launch)
                continue
            seen.add(r[idx['Kernel Name']])
            f.write(f"== {r[idx['Kernel Name']]}\n")
            for w in WANT:
                if w in idx:
                    f.write(f"   {w:72s} {r[idx[w]]} {units[idx[w]]}\n")
    print(f"wrote {dst}")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
