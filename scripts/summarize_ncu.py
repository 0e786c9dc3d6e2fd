"""Summarise an .ncu-rep (read on the CPU box) into a small markdown file for profiles/.

    python scripts/summarize_ncu.py gpurun_out/prof.ncu-rep profiles/r01_xxx.md "title"
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_active.avg.per_cycle_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum",
    "sm__cycles_active.avg",
]


def run(args):
    return subprocess.run(["ncu", "-i", *args], capture_output=True, text=True).stdout


def main(rep, out, title):
    raw = list(csv.reader(io.StringIO(run([rep, "--page", "raw", "--csv"]))))
    hdr, units, vals = raw[0], raw[1], raw[2]
    name = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
    lines = [f"# {title}", "", f"kernel: `{name}`", "", f"source: `{rep}` (ncu --set full --clock-control none)", "",
             "| metric | value | unit |", "|---|---|---|"]
    for h, u, v in zip(hdr, units, vals):
        if any(h == k or h.endswith("." + k) for k in KEYS) or h in KEYS:
            lines.append(f"| {h} | {v} | {u} |")
    stall = [(h, float(v)) for h, v in zip(hdr, vals)
             if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")]
    lines += ["", "## warp stall reasons (warps stalled per issue-active cycle)", "", "| reason | ratio |", "|---|---|"]
    for h, v in sorted(stall, key=lambda x: -x[1])[:9]:
        lines.append(f"| {h.split('issue_stalled_')[1].replace('_per_issue_active.ratio', '')} | {v:.3f} |")
    src = list(csv.reader(io.StringIO(run([rep, "--page", "source", "--csv"]))))
    if len(src) > 3:
        h2 = src[1]
        ix = {h: i for i, h in enumerate(h2)}
        data = src[2:]
        tot = sum(int(r[ix["# Samples"]]) for r in data) or 1
        ops = collections.Counter()
        for r in data:
            try:
                ops[r[ix["Source"]].split()[0 if not r[ix["Source"]].strip().startswith("@") else 1]] += float(
                    r[ix["Instructions Executed"]])
            except (ValueError, IndexError):
                pass
        lines += ["", "## executed warp instructions by opcode (top 12)", "", "| opcode | warp instructions |", "|---|---|"]
        for op, n in ops.most_common(12):
            lines.append(f"| {op} | {n:.3g} |")
        lines += ["", "## hottest instructions by stall samples", "", "| # | SASS | samples % | executed |", "|---|---|---|---|"]
        for i in sorted(range(len(data)), key=lambda i: -int(data[i][ix["# Samples"]]))[:12]:
            r = data[i]
            lines.append(f"| {i} | `{r[ix['Source']].strip()[:70]}` | {100 * int(r[ix['# Samples']]) / tot:.1f} | "
                         f"{float(r[ix['Instructions Executed']] or 0):.3g} |")
    open(out, "w").write("\n".join(lines) + "\n")
    print("wrote", out)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else sys.argv[1])
