"""Summarise an .ncu-rep (read here, no GPU needed) into a small markdown file for profiles/.
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01_play_kernel.md "title" [sims_in_launch]"""
import csv, io, subprocess, sys

rep, out, title = sys.argv[1], sys.argv[2], sys.argv[3]
sims = float(sys.argv[4]) if len(sys.argv) > 4 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
m = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__waves_per_multiprocessor", "sm__warps_active.avg.per_cycle_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__icc_request_hit_rate.pct",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
lines = [f"# {title}", "", f"source: `{rep}` (ncu --set full --clock-control none --import-source on)", "", "| metric | value |", "|---|---|"]
for w in want:
    if w in m:
        lines.append(f"| {w} | {m[w][0]} {m[w][1]} |")
if sims and "smsp__inst_executed.sum" in m:
    lines.append(f"| warp instructions per simulation | {float(m['smsp__inst_executed.sum'][0]) / sims:.0f} |")
    tr = float(m["dram__bytes_read.sum"][0]) + float(m["dram__bytes_write.sum"][0])
    unit = m["dram__bytes_read.sum"][1]
    lines.append(f"| DRAM traffic per simulation | {tr / sims * {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1}.get(unit, 1):.0f} B |")
lines += ["", "## warp stall reasons (cycles per issued instruction)", "", "| reason | value |", "|---|---|"]
st = [(h, float(v)) for h, (v, u) in m.items() if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and "not_issued" not in h]
for h, v in sorted(st, key=lambda x: -x[1])[:10]:
    lines.append(f"| {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')} | {v:.2f} |")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur, h2, per = None, None, []
for r in csv.reader(io.StringIO(src)):
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] == "Line No": h2 = r; continue
    if h2 is None or r[0] == "Function Name": continue
    try: ln = int(r[0])
    except ValueError: continue
    if r[2] != "-": continue
    try: per.append((int(r[h2.index("Instructions Executed")]), int(r[h2.index("# Samples")]), cur, ln, r[1].strip()[:90]))
    except ValueError: pass
tot, ts = sum(p[0] for p in per) or 1, sum(p[1] for p in per) or 1
lines += ["", "## hottest source lines (share of executed instructions / of stall samples)", "", "| inst % | samples % | line |", "|---|---|---|"]
for n, s, f, ln, text in sorted(per, reverse=True)[:25]:
    lines.append(f"| {n / tot * 100:.1f} | {s / ts * 100:.1f} | `{f}:{ln}` {text.replace('|', '/')} |")
open(out, "w").write("\n".join(lines) + "\n")
print("wrote", out)
