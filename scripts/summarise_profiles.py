"""Turn ncu outputs brought back in gpurun_out/ into the small text summaries kept under profiles/.

    python scripts/summarise_profiles.py <round-tag> <launch-list.csv> [<report.ncu-rep> ...]
"""
import collections
import csv
import re
import subprocess
import sys

tag, launches, reports = sys.argv[1], sys.argv[2], sys.argv[3:]
rows = [r for r in csv.reader(open(launches)) if len(r) > 5]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
seq = []
for r in rows[1:]:
    name = re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("bl::", "").replace("<unnamed>::", "")
    seq.append((name, float(r[vi].replace(",", "")) / 1e3))
agg = collections.OrderedDict()
for n, t in seq:
    a = agg.setdefault(n, [0, 0.0])
    a[0] += 1
    a[1] += t
tot = sum(a[1] for a in agg.values())
out = [f"# ncu launch list summary ({tag})", "",
       "Command: `ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv python bench.py --quick --steps 1 --warmup 0`",
       "(one bench step = one forward + one adjoint sweep, n = 1M, depth 100, fp32; per-launch times are cold-cache and",
       "serialised by the profiler: compare SHARES, not absolutes).", "",
       f"launches: {len(seq)}, sum of kernel durations: {tot / 1e3:.2f} ms", "",
       "| kernel | launches | total us | avg us | share |", "|---|---:|---:|---:|---:|"]
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append(f"| `{k}` | {c} | {t:.1f} | {t / c:.2f} | {100 * t / tot:.1f}% |")
with open(f"profiles/{tag}_launches_summary.md", "w") as f:
    f.write("\n".join(out) + "\n")
with open(f"profiles/{tag}_launches.csv", "w") as f:
    f.write("index,kernel,duration_us\n")
    for i, (n, t) in enumerate(seq):
        f.write(f'{i},"{n}",{t:.3f}\n')

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit_shared_mem", "sm__cycles_elapsed.max",
        "smsp__cycles_active.avg"]
for rep in reports:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines()))
    h, units = rr[0], rr[1]
    lines = [f"# ncu --set full summary: {rep.split('/')[-1]} ({tag})", ""]
    for r in rr[2:]:
        lines.append(f"## {r[h.index('Kernel Name')]}")
        for w in WANT:
            if w in h:
                lines.append(f"- `{w}` = {r[h.index(w)]} {units[h.index(w)]}")
        lines.append("")
    name = rep.split("/")[-1].replace(".ncu-rep", "")
    with open(f"profiles/{tag}_{name}.md", "w") as f:
        f.write("\n".join(lines) + "\n")
print("wrote profiles/")
