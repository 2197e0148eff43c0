"""Turn gpurun_out ncu outputs into the committed text summaries under profiles/."""
import csv, collections, re, subprocess, sys, os
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
out = open(f"profiles/{tag}_summary.md", "w")
def P(*a):
    print(*a); print(*a, file=out)
# 1. launch list
lines = [l for l in open(f"gpurun_out/{tag}_launches.csv") if not l.startswith("==")]
rows = list(csv.DictReader(lines))
agg = collections.OrderedDict()
for row in rows:
    name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "")[:64]
    v = float(row["Metric Value"].replace(",", "")); u = row["Metric Unit"]
    v = v / 1e3 if u in ("ns", "nsecond") else v * 1e3 if u in ("ms", "msecond") else v * 1e6 if u in ("s", "second") else v
    agg.setdefault(name, []).append(v)
tot = sum(sum(v) for v in agg.values())
P(f"# {tag}: ncu launch list of `NFA_BENCH_CLOCK_LOAD_STEPS=0 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-reference-cuda`")
P("(`ncu --metrics gpu__time_duration.sum --clock-control none`; serialised, cold-cache: compare SHARES)\n")
P("| kernel | launches | mean us | share |"); P("|---|---:|---:|---:|")
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    P(f"| `{k}` | {len(v)} | {sum(v)/len(v):.1f} | {100*sum(v)/tot:.1f}% |")
P("\nNotes: the launch list covers the whole bench command -- 5 steps of the value arm (2 timed + 3 warm-up), 2+2 of")
P("the e2e arm, 5 of the pipelined arm, and the per-kernel roofline timings (11 direct launches of each of the four")
P("kernels and 11 sampling calls, each preceded by the 256 MB `FillFunctor<unsigned char>` L2 flush, which is not part")
P("of a step, and 11 `fill_` launches of 136 MB, the write-only reference for expand).  Per step: march, offsets, expand,")
P("composite fwd, ATen's mse / mean / mse-backward / fill, composite bwd.")
# 2. full profiles
import glob
rr = []
base = lambda name: re.sub(r"[<(].*", "", name).replace("void ", "")
refreshed = set()
# the capture of the final kernels of a config-2 step (if the step's kernels changed after the full capture) first:
# its rows replace the older ones of the same kernels
reps = sorted(glob.glob(f"gpurun_out/{tag}_step_kernels*.ncu-rep")) + sorted(glob.glob(f"gpurun_out/{tag}_kernels*.ncu-rep"))
for rep in reps:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    part = list(csv.reader(raw.splitlines()))
    ki = part[0].index("Kernel Name")
    if "_step_kernels" in rep:
        refreshed |= {base(r[ki]) for r in part[2:]}
    else:
        part = part[:2] + [r for r in part[2:] if base(r[ki]) not in refreshed]
    if not rr:
        rr = part
    elif part[0] == rr[0]:
        rr += part[2:]
    else:  # different metric sets: align on the first report's header
        idx = [part[0].index(h) if h in part[0] else None for h in rr[0]]
        rr += [[row[i] if i is not None else "" for i in idx] for row in part[2:]]
hdr, units = rr[0], rr[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio"]
P(f"\n# {tag}: `ncu --set full --clock-control none` per kernel (one launch each, from scripts/profile_kernels.py;")
P("the kernels of the config-2 step -- march, offsets, expand, composite fwd / bwd -- from the capture of the final build)\n")
seen = set()
for vals in rr[2:]:
    kn = re.sub(r"\(.*", "", vals[hdr.index("Kernel Name")]).replace("void ", "")
    if kn in seen: continue
    seen.add(kn)
    P(f"## `{kn}`\n"); P("| metric | value | unit |"); P("|---|---:|---|")
    for w in want:
        if w in hdr:
            i = hdr.index(w); P(f"| {w} | {vals[i]} | {units[i]} |")
    stalls = [(h, float(v.replace(',', ''))) for h, v in zip(hdr, vals) if h.startswith("smsp__pcsamp_warps_issue_stalled_") and "not_issued" not in h and v.replace(',', '').replace('.', '').isdigit()]
    ts = sum(v for _, v in stalls) or 1
    top = sorted(stalls, key=lambda x: -x[1])[:5]
    P("| top stall reasons (pc sampling) | " + ", ".join(f"{h.replace('smsp__pcsamp_warps_issue_stalled_', '')} {100*v/ts:.0f}%" for h, v in top) + " | |\n")
out.close()
