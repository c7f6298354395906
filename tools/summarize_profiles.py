"""Summarises gpurun_out ncu artefacts into profiles/ (tracked):  python tools/summarize_profiles.py <round-tag>"""
import collections, csv, os, subprocess, sys
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.makedirs(os.path.join(root, "profiles"), exist_ok=True)

# ---- launch list --------------------------------------------------------------------------------------
lp = os.path.join(root, "gpurun_out", "launches_%s.csv" % tag)
rows = [r for r in csv.reader(open(lp, errors="ignore")) if len(r) > 5]
hdr = next(r for r in rows if "Kernel Name" in r)
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows:
    if r is hdr or len(r) <= iv or r[hdr.index("Metric Name")] != "gpu__time_duration.sum":
        continue
    name = r[ik].split("(")[0]
    v = float(r[iv].replace(",", ""))
    u = r[iu]
    us = v / 1e3 if u in ("ns", "nsecond") else v if u in ("us", "usecond") else v * 1e3
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += us
tot = sum(a[1] for a in agg.values())
with open(os.path.join(root, "profiles", "launches_%s.md" % tag), "w") as f:
    f.write("# Kernel launch list (%s)\n\n`ncu --metrics gpu__time_duration.sum --clock-control none` over\n"
            "`python bench.py --steps 2 --warmup 3 --pairs 100000 --no-cpu-baseline --no-msa` (cold-cache, serialised: compare shares).\n\n"
            "| kernel | launches | total ms | share |\n|---|---:|---:|---:|\n" % tag)
    for k, (n, us) in sorted(agg.items(), key=lambda x: -x[1][1]):
        f.write("| `%s` | %d | %.3f | %.1f%% |\n" % (k, n, us / 1e3, 100 * us / tot))
    f.write("\nTotal %.1f ms over %d launches.\n" % (tot / 1e3, sum(a[0] for a in agg.values())))

# ---- per-kernel metrics -------------------------------------------------------------------------------
def load(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(out.splitlines()))


rep = os.path.join(root, "gpurun_out", "prof_%s.ncu-rep" % tag)
rows = load(rep)
hdr = rows[0]
# seed_win_kernel rewrites its input in place: captured separately under application replay
rep_win = os.path.join(root, "gpurun_out", "prof_win_%s.ncu-rep" % tag)
rows_win = load(rep_win) if os.path.exists(rep_win) else []
want = [("gpu__time_duration.sum", "duration"), ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
        ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "ALU(DPX) pipe active %"),
        ("sm__inst_executed_pipe_alu.sum", "ALU-pipe instr"), ("smsp__inst_executed.sum", "warp instr"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("smsp__thread_inst_executed_per_inst_executed.ratio", "threads/instr"),
        ("launch__registers_per_thread", "regs"), ("launch__shared_mem_per_block", "smem/block"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_sb"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_sb"),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait"),
        ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math_pipe"),
        ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall branch")]
units = rows[1]
seen = set()
with open(os.path.join(root, "profiles", "kernels_%s.md" % tag), "w") as f:
    f.write("# Per-kernel ncu --set full summary (%s)\n\nFrom `gpurun_out/prof_%s.ncu-rep` (`ncu --set full --clock-control none --import-source on`, "
            "tools/map_probe.py 70000 2: the mapping pipeline at 70k pairs). Times under ncu are not bench values.\n" % (tag, tag))
    rep_gen = os.path.join(root, "gpurun_out", "prof_gen_%s.ncu-rep" % tag)  # optional: the stream kernel captured on its own
    rows_gen = load(rep_gen) if os.path.exists(rep_gen) else []
    for rws, note in ((rows_gen, " (captured on its own: `-k regex:seed_gen -s 1 -c 1`, same probe)"), (rows, ""), (rows_win, " (`--replay-mode application`, tools/map_probe.py 20000 2: the kernel rewrites its "
                                               "multi-GB input slot in place, which kernel replay does not restore)")):
        if not rws:
            continue
        h, un = rws[0], rws[1]
        for r in rws[2:]:
            name = r[h.index("Kernel Name")].split("(")[0]
            if name in seen:
                continue
            seen.add(name)
            f.write("\n## `%s`%s\n\n| metric | value |\n|---|---|\n" % (name, note))
            for m, label in want:
                if m in h:
                    f.write("| %s (`%s`) | %s %s |\n" % (label, m, r[h.index(m)], un[h.index(m)]))
print(open(os.path.join(root, "profiles", "launches_%s.md" % tag)).read())
