"""profiles/kernel_metrics_<tag>.json: what bench.py quotes next to the live timings of the two seeding kernels (DRAM bytes per
read pair, issue-slot utilisation, warp instructions per read orientation), taken from the ncu reports of tools/profile_round.sh.
These are STATIC numbers of the named capture, labelled as such in the bench line; the timings beside them are live."""
import csv, json, os, subprocess, sys
tag = sys.argv[1] if len(sys.argv) > 1 else "r2a"
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return rows[0], rows[1], rows[2:]


def val(r, h, u, k):
    v = float(r[h.index(k)].replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(u[h.index(k)], 1)


out = {"source": "profiles/kernels_%s.md = gpurun_out/prof_%s.ncu-rep (ncu --set full --clock-control none, kernel replay, "
                 "tools/map_probe.py 70000 2: 280 000 read orientations per seeding launch) and prof_win_%s.ncu-rep (seed_win_kernel, "
                 "--replay-mode application, 20 000 pairs = 80 000 orientations)" % (tag, tag, tag), "kernels": {}}
for rep, kern, orient in (("prof_%s.ncu-rep" % tag, "seed_gen_kernel", 280000), ("prof_win_%s.ncu-rep" % tag, "seed_win_kernel", 80000)):
    alt = os.path.join(root, "gpurun_out", "prof_gen_%s.ncu-rep" % tag)
    if kern == "seed_gen_kernel" and os.path.exists(alt):
        rep = os.path.basename(alt)
    h, u, rows = raw(os.path.join(root, "gpurun_out", rep))
    r = next(r for r in rows if kern in r[h.index("Kernel Name")])
    out["kernels"][kern] = {
        "orientations_in_launch": orient,
        "dram_bytes_per_pair": (val(r, h, u, "dram__bytes_read.sum") + val(r, h, u, "dram__bytes_write.sum")) / (orient / 4),
        "issue_active_frac": val(r, h, u, "smsp__issue_active.avg.pct_of_peak_sustained_active") / 100,
        "alu_pipe_active_frac": val(r, h, u, "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active") / 100,
        "warp_instr_per_orientation": val(r, h, u, "smsp__inst_executed.sum") / orient,
        "threads_per_instr": val(r, h, u, "smsp__thread_inst_executed_per_inst_executed.ratio"),
        "duration_under_ncu_ms": val(r, h, u, "gpu__time_duration.sum")}
json.dump(out, open(os.path.join(root, "profiles", "kernel_metrics_r2.json"), "w"), indent=1)
print(json.dumps(out["kernels"], indent=1))
