"""Warp-instruction / stall-sample share of source-line ranges of one kernel in an ncu report (--import-source on):
   ncu_ranges.py rep kernel_regex file name:a-b [name:a-b ...]"""
import csv, subprocess, sys
rep, kern, fname = sys.argv[1], sys.argv[2], sys.argv[3]
rg = []
for a in sys.argv[4:]:
    n, r = a.split(":"); lo, hi = r.split("-"); rg.append((n, int(lo), int(hi)))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name", "regex:" + kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur = None; lines = []
for r in rows:
    if len(r) >= 2 and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if len(r) > 8 and r[0].isdigit():
        try: lines.append((cur, int(r[0]), float(r[4]), float(r[7]), float(r[8])))
        except ValueError: pass
tots = sum(l[2] for l in lines); toti = sum(l[3] for l in lines); tott = sum(l[4] for l in lines)
print("total samples %.0f warp-instr %.4e (avg thr %.1f)" % (tots, toti, tott / max(toti, 1)))
acc = 0
for n, lo, hi in rg:
    s = [l for l in lines if l[0] == fname and lo <= l[1] <= hi]
    si = sum(l[3] for l in s); acc += si
    print("%-12s %5d-%-5d smp %5.1f%% ins %5.1f%% thr %4.1f" % (n, lo, hi, 100 * sum(l[2] for l in s) / tots, 100 * si / toti, sum(l[4] for l in s) / max(1, si)))
print("rest ins %5.1f%%" % (100 * (toti - acc) / toti))
