"""Per-source-line totals (stall samples, warp instructions, avg threads) of a kernel from an ncu report captured with
--import-source on:  ncu_lines.py rep [top_n] [file_substr]"""
import csv, subprocess, sys
rep = sys.argv[1]; n = int(sys.argv[2]) if len(sys.argv) > 2 else 40; want = sys.argv[3] if len(sys.argv) > 3 else ""
extra = []
for a in sys.argv:
    if a.startswith("--kernel="): extra = ["--kernel-name", "regex:" + a.split("=", 1)[1]]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"] + extra, capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur = None; lines = []
for r in rows:
    if len(r) >= 2 and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if len(r) > 8 and r[0] not in ("", "Line No") and r[0].isdigit():
        try: lines.append((cur, int(r[0]), r[1].strip(), float(r[4]), float(r[7]), float(r[8])))
        except ValueError: pass
tots = sum(l[3] for l in lines); toti = sum(l[4] for l in lines); tott = sum(l[5] for l in lines)
print("total samples %.0f warp-instr %.3e thread-instr %.3e (avg thr %.1f)" % (tots, toti, tott, tott / max(toti, 1)))
sel = [l for l in lines if want in l[0]]
key = (lambda l: -l[4]) if "--by-instr" in sys.argv else (lambda l: -l[3])
for l in sorted(sel, key=key)[:n]:
    print("%-12s %5d  smp %5.1f%%  ins %5.1f%%  thr %4.1f  %s" % (l[0][:12], l[1], 100 * l[3] / tots, 100 * l[4] / toti, l[5] / max(l[4], 1), l[2][:110]))
if "--ranges" in sys.argv:
    rg = [("setup", 234, 335), ("A", 336, 398), ("B1", 399, 456), ("B2", 457, 506), ("Bhash", 507, 576), ("MSA", 577, 649), ("C-flag", 650, 731), ("C-sweep", 732, 817),
          ("C-compact", 818, 837), ("C-serial", 838, 964), ("D", 965, 1040), ("E", 1041, 1102), ("F", 1103, 1148), ("G-prune", 1149, 1191),
          ("radix", 139, 192), ("scan", 130, 137), ("kmerid", 100, 128)]
    for name, a, b in rg:
        s = [l for l in lines if l[0] == "seed.cuh" and a <= l[1] <= b]
        print("%-10s smp %5.1f%% ins %5.1f%% thr %4.1f" % (name, 100 * sum(l[3] for l in s) / tots, 100 * sum(l[4] for l in s) / toti, sum(l[5] for l in s) / max(1, sum(l[4] for l in s))))
    s = [l for l in lines if l[0] != "seed.cuh"]
    print("%-10s smp %5.1f%% ins %5.1f%%" % ("other files", 100 * sum(l[3] for l in s) / tots, 100 * sum(l[4] for l in s) / toti))
