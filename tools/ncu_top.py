"""Print the top stall-sampled SASS lines of a kernel from an ncu report: ncu_top.py rep [launch_idx] [n]"""
import csv, subprocess, sys
rep = sys.argv[1]; idx = sys.argv[2] if len(sys.argv) > 2 else "0"; n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", idx, "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
data = []
for r in rows:
    if len(r) > 6 and r[0].startswith("0x"):
        data.append(r)
    elif len(r) >= 2 and r[0] == "Kernel Name":
        print(r[1][:120])
tot = sum(float(r[2]) for r in data)
print("total samples", tot, "instrs", len(data))
for i, r in enumerate(data):
    r.append(i)
for r in sorted(data, key=lambda r: -float(r[2]))[:n]:
    print("%6d %5.1f%%  #%4d exec=%s  %s" % (float(r[2]), 100 * float(r[2]) / tot, r[-1], r[5], r[1].strip()[:100]))
