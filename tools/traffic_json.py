"""profiles/traffic_<tag>.json: DRAM bytes per kernel launch and per read pair from the ncu reports of tools/profile_round.sh."""
import csv, json, os, subprocess, sys
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return rows[0], rows[1], rows[2:]


def tobytes(v, u):
    return float(v.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(u, 1)


res = {}
# pairs in the captured launch: map_probe 100000 -> seeding sub-chunks of 149 796 orientations, SW / pileup 100 000 pairs
pairs = {"seed_gen_kernel": 149796 / 4, "pair_select_kernel": 149796 / 4, "sw_fill_kernel": 100000, "sw_traceback_kernel": 100000,
         "pu_count_kernel": 100000, "pu_emit_kernel": 100000}
h, u, rows = raw(os.path.join(root, "gpurun_out", "prof_%s.ncu-rep" % tag))
for r in rows:
    name = r[h.index("Kernel Name")].split("(")[0].replace("void ", "").split("<")[0]
    if name in res:
        continue
    rd = tobytes(r[h.index("dram__bytes_read.sum")], u[h.index("dram__bytes_read.sum")])
    wr = tobytes(r[h.index("dram__bytes_write.sum")], u[h.index("dram__bytes_write.sum")])
    res[name] = {"dram_read_bytes": rd, "dram_write_bytes": wr,
                 "duration": r[h.index("gpu__time_duration.sum")] + " " + u[h.index("gpu__time_duration.sum")], "pairs_in_launch": pairs.get(name)}
h, u, rows = raw(os.path.join(root, "gpurun_out", "prof_win_%s.ncu-rep" % tag))
r = rows[0]
res["seed_win_kernel"] = {"dram_read_bytes": tobytes(r[h.index("dram__bytes_read.sum")], u[h.index("dram__bytes_read.sum")]),
                          "dram_write_bytes": tobytes(r[h.index("dram__bytes_write.sum")], u[h.index("dram__bytes_write.sum")]),
                          "duration": r[h.index("gpu__time_duration.sum")] + " " + u[h.index("gpu__time_duration.sum")],
                          "pairs_in_launch": 20000, "replay": "application"}
per_pair = {k: (v["dram_read_bytes"] + v["dram_write_bytes"]) / v["pairs_in_launch"] for k, v in res.items() if v.get("pairs_in_launch")}
json.dump({"source": "gpurun_out/prof_%s.ncu-rep (ncu --set full, kernel replay, tools/map_probe.py 100000 2) and prof_win_%s.ncu-rep "
                     "(seed_win_kernel, --replay-mode application, 20000 pairs); tools/profile_round.sh" % (tag, tag),
           "kernels": res, "dram_bytes_per_pair": per_pair}, open(os.path.join(root, "profiles", "traffic_%s.json" % tag), "w"), indent=1)
print(json.dumps(per_pair, indent=1))
