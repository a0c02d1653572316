"""Per-kernel table of the last marching-cubes call in an ncu launch list (gpu__time_duration + dram bytes).  Usage: python scripts/mc_launch_table.py gpurun_out/r2_mc_launches.csv"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
start = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[start]
ki, mi, vi, idi, ui = (hdr.index(x) for x in ("Kernel Name", "Metric Name", "Metric Value", "ID", "Metric Unit"))
d = collections.OrderedDict()
for r in rows[start + 1:]:
    if len(r) > vi:
        d.setdefault(r[idi], {"k": r[ki]})[r[mi]] = (float(r[vi].replace(",", "")), r[ui])
ids = list(d)
nodes = [i for i in ids if "mc_nodes" in d[i]["k"]]
last = nodes[-2]                     # the last 512^3 call (the very last one is the 128^3 sub-volume)
tot = 0.0
for i in ids[ids.index(last):]:
    e = d[i]
    if "mc_nodes" in e["k"] and i != last:
        break
    t = e["gpu__time_duration.sum"]
    us = t[0] / 1e3 if t[1] == "ns" else t[0]
    tot += us
    name = e["k"].replace("<unnamed>::", "").split("(")[0][:48]
    print(f"{name:48s} {us:9.1f} us   dram rd {e['dram__bytes_read.sum'][0] / 1e6:8.1f} MB  wr {e['dram__bytes_write.sum'][0] / 1e6:8.1f} MB")
print(f"{'sum of kernels (cold caches, serialised)':48s} {tot:9.1f} us")
