"""Per-source-line instruction counts and stall samples of one kernel from an ncu report captured with --import-source on:
    python scripts/ncu_lines.py report.ncu-rep kernel_regex [top]"""
import csv, subprocess, sys, collections
rep, rx = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 45
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + rx],
                     capture_output=True, text=True).stdout
cur, hdr = None, None
agg = collections.OrderedDict()
seen_fn = 0
for r in csv.reader(raw.splitlines()):
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]; continue
    if len(r) == 2 and r[0] == "Function Name":
        continue
    if r and r[0] == "Line No":
        hdr = r; iI = hdr.index("Instructions Executed"); iS = hdr.index("# Samples"); continue
    if hdr and r and r[0] not in ("",) and r[2] == "-":
        try:
            key = (cur, int(r[0]), r[1].strip()[:90])
            a = agg.setdefault(key, [0, 0]); a[0] += int(r[iI]); a[1] += int(r[iS])
        except ValueError:
            pass
tot_i = sum(v[0] for v in agg.values()); tot_s = sum(v[1] for v in agg.values())
print(f"total warp instructions {tot_i}, samples {tot_s}")
print("by file:")
byf = collections.Counter(); bys = collections.Counter()
for (f, l, s), v in agg.items():
    byf[f] += v[0]; bys[f] += v[1]
for f, n in byf.most_common():
    print(f"  {f:28s} inst {100*n/tot_i:5.1f}%  samples {100*bys[f]/max(tot_s,1):5.1f}%")
print(f"top {top} lines by samples:")
for (f, l, s), v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"  {f}:{l:<4d} inst {100*v[0]/tot_i:5.2f}% samp {100*v[1]/max(tot_s,1):5.2f}%  {s}")
