"""SASS opcode histogram of the tensor-core kernels in the built library: the mnemonics that prove tcgen05 /
TMEM / TMA (UTCHMMA = tcgen05.mma, UTCBAR = tcgen05.commit, LDTM / STTM = tcgen05.ld / st, UBLKCP = cp.async.bulk, SYNCS = mbarrier,
RED = red.global).   usage: python scripts/sass_histogram.py > profiles/r2_sass_histogram.txt"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "mipsfusion_b200/libmipsfusion_b200.so"
want = ("field_bwd_tc2_kernel", "field_fwd_tc3_kernel", "field_bwd_tc_kernel", "adam_sharded_kernel", "adam_pair_kernel")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
demangle = lambda s: subprocess.run(["cu++filt", s], capture_output=True, text=True).stdout.strip()
hist, name = {}, None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = m.group(1) if any(w in m.group(1) for w in want) else None
        if name:
            hist[name] = collections.Counter()
        continue
    if name:
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m:
            hist[name][m.group(1)] += 1
for name, h in hist.items():
    print(f"\n{demangle(name)[:150]}\n  {sum(h.values())} instructions")
    key = [o for o in ("UTCHMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UTMALDG", "SYNCS", "REDG", "RED", "ATOMG", "ATOMS", "MUFU", "LDG", "STG", "LDS", "STS", "BAR", "FFMA", "HFMA2") if o in h]
    print("  key: " + ", ".join(f"{o} {h[o]}" for o in key))
    print("  top: " + ", ".join(f"{o} {c}" for o, c in h.most_common(14)))
