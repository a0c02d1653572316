"""Cycle stamps of the four-role tensor-core backward experiment (impl 3; CTA 0, second tile) and an A/B timing against the
three-role (0, default) and single-role (1) kernels."""
import ctypes as C, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench, helpers as H
from mipsfusion_b200 import _lib as L
from mipsfusion_b200.mapper import FusedMapper
cfg, of = bench.build_model()
model = H.cuda_model(cfg, H.state_of(of))
ro, rd, rgb, d, _ = bench.make_inputs(0)
ro, rd, rgb, d = (t.cuda().contiguous() for t in (ro, rd, rgb, d))
m = FusedMapper(model)
L.call("mf_debug_kernel_timer", 1)
for impl in (3, 0, 1, 3):
    L.call("mf_set_bwd_impl", impl)
    for _ in range(3): m.step(ro, rd, rgb, d)
    ts = []
    for _ in range(10):
        m.step(ro, rd, rgb, d)
        ms = C.c_float(); L.call("mf_debug_kernel_ms", 1, C.byref(ms)); ts.append(ms.value)
    print("bwd impl", impl, "kernel ms", sorted(ts)[len(ts) // 2], "min", min(ts), "tc error", L.lib().mf_tc_check_error(), flush=True)
L.call("mf_debug_kernel_timer", 0)
L.call("mf_debug_profile", 1, None)
m.step(ro, rd, rgb, d); torch.cuda.synchronize()
buf = (C.c_longlong * 64)()
L.call("mf_debug_profile", 0, C.cast(buf, C.c_void_p))
t = list(buf)
names = ["load idx, d_raw, e/grid words (+ rd3, rd2)", "L1 + epi1 (H1 -> R1, scratch)", "L2 (+ rd4) + epi2", "L3 + epi3 + logit exchange", "softmax + U + dZ3 -> R1",
         "z3, dgrad3 wait, rd0", "dH -> R2, zh, issue dgrad2", "dgrad2 wait", "dZ1 -> R1, z1"]
for i in range(9):
    print(f"chain {names[i]:44s} {t[i+1]-t[i]:8d} cycles")
print("chain tile total", t[9] - t[0])
cn = ["z3 seen", "P0 issued", "P1 issued", "read-out 3 done", "zh seen", "P2 issued", "read-out 2 done", "z1 seen", "P3 issued", "P4 issued", "read-out 1 done"]
for j, n in enumerate(cn):
    print(f"copy {n:18s} at +{t[16+j]-t[0]:d}")
print(f"scatter: released at +{t[28]-t[0]:d}, reductions take {t[29]-t[28]:d} cycles")
