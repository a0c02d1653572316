"""Cycle stamps of the role-split tensor-core backward (CTA 0, second tile) and an A/B timing against the single-role kernel."""
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
for impl in (0, 2, 0, 2, 1, 0):
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
names = ["load idx, d_raw, e/grid words", "L1 + epi1 (H1 -> R1, scratch)", "L2 + epi2", "L3 + epi3 (logits, H3 -> scratch)", "softmax + dZ3 -> R1",
         "issue dgrad3 + products 0, 1", "dgrad3 wait", "dH -> R2 + issue dgrad2", "product 2 + dgrad2 wait", "dZ1 -> R1 + products 3, 4"]
for i in range(10):
    print(f"chain {names[i]:28s} {t[i+1]-t[i]:8d} cycles")
print("chain tile total", t[10] - t[0])
for L_ in range(3):
    print(f"wgrad layer {3 - L_}: ready at +{t[16+L_]-t[0]:d}, read-out {t[20+L_]-t[16+L_]:d} cycles")
print(f"scatter: dgrad3 seen at +{t[24]-t[0]:d}, released at +{t[26]-t[0]:d}, reductions take {t[25]-t[26]:d} cycles")

print("product 0, per quarter (h = 0 warp): [regs ready, buffer free, stores done, MMAs issued] relative to tile start")
for q in range(4):
    print("  q%d" % q, [t[32 + 4 * q + j] - t[0] for j in range(4)])
