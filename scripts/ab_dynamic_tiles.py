"""A/B of the forward kernel's tile scheduling on one box: dynamic (global tile counter) vs static striding, for the C1 map-step
forward (CUDA events around the kernel) and one RandomOptimizer iteration at 1024 candidates x 2048 pixels."""
import ctypes as C, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench, helpers as H
from mipsfusion_b200 import _lib as L
from mipsfusion_b200.mapper import FusedMapper
dev = torch.device("cuda", 0)
cfg, of = bench.build_model()
model = H.cuda_model(cfg, H.state_of(of))
ro, rd, rgb, d, _ = bench.make_inputs(0)
ro, rd, rgb, d = (t.cuda().contiguous() for t in (ro, rd, rgb, d))
m = FusedMapper(model)
L.call("mf_debug_kernel_timer", 1)
for on in (1, 0, 1, 0):
    L.call("mf_set_dynamic_tiles", on)
    for _ in range(3): m.step(ro, rd, rgb, d)
    ts = []
    for _ in range(20):
        m.step(ro, rd, rgb, d)
        ms = C.c_float(); L.call("mf_debug_kernel_ms", 0, C.byref(ms)); ts.append(ms.value)
    print("dynamic tiles", on, "map forward kernel ms: median", sorted(ts)[len(ts) // 2], "min", min(ts), flush=True)
L.call("mf_debug_kernel_timer", 0)
model2 = H.cuda_model(cfg, H.state_of(of))
for on in (1, 0, 1, 0):
    L.call("mf_set_dynamic_tiles", on)
    r = bench.tracking_bench(model2, cfg, dev, iters=5)
    print("dynamic tiles", on, "RO ms / iteration", r["tracking_ms_per_ro_iteration"], flush=True)
L.call("mf_set_dynamic_tiles", 1)
