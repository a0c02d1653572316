"""Prints the cycle breakdown of one tile of the tensor-core backward kernel (CTA 0, second tile)."""
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
for _ in range(3): m.step(ro, rd, rgb, d)
L.call("mf_debug_profile", 1, None)
m.step(ro, rd, rgb, d); torch.cuda.synchronize()
buf = (C.c_longlong * 64)()
L.call("mf_debug_profile", 0, C.cast(buf, C.c_void_p))
t = list(buf)[:13]
names = ["copyW1+encode", "L1+epi1", "L2+epi2", "L3+epi3", "heads+dz3", "wgrad3", "dgrad3+scatter", "dH", "wgrad2", "dgrad2", "wgrad1", "dx", "-"]
for i in range(12):
    print(f"{names[i]:18s} {t[i+1]-t[i]:8d} cycles")
print("total", t[12]-t[0])

f = list(buf)[32:36]
print("fwd kernel: encode", f[1]-f[0], " mlp", f[2]-f[1], " store", f[3]-f[2], " total", f[3]-f[0])
