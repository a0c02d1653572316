"""Cycle breakdown of one tile of the single-pipeline tensor-core forward kernel (decoder impl 2; CTA 0, second tile)."""
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
impl = int(sys.argv[1]) if len(sys.argv) > 1 else 0
L.call("mf_set_decoder_impl", impl)
m = FusedMapper(model)
for _ in range(3): m.step(ro, rd, rgb, d, update=False)
L.call("mf_debug_profile", 1, None)
m.step(ro, rd, rgb, d, update=False); torch.cuda.synchronize()
buf = (C.c_longlong * 64)()
L.call("mf_debug_profile", 0, C.cast(buf, C.c_void_p))
f = list(buf)[32:36]
b = list(buf)
if impl == 0:
    print("kernel start -> first tile", b[16] - b[29], " producer tile starts (deltas):", [b[17 + i] - b[16 + i] for i in range(0, 9)], " total", b[28] - b[29])
    pn = ["point+wait empty", "freq", "gathers", "wait_st+arrive"]
    print("producer:", {pn[i]: b[41 + i] - b[40 + i] for i in range(4)}, "total", b[44] - b[40])
    cn = ["wait full", "L1 round", "epi1", "L2 round", "epi2", "L3 round", "epi3+logits", "heads", "store"]
    print("consumer:", {cn[i]: b[49 + i] - b[48 + i] for i in range(9)}, "total", b[57] - b[48])
print("fwd kernel (512 threads, 1 tile): encode", f[1]-f[0], " mlp", f[2]-f[1], " store", f[3]-f[2], " total", f[3]-f[0])
m.timing = {}
for _ in range(5): m.step(ro, rd, rgb, d, update=False)
torch.cuda.synchronize()
print({k: sum(a.elapsed_time(b) for a, b in v) / len(v) for k, v in m.timing.items()})
