"""Per-CTA lifetimes (%globaltimer) of the forward and backward tensor-core kernels of one C1 map step: launch ramp, tail, imbalance."""
import ctypes as C, os, sys
import numpy as np, torch
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
buf = (C.c_longlong * 1024)()
L.call("mf_debug_profile_all", C.cast(buf, C.c_void_p), 1024)
L.call("mf_debug_profile", 0, None)
t = np.array(list(buf), dtype=np.int64)
n = int(torch.cuda.get_device_properties(0).multi_processor_count)
for name, off in (("field_fwd_tc3", 64), ("field_bwd_tc2", 576)):
    s, e = t[off:off + 2 * n:2], t[off + 1:off + 1 + 2 * n:2]
    ok = (s > 0) & (e > 0)
    s, e = s[ok], e[ok]
    t0 = s.min()
    dur = e - s
    print(f"{name}: {ok.sum()} CTAs; kernel span {(e.max() - t0) / 1e3:.1f} us; starts spread {(s.max() - t0) / 1e3:.1f} us; "
          f"ends from {(e.min() - t0) / 1e3:.1f} to {(e.max() - t0) / 1e3:.1f} us; CTA lifetime min / median / max "
          f"{dur.min() / 1e3:.1f} / {np.median(dur) / 1e3:.1f} / {dur.max() / 1e3:.1f} us")
    order = np.argsort(e)
    print("   end times (us) by decile:", [round(float((e[order[int(q * (len(e) - 1))]] - t0) / 1e3), 1) for q in np.linspace(0, 1, 11)])
