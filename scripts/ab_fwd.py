"""A/B of the forward kernels (decoder impl 0 = producer/consumer, 3 = dual pipeline, 2 = single pipeline) on the
tracking and joint-query shapes."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench, helpers as H
from mipsfusion_b200 import _lib as L
dev = torch.device("cuda", 0)
cfg, of = bench.build_model()
model = H.cuda_model(cfg, H.state_of(of))
for impl in [int(a) for a in sys.argv[1:]] or [0, 3, 2]:
    L.call("mf_set_decoder_impl", impl)
    t = bench.tracking_bench(model, cfg, dev)
    j = bench.joint_query_bench(dev, res=256)
    print("impl", impl, "RO ms/iter", round(t["tracking_ms_per_ro_iteration"], 4), " joint 256^3 s", round(j["joint_query_s"], 4), flush=True)
