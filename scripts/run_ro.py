"""Runs a few RandomOptimizer scoring passes at the BASELINE tracking shape (for profiling)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench, helpers as H
cfg, of = bench.build_model()
model = H.cuda_model(cfg, H.state_of(of))
print(bench.tracking_bench(model, cfg, torch.device("cuda", 0), iters=int(sys.argv[1]) if len(sys.argv) > 1 else 5))
