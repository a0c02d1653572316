"""A/B of a library build (MIPSFUSION_B200_LIB=...): RandomOptimizer iteration, joint query 512^3 x 16, map step."""
import json, os, subprocess, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench, helpers as H
cfg, of = bench.build_model()
model = H.cuda_model(cfg, H.state_of(of))
dev = torch.device("cuda", 0)
t = bench.tracking_bench(model, cfg, dev, iters=5)
print("lib", os.environ.get("MIPSFUSION_B200_LIB", "default"))
print("  RO candidates/s", round(t["tracking_pose_candidates_per_s"]), "ms/iter", t.get("tracking_ms_per_ro_iteration"))
j = bench.joint_query_bench(dev)
print("  joint query s", j["joint_query_s"])
r = bench.render_full_bench(dev)
print("  render_full ms", r["render_full_img_ms"])
out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--quick", "--steps", "20", "--warmup", "5"], capture_output=True, text=True, env=os.environ)
print("  map step", out.stdout.strip().splitlines()[-1][:200])
