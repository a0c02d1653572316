"""cProfile of the drop-in mapping step (JointEncoding.forward + backward + FusedAdam.step) to see host overhead."""
import cProfile, pstats, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench, helpers as H
import mipsfusion_b200 as mf
cfg, of = bench.build_model()
model = H.cuda_model(cfg, H.state_of(of))
ro, rd, rgb, d, _ = bench.make_inputs(0)
host = torch.cat([ro, rd, rgb, d], -1).contiguous().pin_memory()
opt = mf.create_map_optimizer(model, 1e-2, 1e-2)
tw = cfg["training"]
def step():
    b = host.to("cuda", non_blocking=True)
    ret = model(b[:, 0:3], b[:, 3:6], b[:, 6:9], b[:, 9:10])
    loss = tw["rgb_weight"] * ret["rgb_loss"] + tw["sdf_weight"] * ret["sdf_loss"] + tw["fs_weight"] * ret["fs_loss"]
    loss.backward()
    opt.step(zero_grad=True)
    return float(loss.detach())
for _ in range(5): step()
torch.cuda.synchronize(); t = time.perf_counter()
for _ in range(50): step()
torch.cuda.synchronize(); print("ms/step", (time.perf_counter() - t) / 50 * 1e3)
pr = cProfile.Profile(); pr.enable()
for _ in range(50): step()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)

from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(10): step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=60))
