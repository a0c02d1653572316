"""Host timeline of the drop-in mapping step (JointEncoding.forward + loss.backward() + FusedAdam.step on the reference loop's
host batch): perf_counter around each section, no extra synchronisation (the only sync is the loss read-back at the end)."""
import os, sys, time
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
names = ["h2d", "slices", "forward", "loss ops", "backward", "opt.step", "loss read-back (sync)"]
acc = [0.0] * len(names)
def step(rec):
    t = [time.perf_counter()]
    b = host.to("cuda", non_blocking=True); t.append(time.perf_counter())
    a0, a1, a2, a3 = b[:, 0:3], b[:, 3:6], b[:, 6:9], b[:, 9:10]; t.append(time.perf_counter())
    ret = model(a0, a1, a2, a3); t.append(time.perf_counter())
    loss = tw["rgb_weight"] * ret["rgb_loss"] + tw["sdf_weight"] * ret["sdf_loss"] + tw["fs_weight"] * ret["fs_loss"]; t.append(time.perf_counter())
    loss.backward(); t.append(time.perf_counter())
    opt.step(zero_grad=True); t.append(time.perf_counter())
    v = float(loss.detach()); t.append(time.perf_counter())
    if rec:
        for i in range(len(names)): acc[i] += t[i + 1] - t[i]
    return v
for _ in range(10): step(False)
torch.cuda.synchronize(); t0 = time.perf_counter()
n = 100
for _ in range(n): step(True)
torch.cuda.synchronize(); tot = (time.perf_counter() - t0) / n
print(f"ms/step {tot*1e3:.3f}  = {4096/tot/1e6:.2f} M rays/s")
for nm, a in zip(names, acc): print(f"  {nm:24s} {a/n*1e6:8.1f} us")
print("one-launch optimiser step active:", opt.__dict__.get("_flat_state") is not None, "| decoder flat:", model.decoder.flat_storage() is not None)
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(50): step(False)
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(22)
