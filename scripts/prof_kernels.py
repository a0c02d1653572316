"""Runs one path a few times so that ncu can capture its dominant kernel:
    python scripts/prof_kernels.py ro|jq|go|map [iters]
ro: RandomOptimizer scoring at 1024 candidates x 2048 pixels (field_fwd_tc3_kernel<SrcRO, EpiAbsSdf, SDF_ONLY>)
jq: joint query, 128^3 grid x 16 submaps (field_fwd_tc2_kernel<SrcJoint, ...>)
go: gradient pose refinement, 1000 rays x 75 samples with ray gradients (the pose-gradient backward)
map: the C1 map step (FusedMapper.step)"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench, helpers as H
mode = sys.argv[1]
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda", 0)
if mode == "ro":
    cfg, of = bench.build_model()
    model = H.cuda_model(cfg, H.state_of(of))
    print(bench.tracking_bench(model, cfg, dev, iters=iters))
elif mode == "jq512":
    r = bench.joint_query_bench(dev, res=512)
    print({k: v for k, v in r.items() if not isinstance(v, dict)})
elif mode == "jq":
    print(bench.joint_query_bench(dev, res=128))
elif mode == "go":                       # fused gradient pose refinement: ray-gradient-only tensor-core backward
    import mipsfusion_b200 as mf
    cfg = H.make_config(bench.HASH, n_samples_d=50, n_range_d=25)
    cfg["tracking"] = {"lr_rot": 1e-3, "lr_trans": 1e-3, "wait_iters": 100, "best": True}
    of = H.oracle_field(cfg); model = H.cuda_model(cfg, H.state_of(of))
    rays7, _, poses, _ = H.synth_batch_packed(1000, seed=2)
    ref = mf.FusedPoseRefiner(model)
    pose, st = ref.refine(poses[0], rays7[:, :3].to(dev), rays7[:, 3:6].to(dev), rays7[:, 6].to(dev), iters)
    torch.cuda.synchronize()
    print("go ok", pose.cpu().numpy()[:3, 3])
elif mode == "map":
    from mipsfusion_b200.mapper import FusedMapper
    cfg, of = bench.build_model()
    model = H.cuda_model(cfg, H.state_of(of))
    rays_o, rays_d, rgb, d, _ = [t.to(dev).contiguous() for t in bench.make_inputs(0)]
    mapper = FusedMapper(model)
    for _ in range(iters):
        losses = mapper.step(rays_o, rays_d, rgb, d.reshape(-1))
    torch.cuda.synchronize()
    print("map ok", losses.tolist())
