"""Per-piece timing of the data-parallel update (run under torchrun): barrier, sharded Adam kernels, counts all-reduce."""
import ctypes as C, os, sys
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench, helpers as H
from mipsfusion_b200 import _lib as L
from mipsfusion_b200.mapper import FusedMapper
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
os.environ["NCCL_DEBUG"] = "WARN"
dist.init_process_group("nccl", device_id=dev)
cfg, of = bench.build_model()
m = FusedMapper(H.cuda_model(cfg, H.state_of(of)), group=dist.group.WORLD, peer_memory=True)
a = m.arena
bases = (C.c_uint64 * a.world)(*a.peer_bases)
st = L.stream()
MC = [0]
def timeit(fn, n=30):
    for _ in range(5): fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
def grid_kernel(clear=True):
    L.call("mf_adam_step_sharded", bases, a.world, a.rank, a.offsets["p_grid"], a.offsets["g_grid0"], a.offsets["g_grid1"] if clear else -1,
           L.ptr(m.m_grid), L.ptr(m.v_grid), a.sizes["p_grid"], 1e-2, 0.9, 0.99, 1e-15, 0.0, 1, MC[0], st)
def mlp_kernel():
    L.call("mf_adam_step_sharded", bases, a.world, a.rank, a.offsets["p_mlp"], a.offsets["g_mlp0"], a.offsets["g_mlp1"],
           L.ptr(m.m_mlp), L.ptr(m.v_mlp), a.sizes["p_mlp"], 1e-2, 0.9, 0.99, 1e-8, 1e-6, 1, MC[0], st)
cnt = torch.zeros(2, device=dev, dtype=torch.int64)
res = {"barrier": timeit(a.barrier), "grid_kernel+clear": timeit(grid_kernel), "grid_kernel": timeit(lambda: grid_kernel(False)),
       "mlp_kernel": timeit(mlp_kernel), "counts_allreduce": timeit(lambda: dist.all_reduce(cnt)),
       "barrier+grid+barrier": timeit(lambda: (a.barrier(), grid_kernel(), a.barrier())),
       "nccl_allreduce_grid": timeit(lambda: dist.all_reduce(m.g_grid))}
MC[0] = a.multicast_base
res["multicast_base"] = float(a.multicast_base != 0)
if a.multicast_base:
    res["mc_grid_kernel+clear"] = timeit(grid_kernel); res["mc_mlp_kernel"] = timeit(mlp_kernel)
if rank == 0:
    print("world", world, {k: round(v, 1) for k, v in res.items()}, "us", flush=True)
dist.destroy_process_group()
