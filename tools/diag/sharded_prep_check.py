"""2+ GPUs (torchrun): the observation-sharded prep through real NCCL collectives against the replicated prep.
   torchrun --nproc-per-node 2 tools/diag/sharded_prep_check.py [workload] [N]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry
import torch, torch.distributed as dist
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
jp = entry.load_package()
from jointposteriors_jl_b200 import workloads, distributed as D
from jointposteriors_jl_b200.model import Context, JointPosterior
name = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
N = int(float(sys.argv[2])) if len(sys.argv) > 2 else 300000
wl = workloads.WORKLOADS[name](N=N)
ctx = Context.get(lr)
ctx.use_stream(torch.cuda.current_stream(dev).cuda_stream)
M = jp.Model(wl["params"], device=lr)
dd = ctx.upload(wl["data"])
x, U, neg_min = jp.mode(M, dd)
grid = ctx.grid(0, U.shape[1], wl["level"])
Mtot = int(jp.lib().jp_grid_size(grid))
b, e = D.shard_bounds(Mtot, rank, world)
res = {}
for mode, mb in (("replicated", "1e9"), ("sharded", "0")):
    os.environ["JP_SHARDED_PREP_MIN_MB"] = mb
    post = JointPosterior(M, dd, grid, x, U, neg_min, node_range=(b, e))
    loc = D.CudaLocal(post)
    for _ in range(3):
        D.fit_sharded(loc)
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    for _ in range(10):
        D.fit_sharded(loc)
    torch.cuda.synchronize(); dist.barrier()
    ms = (time.perf_counter() - t0) * 100
    res[mode] = (post.density.copy(), post.logdens.copy(), ms, post.diagnostics)
    assert loc.last_prep == mode, (loc.last_prep, mode, post.diagnostics)
    post.free()
d0, l0, t0_, g0 = res["replicated"]; d1, l1, t1_, g1 = res["sharded"]
err = np.max(np.abs(d0 - d1)) / np.max(np.abs(d0))
print("rank %d nodes %d: replicated %.3f ms  sharded %.3f ms  density diff %.2e  logdens diff %.2e  NC %d/%d" %
      (rank, e - b, t0_, t1_, err, np.max(np.abs(l0 - l1)), g0["series_terms"], g1["series_terms"]), flush=True)
assert err < 1e-7
dist.barrier(); dist.destroy_process_group()
