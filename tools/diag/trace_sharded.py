"""Stage timeline of one node-sharded step (jp_fit_p2p + jp_marginal_coords_p2p) on rank 0 of a torchrun job:
   python -m torch.distributed.run --nproc-per-node N tools/diag/trace_sharded.py [cfg3|cfg4|cfg5]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry
import torch
import torch.distributed as dist
jp = entry.load_package()
from jointposteriors_jl_b200 import workloads, distributed as D
from jointposteriors_jl_b200.model import Context, JointPosterior

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
name = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
wl = workloads.cfg3_logistic(N=100_000 * world) if name == "cfg3" else workloads.WORKLOADS[name]()
ctx = Context.get(lr)
ctx.use_stream(torch.cuda.current_stream(dev).cuda_stream)
M = jp.Model(wl["params"], device=lr)
obs_mode = os.environ.get("SHARD", "nodes") == "obs"
coords = list(range(wl["d"]))
if obs_mode:
    from jointposteriors_jl_b200.model import DeviceData
    N = wl["data"].records()[0].shape[0]
    dd = DeviceData(ctx, wl["data"], rows=D.row_slice(N, rank, world)[:2])
    comm = D.comm_for(ctx, 0, None, min_bulk=1 << 24)
    x, U, neg_min = D.mode_p2p(M, dd, comm)
    grid = ctx.grid(0, U.shape[1], wl["level"])
    post = JointPosterior(M, dd, grid, x, U, neg_min)
    sp = D.ObsShardedPosterior(post, comm, None)
else:
    dd = ctx.upload(wl["data"])
    x, U, neg_min = jp.mode(M, dd)
    grid = ctx.grid(0, U.shape[1], wl["level"])
    Mtot = int(jp.lib().jp_grid_size(grid))
    post = JointPosterior(M, dd, grid, x, U, neg_min, node_range=D.shard_bounds(Mtot, rank, world))
    loc = D.CudaLocal(post)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
tok = torch.zeros(1, device=dev)
for it in range(6):
    flush.zero_()
    dist.all_reduce(tok)
    torch.cuda.synchronize()
    ctx.trace(True)
    if obs_mode:
        sp.refit()
        sp.marginals(coords)
    else:
        D.fit_sharded(loc)
        D.marginals_sharded(loc, coords)
    tr = ctx.trace_dump()
    ctx.trace(False)
if rank == 0:
    print("%s on %d ranks, rank 0: last of 6 steps (L2 flushed before each), prep %s" % (name, world, "observation-sharded-p2p" if obs_mode else loc.last_prep))
    for nm, us in tr:
        print("  %-30s %9.1f us" % (nm, us))
dist.barrier()
dist.destroy_process_group()
