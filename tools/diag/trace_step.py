"""Stage timeline of one device-resident step (fit + marginals of every coordinate) from the library's own CUDA-event marks
(jp_ctx_trace): python tools/diag/trace_step.py [cfg3|cfg4|cfg5] -- prints microseconds since the start of the fit."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry
import torch
jp = entry.load_package()
from jointposteriors_jl_b200 import workloads
from jointposteriors_jl_b200.model import Context, JointPosterior

name = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
wl = workloads.WORKLOADS[name]()
dev = torch.device("cuda", 0)
ctx = Context.get(0)
ctx.use_stream(torch.cuda.current_stream(dev).cuda_stream)
M = jp.Model(wl["params"], device=0)
dd = ctx.upload(wl["data"])
x, U, neg_min = jp.mode(M, dd)
grid = ctx.grid(0, U.shape[1], wl["level"])
post = JointPosterior(M, dd, grid, x, U, neg_min)
coords = list(range(wl["d"]))
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for it in range(6):
    flush.zero_()
    torch.cuda.synchronize()
    ctx.trace(True)
    post.evaluate()
    jp.marginals(post, coords)
    tr = ctx.trace_dump()
    ctx.trace(False)
print("%s: last of 6 steps (L2 flushed before each)" % name)
prev = {}
for nm, us in tr:
    print("  %-28s %9.1f us" % (nm, us))
