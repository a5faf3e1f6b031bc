"""End-to-end step times at cfg3 without tracing (fresh upload + posterior per step), as bench.py's e2e leg does it."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry
import torch
jp = entry.load_package()
from jointposteriors_jl_b200 import workloads
from jointposteriors_jl_b200.model import Context, JointPosterior

wl = workloads.cfg3_logistic()
data = wl["data"]
obs, hyper = data.records()
dev = torch.device("cuda", 0)
ctx = Context.get(0)
ctx.use_stream(torch.cuda.current_stream(dev).cuda_stream)
M = jp.Model(wl["params"], device=0)
dd = ctx.upload(data)
x, U, neg_min = jp.mode(M, dd)
grid = ctx.grid(0, U.shape[1], wl["level"])
pin = torch.from_numpy(obs).pin_memory()
hd = type(data).__new__(type(data)); hd.__dict__.update(data.__dict__); hd._obs = pin.numpy()
coords = list(range(10))
keep = os.environ.get("KEEP_RESIDENT")
if keep:
    post = JointPosterior(M, dd, grid, x, U, neg_min)
    for _ in range(3):
        post.evaluate(); jp.marginals(post, coords)
ts = []
for it in range(20):
    t0 = time.perf_counter()
    d1 = ctx.upload(hd)
    pe = JointPosterior(M, d1, grid, x, U, neg_min)
    pe.evaluate()
    t1 = time.perf_counter()
    r = jp.marginals(pe, coords)
    t2 = time.perf_counter()
    q = pe.density
    t3 = time.perf_counter()
    pe.free(); d1.free()
    ts.append(((t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, (time.perf_counter() - t0) * 1e3))
a = np.array(ts[8:])
print("median over 12 steps: enqueue %.3f  marginals(blocking) %.3f  density %.3f  total %.3f ms" % tuple(np.median(a, axis=0)))
