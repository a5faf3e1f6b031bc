"""Stage timeline of one END-TO-END step at cfg3 (fresh upload + posterior per step), host-decided and device-decided series
length: python tools/diag/trace_e2e.py"""
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
for it in range(8):
    torch.cuda.synchronize()
    ctx.trace(True)
    t0 = time.perf_counter()
    d1 = ctx.upload(hd)
    pe = JointPosterior(M, d1, grid, x, U, neg_min)
    pe.evaluate()
    t1 = time.perf_counter()
    r = jp.marginals(pe, coords)
    t2 = time.perf_counter()
    q = pe.density
    t3 = time.perf_counter()
    tr = ctx.trace_dump()
    ctx.trace(False)
    pe.free(); d1.free()
print("host: enqueue %.3f ms, marginals (blocking) %.3f ms, density %.3f ms, total %.3f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, (t3 - t0) * 1e3))
for nm, us in tr:
    print("  %-28s %9.1f us" % (nm, us))
