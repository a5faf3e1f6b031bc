"""Host-side breakdown of one end-to-end step at cfg3 (diagnostic, 1 GPU): where the time outside the kernels goes."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry
import torch
jp = entry.load_package()
from jointposteriors_jl_b200 import workloads, distributed as D
from jointposteriors_jl_b200.model import Context, JointPosterior, DeviceData

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
T = {}
def tic(name, f):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = f(); torch.cuda.synchronize()
    T.setdefault(name, []).append((time.perf_counter() - t0) * 1e3); return r
for it in range(8):
    d1 = tic("upload(jp_data_upload)", lambda: ctx.upload(hd))
    d2 = tic("upload(torch copy + adopt)", lambda: DeviceData(ctx, hd, device_obs=D.gather_rows(pin, dev, world=1, rank=0)))
    pe = tic("posterior_create", lambda: JointPosterior(M, d1, grid, x, U, neg_min))
    tic("evaluate", lambda: pe.evaluate())
    tic("evaluate again (state cached)", lambda: pe.evaluate())
    tic("marginals", lambda: jp.marginals(pe, coords))
    tic("density D2H", lambda: pe.density)
    tic("free", lambda: (pe.free(), d1.free(), d2.free()))
    def whole():
        d = ctx.upload(hd); p = JointPosterior(M, d, grid, x, U, neg_min); p.evaluate(); r = jp.marginals(p, coords); q = p.density; p.free(); d.free(); return r
    tic("whole step, no syncs inside", whole)
for k, v in T.items():
    print("%-34s median %.3f ms  min %.3f" % (k, float(np.median(v[2:])), min(v[2:])))
