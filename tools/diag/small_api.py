"""Where the public call goes at the reference's own sizes (cfg1 README Example 1, cfg2 eight schools): host phases of
fit(model, data, level) + marginals of every coordinate, and the mode finder's evaluation count.
python tools/diag/small_api.py"""
import os, sys, time, ctypes as C
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry
import torch
jp = entry.load_package()
from jointposteriors_jl_b200 import workloads, model as Mod
from jointposteriors_jl_b200._lib import lib
from jointposteriors_jl_b200.model import Context, JointPosterior

dev = torch.device("cuda", 0)
ctx = Context.get(0)
ctx.use_stream(torch.cuda.current_stream(dev).cuda_stream)
for name, mk in (("cfg1", workloads.cfg1_binary_classification), ("cfg2", workloads.cfg2_eight_schools)):
    wl = mk()
    data = wl["data"]
    M = jp.Model(wl["params"], device=0)
    coords = list(range(wl["d"]))
    for _ in range(5):
        p = jp.fit(M, data, wl["level"]); r = jp.marginals(p, coords); p.free()
    rows = []
    for it in range(7):
        torch.cuda.synchronize()
        t = [time.perf_counter()]
        dd = Mod._as_device_data(M, data); t.append(time.perf_counter())
        x, H, nm = Mod._native_mode(M, dd, np.zeros(M.d), False); t.append(time.perf_counter())
        g, itn, ok = C.c_double(), C.c_int(), C.c_int()
        lib().jp_mode_report(C.byref(g), C.byref(itn), C.byref(ok))
        U = Mod.deduce_scale(M, M.hessian_scale * H); t.append(time.perf_counter())
        grid = M.ctx.grid(M.build.rule.rule_id, U.shape[1], wl["level"]); t.append(time.perf_counter())
        pe = JointPosterior(M, dd, grid, x, Mod.colmajor(U), nm); t.append(time.perf_counter())
        pe.evaluate(); t.append(time.perf_counter())
        r = jp.marginals(pe, coords); t.append(time.perf_counter())
        pe.free(); t.append(time.perf_counter())
        rows.append(np.diff(t) * 1e3)
    med = np.median(np.array(rows), axis=0)
    names = ["upload", "mode", "deduce_scale", "grid(cached)", "posterior_create", "fit_enqueue", "marginals_blocking", "free"]
    print(name, "total %.3f ms;" % med.sum(), ", ".join("%s %.3f" % (n, v) for n, v in zip(names, med)), "; mode iterations", itn.value, "grad", g.value)
    # evaluation count and per-evaluation cost of the mode finder
    code = np.ascontiguousarray(M.transform, dtype=np.int32)
    xx = np.zeros(M.d); HH = np.zeros((M.d, M.d), order="F"); nmn = C.c_double(); ev = C.c_int()
    dd = Mod._as_device_data(M, data)
    t0 = time.perf_counter()
    lib().jp_mode(ctx.handle, dd.handle, C.c_int(M.d), code.ctypes.data_as(C.c_void_p), C.c_int(0), xx.ctypes.data_as(C.c_void_p), HH.ctypes.data_as(C.c_void_p), C.byref(nmn), C.byref(ev))
    t1 = time.perf_counter()
    print("   jp_mode alone %.3f ms, %d batched evaluations -> %.1f us each; x =" % ((t1 - t0) * 1e3, ev.value, (t1 - t0) * 1e6 / max(1, ev.value)), np.round(xx, 4))
    K = 1 + 2 * (2 * M.d + 2 * M.d * (M.d - 1))
    X = np.zeros((K, M.d)); out = np.zeros(K)
    t0 = time.perf_counter()
    for _ in range(50):
        lib().jp_log_density_points(ctx.handle, dd.handle, C.c_int(M.d), code.ctypes.data_as(C.c_void_p), C.c_longlong(K), X.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
    print("   jp_log_density_points of %d points: %.1f us per call" % (K, (time.perf_counter() - t0) * 1e6 / 50))
