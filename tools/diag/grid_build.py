"""Stage 1 timing: cold (first build in the process) and warm (second context, pool and kernels warm) Smolyak grid builds at the
BASELINE shapes; wall clock around the blocking jp_grid_get."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry
jp = entry.load_package()
from jointposteriors_jl_b200.model import Context
import ctypes as C
shapes = [("cfg1", 0, 3, 5), ("cfg2", 0, 10, 5), ("cfg3", 0, 10, 6), ("cfg4", 0, 20, 5), ("cfg5", 0, 30, 4), ("cfg1-KP7", 1, 3, 7)]
ctxs = [Context(0) for _ in range(4)]
for name, rule, d, L in shapes:
    ts = []
    for cx in ctxs:
        cx.sync()
        t0 = time.perf_counter()
        g = cx.grid(rule, d, L)
        cx.sync()
        ts.append((time.perf_counter() - t0) * 1e3)
    nm, npre = C.c_longlong(), C.c_longlong()
    jp.lib().jp_grid_build_stats(g, C.byref(nm), C.byref(npre))
    print("%-9s rule %d d=%2d L=%d: nodes %7d  pre-merge points %7d  multi-indices %6d | build ms: first %.2f, then %s" % (
        name, rule, d, L, jp.lib().jp_grid_size(g), npre.value, nm.value, ts[0], " ".join("%.2f" % t for t in ts[1:])))
