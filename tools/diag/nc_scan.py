"""Which series length the a-priori gate picks for a few synthetic GLM shapes (diagnostic)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as entry
jp = entry.load_package()
from conftest import synth_glm
for kind, N, d, level, xs in [("logistic", 50000, 6, 5, 1.0), ("logistic", 50000, 6, 6, 1.0), ("logistic", 20000, 6, 5, 1.0), ("logistic", 100000, 4, 7, 1.0),
                              ("poisson", 60000, 5, 5, 0.3), ("poisson", 60000, 5, 6, 0.3), ("poisson", 20000, 4, 6, 0.5), ("logistic", 200000, 3, 8, 1.0),
                              ("logistic", 30000, 8, 5, 1.0), ("poisson", 100000, 6, 5, 0.5)]:
    X, y = synth_glm(9, N, d, kind, xs)
    data = (jp.LogisticData if kind == "logistic" else jp.PoissonData)(X, y, 10.0)
    M = jp.Model((jp.RealVector(d),))
    try:
        post = jp.fit(M, data, level, path=jp.PATH_TC)
        print(kind, N, d, level, xs, "->", post.diagnostics)
    except Exception as e:
        print(kind, N, d, level, xs, "-> refused:", str(e)[-160:])
