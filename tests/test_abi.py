"""The C-ABI library loads without a GPU, exports every symbol include/jpcuda.h declares, its host-side
helpers match the oracle bit for bit, and every compute entry point fails loudly when there is no device."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT


def header_symbols():
    src = open(os.path.join(ROOT, "include", "jpcuda.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(jp_[a-z0-9_]+)\s*\(", src)))


def test_exports_match_header(jp):
    from jointposteriors_jl_b200 import _lib
    declared = header_symbols()
    assert sorted(_lib.SYMBOLS) == declared
    L = jp.lib()
    for name in declared:
        assert hasattr(L, name), "libjpcuda.so does not export %s" % name
    out = subprocess.check_output(["nm", "-D", "--defined-only", os.path.join(ROOT, "jointposteriors.jl_b200", "libjpcuda.so")], text=True)
    exported = sorted(set(re.findall(r" T (jp_[a-z0-9_]+)$", out, flags=re.M)))
    assert [s for s in declared if s not in exported] == []


def test_built_for_sm100a_only():
    so = os.path.join(ROOT, "jointposteriors.jl_b200", "libjpcuda.so")
    out = subprocess.check_output(["cuobjdump", "-lelf", so], text=True)
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_no_device_is_an_error_not_a_fallback(jp):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(jp.JPError) as e:
        jp.Context(0)
    assert e.value.status == 4   # JP_ERR_NO_DEVICE
    assert "no CPU fallback" in str(e.value)


def test_product_does_not_touch_the_oracle():
    """Nothing under the package may import, link or load oracle/."""
    pkg = os.path.join(ROOT, "jointposteriors.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")) or f == "Makefile":
                txt = open(os.path.join(dirpath, f)).read()
                assert "jporacle" not in txt and "from oracle" not in txt and "import oracle" not in txt, f
    out = subprocess.check_output(["ldd", os.path.join(pkg, "libjpcuda.so")], text=True)
    assert "oracle" not in out


@pytest.mark.parametrize("d", [1, 2, 3, 10, 30, 64])
def test_host_linalg_bit_exact_with_oracle(jp, O, d):
    rng = np.random.default_rng(100 + d)
    A = rng.standard_normal((d, d))
    S = A @ A.T + d * np.eye(d)
    assert np.array_equal(np.triu(jp.chol(S)), np.triu(O.chol(S)))
    ok, U = jp.try_chol(S)
    assert ok and np.array_equal(np.triu(U), np.triu(O.chol(S)))
    assert np.array_equal(np.triu(jp.inv_upper(np.triu(U))), np.triu(O.inv_upper(np.triu(U))))
    assert np.array_equal(jp.inv_chol(S), O.inv_chol(S))
    assert np.array_equal(jp.deduce_scale_dynamic(S), O.deduce_scale_dynamic(S))


def test_try_chol_not_pd(jp):
    ok, _ = jp.try_chol(np.array([[1.0, 2.0], [2.0, 1.0]]))
    assert not ok


def test_reduce_dimensions_matches_oracle(jp, O):
    rng = np.random.default_rng(7)
    Q, _ = np.linalg.qr(rng.standard_normal((6, 6)))
    lam = np.array([0.0, 1e-13, 0.5, 2.0, 9.0, 11.0])
    H = (Q * lam) @ Q.T
    for mr in (0, 2):
        G, Go = jp.reduce_dimensions(H, mr), O.reduce_dimensions(H, mr)
        assert G.shape == Go.shape
        assert np.allclose(G @ G.T, Go @ Go.T, atol=1e-11)
    U = jp.deduce_scale_dynamic(H)
    assert U.shape == (6, 4)
    for g in (0.5, 0.9, 0.97, 0.999):       # LDR{g}
        G, Go = jp.reduce_dimensions_ldr(H, g), O.reduce_dimensions_ldr(H, g)
        assert G.shape == Go.shape and np.allclose(G @ G.T, Go @ Go.T, atol=1e-11)
    with pytest.raises(jp.JPError):
        jp.reduce_dimensions_ldr(H, 1.5)


def test_quantile_cdf_bit_exact_with_oracle(jp, O):
    rng = np.random.default_rng(3)
    vn = np.linspace(-2.0, 3.0, 100)
    wn = np.clip(np.linspace(0, 1, 100) + 0.03 * rng.standard_normal(100), -0.05, 1.05)
    wn[0], wn[-1] = 0.0, 1.0
    g = jp.Grid(wn, vn)
    for p in [0.0, 1e-9, 0.025, 0.25, 0.49999, 0.5, 0.75, 0.975, 1 - 1e-9, 1.0]:
        assert jp.quantile(g, p) == O.quantile(wn, vn, p)
    for x in [-3.0, -2.0, -1.99, 0.0, 0.123, 2.99, 3.0, 4.0]:
        assert jp.cdf(g, x) == O.cdf(wn, vn, x)


def test_rule_tables_identical_to_oracle(jp, O):
    L = jp.lib()
    for rule in (0, 1):
        npts, nodes, weights = O.rule_info(rule)
        lv, nm = C.c_int(), C.c_int()
        assert L.jp_rule_info(rule, C.byref(lv), C.byref(nm), None, None, None) == 0
        n2 = np.zeros(lv.value, dtype=np.int32)
        z2 = np.zeros(nm.value)
        w2 = np.zeros((lv.value, nm.value))
        assert L.jp_rule_info(rule, C.byref(lv), C.byref(nm), n2.ctypes.data_as(C.c_void_p),
                              z2.ctypes.data_as(C.c_void_p), w2.ctypes.data_as(C.c_void_p)) == 0
        assert np.array_equal(n2, npts) and np.array_equal(z2, nodes) and np.array_equal(w2, weights)


def test_model_declaration_styles(jp):
    """Tuple API (reference test/runtests.jl:5) and struct API (reference README.md:30-33)."""
    m1 = jp.Model((jp.ProbabilityVector(3),))

    class BinaryClassification(jp.parameter):
        p = jp.ProbabilityVector(3)

    m2 = jp.Model(BinaryClassification)
    assert m1.d == m2.d == 3 and list(m1.transform) == list(m2.transform) == [2, 2, 2]
    m3 = jp.Model((jp.RealVector(1), jp.PositiveVector(1), jp.RealVector(8)), jp.SmolyakRaw[jp.KronrodPatterson])
    assert m3.d == 10 and list(m3.transform) == [0, 1] + [0] * 8 and m3.build.rule.rule_id == 1 and m3.build.raw
    with pytest.raises(TypeError):
        jp.Model(3)
    # a Simplex(n) block takes n - 1 coordinates; its code word names the block's first coordinate and length
    m4 = jp.Model((jp.RealVector(2), jp.Simplex(4), jp.PositiveVector(1)))
    sx = 4 | (2 << 8) | (3 << 16)
    assert m4.d == 6 and list(m4.transform) == [0, 0, sx, sx, sx, 1]
    with pytest.raises(ValueError):
        jp.Simplex(1)


def test_coordinate_selector_detection(jp):
    from jointposteriors_jl_b200.params import probe_coordinate
    m = jp.Model((jp.RealVector(1), jp.PositiveVector(1), jp.RealVector(8)))
    assert probe_coordinate(lambda t: t.p3[4], m.blocks) == 6
    assert probe_coordinate(lambda t: t.p2[0], m.blocks) == 1
    assert probe_coordinate(lambda t: t.p3[4] * 2, m.blocks) is None
    assert probe_coordinate(lambda t: t.p3[1] - t.p3[0], m.blocks) is None
    # functions that are the identity on part of the range are NOT selectors (reference evaluates f at every node)
    for g in (lambda t: abs(t.p2[0]), lambda t: max(t.p2[0], 0), lambda t: np.clip(t.p2[0], 0, 2),
              lambda t: t.p2[0] if t.p2[0] > 0 else -t.p2[0], lambda t: float(t.p2[0]), lambda t: t.p2[0] + 0.0):
        assert probe_coordinate(g, m.blocks) is None
    m1 = jp.Model((jp.ProbabilityVector(3),))
    assert probe_coordinate(lambda p: p[0], m1.blocks) == 0
    # Simplex: the stored components are coordinates, the implied last one is a host closure
    m2 = jp.Model((jp.RealVector(1), jp.Simplex(3)))
    assert probe_coordinate(lambda t: t.p2[1], m2.blocks) == 2
    assert probe_coordinate(lambda t: t.p2[2], m2.blocks) is None


def test_data_validation(jp):
    with pytest.raises(ValueError):
        jp.LogisticData(np.zeros((0, 3)), np.zeros(0))
    with pytest.raises(ValueError):
        jp.BinaryClassificationData([], [], 9)
    with pytest.raises(ValueError):
        jp.LogisticData(np.zeros((4, 3)), np.zeros(5))
    obs, hyper = jp.BinaryClassificationData([0, 1], [2, 3], 9, βm=2, βp=2).records()
    assert obs.tolist() == [[0, 2, 9], [1, 3, 8]] and hyper.tolist() == [0, 1, 0, 1, 0, 0]


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` runs on the host cores alone and prints ONE JSON line with the contract's keys."""
    import json
    import sys
    out = subprocess.check_output([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                                   "--warmup", "0"], text=True, stderr=subprocess.DEVNULL, timeout=600)
    lines = [l for l in out.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "dtype",
              "data", "config", "impl", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["metric"] == "node_x_obs_log_density_evals_per_sec" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0


def test_bench_product_arm_needs_a_gpu():
    """Without a CUDA device the product arm refuses to run: there is no CPU fallback to time."""
    import sys
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "0"], text=True,
                       capture_output=True, timeout=600)
    assert r.returncode != 0 and "no CUDA device" in (r.stderr + r.stdout)


def test_fold_tables_are_what_the_generator_produces():
    """csrc/jp_fold_tables.h (economisation constants of the tensor-core series) is the output of tools/gen_fold.py:
    constants agree to 1e-9, and the tabulated error bounds really bound the error of the tabulated constants."""
    import sys
    txt = open(os.path.join(ROOT, "jointposteriors.jl_b200", "csrc", "jp_fold_tables.h")).read()

    def table(name, two_d):
        body = txt[txt.index(name):]
        body = body[body.index("{") + 1:body.index("};")]
        vals = [float(v) for v in re.findall(r"[-+]?\d+\.?\d*(?:[eE][-+]?\d+)?", body)]
        return np.array(vals).reshape(4, 5) if two_d else np.array(vals)

    ko, ke = table("k_fold_odd[TC_NFOLD][5]", True), table("k_fold_even[TC_NFOLD][5]", True)
    eo, ee = table("k_fold_eps_odd[TC_NFOLD]", False), table("k_fold_eps_even[TC_NFOLD]", False)
    go, ge = table("k_fold_grow_odd[TC_NFOLD]", False), table("k_fold_grow_even[TC_NFOLD]", False)
    x = np.linspace(0, 1, 200001)
    for j, NC in enumerate((4, 6, 8, 10)):
        odd = sum(ko[j][m] * x ** (3 + 2 * m) for m in range(NC // 2))
        even = sum(ke[j][m] * x ** (4 + 2 * m) for m in range(NC // 2))
        assert np.max(np.abs(x ** (NC + 3) - odd)) <= eo[j] and np.max(np.abs(x ** (NC + 4) - even)) <= ee[j]
        assert abs(go[j] - (1 + np.abs(ko[j]).sum())) < 1e-5 and abs(ge[j] - (1 + np.abs(ke[j]).sum())) < 1e-5
        assert np.all(ko[j][NC // 2:] == 0) and np.all(ke[j][NC // 2:] == 0)
    out = subprocess.check_output([sys.executable, os.path.join(ROOT, "tools", "gen_fold.py")], text=True, timeout=600)
    new = np.array([float(v) for v in re.findall(r"[-+]?\d+\.\d+(?:[eE][-+]?\d+)?", out[out.index("k_fold_odd"):])])
    old = np.array([float(v) for v in re.findall(r"[-+]?\d+\.\d+(?:[eE][-+]?\d+)?", txt[txt.index("k_fold_odd"):])])
    assert new.shape == old.shape and np.allclose(new, old, rtol=1e-6, atol=1e-9)


def test_rank_policies_in_deduce_scale(jp, O):
    """deduce_scale!(M, H, R) for R = Dynamic / Full / FixedRank{p} / LDR{g} (reference src/joint_posterior.jl:136-144)."""
    from jointposteriors_jl_b200 import model
    rng = np.random.default_rng(2)
    A = rng.standard_normal((5, 5))
    H = A @ A.T + 0.1 * np.eye(5)

    class FakeModel:
        pass
    m = FakeModel()
    for rank, p in ((jp.Dynamic, 5), (jp.Full, 5), (jp.FixedRank(2), 2), (jp.LDR(0.5), None), (jp.LDR(0.999), None)):
        m.rank = rank
        U = model.deduce_scale(m, H)
        assert U.shape[0] == 5 and (p is None or U.shape[1] == p)
        if U.shape[1] == 5:
            assert np.allclose(U @ U.T, np.linalg.inv(H), rtol=1e-9)
    m.rank = jp.LDR(0.5)
    assert model.deduce_scale(m, H).shape == O.reduce_dimensions_ldr(H, 0.5).shape
    with pytest.raises(ValueError):
        jp.LDR(1.0)


def test_smooth_cdf_host_functions(jp, O):
    """jp_smooth_cdf_eval / _pdf_eval / _quantile_eval (reference src/interp.jl:365-374) are host code: checked here against
    the oracle's restatement for a NestedPolyGLM given by its coefficients, no GPU involved."""
    from jointposteriors_jl_b200 import _lib
    V = np.stack([np.linspace(-2, 2, 50) ** k for k in range(10)], 1)
    phi = np.random.default_rng(0).standard_normal(9) * 0.4
    _, _, beta, theta = O.smooth_objective(V, np.linspace(0, 1, 50), phi)
    c = _lib.SmoothCDF()
    c.beta[:], c.theta[:] = list(beta), list(theta)
    c.mu, c.sigma = 0.3, 1.7
    fit = dict(beta=beta, theta=theta, mu=0.3, sigma=1.7)
    L = _lib.lib()
    for p in (1e-5, 0.001, 0.025, 0.1, 0.3, 0.5, 0.7, 0.9, 0.975, 0.999, 1 - 1e-10):
        q = L.jp_smooth_quantile_eval(C.byref(c), p)
        assert abs(q - O.smooth_quantile(fit, p)) < 1e-12 * abs(q)
        assert abs(L.jp_smooth_cdf_eval(C.byref(c), q) - p) < 1e-13
    for x in (-3.0, 0.0, 0.3, 1.0, 4.0):
        assert abs(L.jp_smooth_cdf_eval(C.byref(c), x) - O.smooth_cdf(fit, x)) < 1e-15
        assert abs(L.jp_smooth_pdf_eval(C.byref(c), x) - O.smooth_pdf(fit, x)) <= 1e-14 * O.smooth_pdf(fit, x)
    assert L.jp_smooth_quantile_eval(C.byref(c), 0.0) == -np.inf and L.jp_smooth_quantile_eval(C.byref(c), 1.0) == np.inf


def test_nested_poly_glm_host_mirror(jp, O):
    """The Python mirror of the reference's NestedPolyGLM methods (cdf / pdf / quantile dispatch of src/marginal_posterior.jl:
    140-148 onto src/interp.jl:365-374), vectorised, from a coefficient set -- host code only."""
    from jointposteriors_jl_b200 import _lib
    V = np.stack([np.linspace(-2, 2, 40) ** k for k in range(10)], 1)
    phi = np.array([-0.3, 0.2, 0.4, 0.1, 0.3, -0.2, 0.05, -2.0, 0.5])
    _, _, beta, theta = O.smooth_objective(V, np.linspace(0, 1, 40), phi)
    c = _lib.SmoothCDF()
    c.beta[:], c.theta[:], c.phi[:] = list(beta), list(theta), list(phi)
    c.mu, c.sigma = -1.0, 0.5
    itp = jp.NestedPolyGLM(c)
    fit = dict(beta=beta, theta=theta, mu=-1.0, sigma=0.5)
    ps = np.array([[0.025, 0.25], [0.5, 0.975]])
    qs = jp.quantile(itp, ps)
    assert qs.shape == ps.shape and np.all(np.diff(qs.ravel()) > 0)
    assert np.allclose(qs, [[O.smooth_quantile(fit, p) for p in row] for row in ps], rtol=1e-13)
    assert np.allclose(jp.cdf(itp, qs), ps, atol=1e-13)
    assert np.allclose(jp.pdf(itp, qs), O.smooth_pdf(fit, qs), rtol=1e-13)
    assert np.array_equal(itp.theta, theta) and np.array_equal(itp.β, beta)
    with pytest.raises(TypeError):
        jp.pdf(jp.Grid(np.linspace(0, 1, 100), np.linspace(0, 1, 100)), 0.5)      # a Grid has no pdf in the reference either
