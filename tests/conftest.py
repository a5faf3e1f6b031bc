"""Shared fixtures.  `-m "not gpu"` runs here (no GPU); `-m gpu` runs on a B200 and calls libjpcuda.so
through its C ABI (ctypes).  The oracle (oracle/) is the checker in both."""
import os
import sys

import numpy as np
import pytest

# one hardware queue per stream: the tests that emulate several ranks on one GPU run spinning exchange kernels of one rank
# beside the kernels of another, which must never share a queue (set before CUDA initialises)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
os.environ.setdefault("JP_COMM_TIMEOUT_S", "20")
# ... and no kernel may be loaded lazily while another rank's exchange kernel spins: loading synchronises the context
os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import __graft_entry__ as entry  # noqa: E402


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def O():
    """The CPU oracle (test infrastructure)."""
    from oracle import oracle
    oracle.build()
    oracle.set_precise(True)      # exact observation sums for the GLM families: the arbiter, see oracle/jp_oracle.cpp
    return oracle


@pytest.fixture(scope="session")
def jp():
    """The package (host mirror of the reference API over libjpcuda.so)."""
    so = os.path.join(entry.PKG_DIR, "libjpcuda.so")
    if not os.path.exists(so):
        entry.build()
    return entry.load_package()


@pytest.fixture(scope="session")
def gpu_ctx(jp):
    return jp.Context.get(0)


README_X = [0, 1, 2, 3, 4, 7, 8, 9]
README_FREQ = [10, 2, 2, 1, 2, 3, 2, 16]


def readme_records():
    """README Example 1 data (reference README.md:82-86): records (X, freq, NmX) and Beta(a,b) exponents."""
    X = np.array(README_X, dtype=np.float64)
    f = np.array(README_FREQ, dtype=np.float64)
    obs = np.ascontiguousarray(np.stack([X, f, 9 - X], axis=1))
    hyper = np.array([0.0, 1.0, 0.0, 1.0, 0.0, 0.0])
    return obs, hyper


def fd_hessian(f, x, h=1e-3):
    d = len(x)
    E = np.eye(d) * h
    H = np.zeros((d, d))
    for i in range(d):
        for j in range(i, d):
            H[i, j] = H[j, i] = (f(x + E[i] + E[j]) - f(x + E[i] - E[j]) - f(x - E[i] + E[j]) + f(x - E[i] - E[j])) / (4 * h * h)
    return H


def fd_gradient(f, x, h=1e-5):
    d = len(x)
    E = np.eye(d) * h
    return np.array([(f(x + E[i]) - f(x - E[i])) / (2 * h) for i in range(d)])


def cpu_mode(O, family, code, obs, hyper, x0):
    """Unconstrained posterior mode, Hessian of the negative log-density and its minimum, on the CPU:
    BFGS with central-difference gradients, then Newton polishing on finite-difference derivatives."""
    from scipy.optimize import minimize
    code = np.ascontiguousarray(code, dtype=np.int32)
    f = lambda x: -O.log_density_unc(family, code, x, obs, hyper)
    r = minimize(f, np.asarray(x0, dtype=np.float64), jac=lambda x: fd_gradient(f, x), method="BFGS",
                 options=dict(gtol=1e-9, maxiter=2000))
    x = r.x
    for _ in range(20):
        g, H = fd_gradient(f, x), fd_hessian(f, x)
        step = -np.linalg.solve(H, g)
        if np.max(np.abs(step)) < 1e-10 or f(x + step) > f(x):
            break
        x = x + step
    return x, fd_hessian(f, x), float(f(x))


def synth_glm(seed, N, d, kind, xscale=1.0):
    rng = np.random.Generator(np.random.Philox(key=seed))
    beta = rng.standard_normal(d) / np.sqrt(d)
    X = rng.standard_normal((N, d)) * xscale
    X[:, 0] = 1.0
    eta = X @ beta
    if kind == "logistic":
        y = (rng.random(N) < 1 / (1 + np.exp(-eta))).astype(np.float64)
    else:
        y = rng.poisson(np.exp(eta)).astype(np.float64)
    return X, y


def relerr(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))
