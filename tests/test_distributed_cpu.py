"""World-size-2 `gloo` test of the node-sharded combine (jointposteriors.jl_b200/distributed.py):
the collectives + rank-ordered combine reproduce the single-process oracle, including the Grid tie rule across
shard boundaries.  The per-rank local phase is a numpy stand-in with the same contract as the CUDA
entry points (jp_fit_local_stats, jp_fit_normalise_gathered, jp_marginal_local_moments,
jp_marginal_local_knots_gathered, jp_marginal_combine_gathered); its combine is distributed.reference_*, the
torch restatement the GPU tests hold the CUDA combine kernels against."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class NumpyLocal:
    """Same contract as distributed.CudaLocal, on numpy arrays (CPU)."""

    def __init__(self, a, w, values, m0, D):
        import torch
        self.t = torch
        self.D = D
        self.a, self.w, self.values, self.m0 = a, w, values, m0
        self.density = None

    def _t(self, x):
        return self.t.tensor(np.asarray(x, dtype=np.float64))

    def fit_local_stats(self):
        m = self.a.max()
        self.e = self.w * np.exp(self.a - m)
        return self._t([m, self.e.sum()])

    def fit_normalise_gathered(self, g, rank):
        self.density = self.e * float(self.D.reference_fit_scale(g, rank))

    def moments(self, coords):
        out = []
        for k in coords:
            v = self.values[k]
            out.append([(self.density * v).sum(), (self.density * v * v).sum(), v.min(), v.max()])
        return self._t(out)

    def knots_gathered(self, coords, gm):
        mm = self.t.stack([gm[:, :, 2].amin(dim=0), gm[:, :, 3].amax(dim=0)], dim=1).numpy()
        out = np.zeros((len(coords), 98, 6))
        for j, k in enumerate(coords):
            v = self.values[k]
            for i in range(1, 99):
                x = float(np.longdouble(mm[j, 0]) + (np.longdouble(i) / np.longdouble(99)) * (np.longdouble(mm[j, 1]) - np.longdouble(mm[j, 0])))
                le = v <= x
                S = self.density[le].sum()
                pred = v[le].max() if le.any() else -np.inf
                gt = ~le
                if gt.any():
                    succ = v[gt].min()
                    jj = int(np.nonzero(v == succ)[0][0])
                    out[j, i - 1] = [S, pred, succ, self.m0 + jj, self.density[jj], x]
                else:
                    out[j, i - 1] = [S, pred, np.inf, np.inf, 0, x]
        return self._t(out)

    def combine_gathered(self, gm, gc):
        return tuple(x.numpy() for x in self.D.reference_combine(gm, gc))


class NumpyPrepLocal(NumpyLocal):
    """Adds the observation-sharded prep phases (same contract as CudaLocal.fit_prep_*): the 'prep' of this stand-in
    is a per-observation coefficient c_i = 2 i + 1 and the slice sum of i, the node values a_m then get the global sum
    and the dot of the gathered coefficients added -- so the fit only comes out right if every phase ran and every
    rank saw every slice."""

    def __init__(self, *args, n_obs=1000, world=1):
        super().__init__(*args)
        self.n_obs, self.world = n_obs, world
        self.n_loc = -(-n_obs // world)
        self.rows = self.t.zeros((2, self.n_loc), dtype=self.t.float32)               # this rank's slab: [rows][n_loc]
        self.all_rows = self.t.zeros((world, 2 * self.n_loc), dtype=self.t.float32)   # [world][rows * n_loc]
        self.a0 = self.a.copy()

    def prep_worthwhile(self):
        return True

    def fit_prep_local(self, rank, world):
        lo, hi = min(self.n_obs, rank * self.n_loc), min(self.n_obs, (rank + 1) * self.n_loc)
        i = np.arange(lo, hi, dtype=np.float64)
        self.rows[0, :hi - lo] = self.t.tensor(2 * i + 1, dtype=self.t.float32)
        self.rows[1, :hi - lo] = 1.0
        return self._t([i.sum(), float(hi - lo), 0.0])

    def fit_prep_gathered(self, g, rank):
        self.total = float(g[:, 0].sum())
        assert int(g[:, 1].sum()) == self.n_obs
        return 2

    def fit_coef_slab(self, n_rows, world):
        assert n_rows == 2 and world == self.world
        return self.rows.view(-1), self.all_rows

    def fit_local_stats_prepared(self):
        r = self.all_rows.numpy().astype(np.float64).reshape(self.world, 2, self.n_loc)
        shift = 1e-6 * self.total + 1e-9 * float(r[:, 0].sum()) + float(r[:, 1].sum()) - self.n_obs
        self.a = self.a0 + shift
        return self.fit_local_stats()


def _worker(rank, world, port, a, w, values, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import __graft_entry__ as entry
    entry.load_package()
    from jointposteriors_jl_b200 import distributed as D
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    M = len(a)
    b, e = D.shard_bounds(M, rank, world)
    loc = NumpyLocal(a[b:e], w[b:e], [v[b:e] for v in values], b, D)
    D.fit_sharded(loc)
    mu, sg, vn, wn = D.marginals_sharded(loc, list(range(len(values))))
    # the fit again through the observation-sharded prep protocol (extra all_gather + ONE gather of the rows): a constant
    # shift of every log-density leaves the normalised weights unchanged only if all ranks derived the SAME shift
    ploc = NumpyPrepLocal(a[b:e], w[b:e], [v[b:e] for v in values], b, D, n_obs=1003, world=world)
    D.fit_sharded(ploc)
    assert ploc.last_prep == "sharded"
    n = 1003
    want = 1e-6 * (n * (n - 1) / 2) + 1e-9 * float(n * n)          # sum i, sum (2 i + 1) = n^2; the row of ones cancels n
    assert abs((ploc.a - ploc.a0)[0] - want) < 1e-9 * max(1.0, abs(want)), ((ploc.a - ploc.a0)[0], want)
    assert np.allclose(ploc.density, loc.density, rtol=1e-12, atol=0)
    # row-sliced upload + all_gather (Context.upload_sharded): ragged N, every rank reassembles all records
    obs = np.arange(float(1003 * 4)).reshape(1003, 4)
    full = D.gather_rows(obs, torch.device("cpu"))
    rows_ok = bool(full.shape[0] >= 1003 and np.array_equal(full[:1003].numpy(), obs))
    q.put((rank, b, e, loc.density, mu, sg, vn, wn, rows_ok))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_combine_matches_oracle(O, jp, world):
    import torch.multiprocessing as mp
    rng = np.random.default_rng(5)
    M = 1001
    a = rng.standard_normal(M) * 3
    w = rng.random(M) + 0.05              # signed quadrature weights: 10 % negative
    w[rng.random(M) < 0.1] *= -1
    # value columns: continuous, heavily tied (coordinate-marginal like), and ties straddling the shard cut
    v0 = rng.standard_normal(M)
    v1 = np.round(rng.standard_normal(M) * 2) / 2
    v2 = np.sort(np.round(rng.standard_normal(M) * 4) / 4)
    values = [v0, v1, v2]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, a, w, values, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert all(r[8] for r in res)
    dens = np.concatenate([r[3] for r in res])
    e = w * np.exp(a - a.max())
    assert np.allclose(dens, e / e.sum(), rtol=1e-13, atol=0)
    for r in res[1:]:                      # every rank holds bit-identical results
        for x, y in zip(r[4:8], res[0][4:8]):
            assert np.array_equal(x, y, equal_nan=True)
    _, _, _, _, mu, sg, vn, wn, _ = res[0]
    for k, v in enumerate(values):
        m = O.marginal(v, dens)
        assert np.isclose(mu[k], m["mu"], rtol=1e-11, atol=1e-13)
        assert np.isclose(sg[k], m["sigma"], rtol=1e-10, equal_nan=True)
        assert np.allclose(vn[k], m["value_nodes"], rtol=1e-15, atol=1e-15)
        assert np.allclose(wn[k], m["weight_nodes"], rtol=1e-10, atol=1e-12), np.max(np.abs(wn[k] - m["weight_nodes"]))


def test_row_slices(jp):
    from jointposteriors_jl_b200.distributed import row_slice
    for N, W in [(10, 3), (1003, 8), (8, 8), (5, 8), (100000, 1)]:
        cuts = [row_slice(N, r, W) for r in range(W)]
        assert cuts[0][0] == 0 and cuts[-1][1] == N
        assert all(cuts[i][1] == cuts[i + 1][0] for i in range(W - 1))
        assert all(e - b <= n for b, e, n in cuts) and len({n for _, _, n in cuts}) == 1


def test_shard_bounds(jp):
    from jointposteriors_jl_b200.distributed import shard_bounds
    for M, W in [(10, 3), (1001, 8), (8, 8), (45201, 8)]:
        cuts = [shard_bounds(M, r, W) for r in range(W)]
        assert cuts[0][0] == 0 and cuts[-1][1] == M
        assert all(cuts[i][1] == cuts[i + 1][0] for i in range(W - 1))
        sizes = [e - b for b, e in cuts]
        assert max(sizes) - min(sizes) <= 2
        # interior cuts are odd: a mirror pair of nodes (2j-1, 2j) never straddles two ranks
        assert all(b % 2 == 1 or b in (0, M) for b, _ in cuts[1:])


def test_sharding_geometry_of_the_in_library_exchanges():
    """Host-side geometry of the jp_comm protocols (no GPU): row slices cover the observations exactly once, the bulk-region sizes
    bound what the kernels store (12 coefficient rows of whole 128-observation tiles per rank for node sharding, one (even, odd)
    pair per mirror pair of nodes and rank for observation sharding), and GLM models are recognised for observation sharding."""
    import __graft_entry__ as entry
    jp = entry.load_package()
    from jointposteriors_jl_b200 import distributed as D
    for N, world in ((100000, 8), (30001, 3), (7, 4), (128, 2)):
        cover = []
        for r in range(world):
            b, e, n_loc = D.row_slice(N, r, world)
            assert 0 <= b <= e <= N and e - b <= n_loc
            cover.extend(range(b, e))
        assert cover == list(range(N))
        tiles = -(-N // 128)
        n_loc_tiles = -(-tiles // world)
        assert D.bulk_bytes_for(N, world) == world * 12 * n_loc_tiles * 128 * 4
        assert D.bulk_bytes_for(N, world) >= 12 * 4 * N                      # room for every observation's 12 coefficients
    for M in (495, 115145, 189161):
        assert D.bulk_bytes_obs(M, 8) == 8 * ((M >> 1) + 1) * 16            # pairs (z, -z) plus the origin
    X = np.ones((10, 3))
    y = np.zeros(10)
    assert D.glm_obs_shardable(jp.Model((jp.RealVector(3),)), jp.LogisticData(X, y))
    assert D.glm_obs_shardable(jp.Model((jp.RealVector(3),)), jp.PoissonData(X, y))
    assert not D.glm_obs_shardable(jp.Model((jp.RealVector(2), jp.PositiveVector(1))), jp.NormalLinearData(X[:, :2], y))
    assert not D.glm_obs_shardable(jp.Model((jp.RealVector(2), jp.PositiveVector(1))), jp.LogisticData(X, y))
    # mirror pairs never straddle a node cut
    for M, world in ((115145, 8), (495, 3), (45201, 8)):
        cuts = [D.shard_bounds(M, r, world) for r in range(world)]
        assert cuts[0][0] == 0 and cuts[-1][1] == M and all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))
        assert all(c[0] % 2 == 1 or c[0] == 0 for c in cuts)
