"""Node-sharded fit / marginal across the GPUs of one box (one process per GPU, torch.distributed).

The reference is single-threaded; its `eval_grid!` loop over nodes (reference
src/joint_posterior.jl:180,186) carries no cross-node state until the normalisation, and `marginal`
(reference src/marginal_posterior.jl:117-123, src/interp.jl:448-457) only needs global sums, extrema
and -- for the 100-knot Grid -- per-knot (mass below, predecessor, successor).  So each rank owns a
contiguous block of merged grid nodes and the ranks exchange only:

    fit        all_gather(local max m_r, local sum s_r relative to m_r)             one collective
               tensor-core GLM path: its O(N) FP64 prep is sharded by OBSERVATION instead of being replicated --
               all_gather(slice sums + bounds), then ONE all_gather of the chosen coefficient rows (fit_sharded)
    marginals  all_gather(K x 4 moments/extrema)  ;  all_gather(K x 98 x 6 knot candidates)

Between the collectives the host does no arithmetic: the gathered buffers go straight back into the
library (jp_fit_normalise_gathered, jp_marginal_local_knots_gathered, jp_marginal_combine_gathered),
which combines them in rank order, so all ranks hold bit-identical results.  On CUDA the collectives
are NCCL over NVLink, on CPU (tests) gloo.  The per-rank "local phase" is an object with five methods;
`CudaLocal` calls the C ABI, tests substitute a numpy stand-in built on the `reference_*` functions
below (the same combine written with torch ops, also used to cross-check the CUDA combine).
"""
import ctypes as C

import numpy as np

from ._lib import GRID_KNOTS, check, lib

NK = GRID_KNOTS - 2   # interior knots


def shard_bounds(M, rank, world):
    """Contiguous node block [begin, end) of `rank`.  The grid is in mirror order (node 0 = origin, nodes 2j-1, 2j
    = a pair z, -z that the tensor-core path evaluates together), so interior cuts fall on odd indices: no pair
    straddles two ranks.  Sizes differ by at most two."""
    M, world = int(M), int(world)

    def cut(r):
        if r <= 0:
            return 0
        if r >= world:
            return M
        c = (r * M) // world
        return min(M, c | 1)
    return cut(rank), cut(rank + 1)


def _all_gather(t, group):
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    out = torch.empty((world,) + tuple(t.shape), dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(out.view(-1), t.contiguous().view(-1), group=group)
    return out


def row_slice(N, rank, world):
    """Rows [begin, end) that `rank` uploads; every rank's slot holds ceil(N / world) rows (the last may be short)."""
    n_loc = -(-int(N) // int(world))
    return min(N, rank * n_loc), min(N, (rank + 1) * n_loc), n_loc


def gather_rows(obs, device, group=None, gather=None, rank=None, world=None):
    """All N x ncols records on `device`, this rank having copied only its own row slice from the host.

    obs: the host array [N, ncols] (every rank can name all of it -- the generators are counter-based -- but only
    obs[begin:end] is read here).  One all_gather_into_tensor moves the slices GPU->GPU (NCCL over NVLink on CUDA,
    gloo in the CPU tests).  Returns a [world * n_loc, ncols] tensor whose first N rows are the records."""
    import torch
    import torch.distributed as dist
    if world is None:
        world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    if rank is None:
        rank = dist.get_rank(group) if world > 1 else 0
    N, ncols = obs.shape
    b, e, n_loc = row_slice(N, rank, world)
    mine = torch.empty((n_loc, ncols), dtype=torch.float64, device=device)
    host = torch.from_numpy(obs[b:e]) if not isinstance(obs, torch.Tensor) else obs[b:e]
    mine[:e - b].copy_(host, non_blocking=True)
    if e - b < n_loc:
        mine[e - b:].zero_()       # padding rows of a short last slice (never read: N bounds every kernel)
    if world == 1:
        return mine
    if gather is not None:
        return gather(mine, group).reshape(world * n_loc, ncols)
    full = torch.empty((world * n_loc, ncols), dtype=torch.float64, device=device)
    dist.all_gather_into_tensor(full.view(-1), mine.view(-1), group=group)
    return full


def reference_knot_values(gathered, vmin, vmax):
    """The 100 value knots [K, 100]: end knots are the global extrema, interior knots are the values the
    device computed (slot 5 of every rank's candidates; all ranks derive them from the same (min, max))."""
    import torch
    K = vmin.shape[0]
    x = torch.empty((K, GRID_KNOTS), dtype=torch.float64, device=vmin.device)
    x[:, 1:-1] = gathered[0, :, :, 5]
    x[:, 0] = vmin
    x[:, -1] = vmax
    return x


def reference_combine_knots(gathered, vmin, vmax):
    """gathered: [world, K, 98, 6] per-rank (S, pred, succ, succ_idx, succ_w, -) -> weight_nodes [K, 100].

    Reproduces itp[x] of the reference (interp.jl:28-31,453-455): left knot = LAST element <= x with
    cumulative weight S = sum of all weights <= x; right knot = first element of the next tie group
    (lowest global index among the smallest values > x) with cumulative weight S + its weight."""
    import torch
    tot = gathered[..., 0].sum(dim=0)    # fixed reduction order over the rank axis -> identical on every rank
    pred = gathered[..., 1].amax(dim=0)
    succ = gathered[..., 2].amin(dim=0)
    is_min = gathered[..., 2] == succ[None]
    idx = torch.where(is_min, gathered[..., 3], torch.full_like(gathered[..., 3], float("inf")))
    owner = idx.argmin(dim=0, keepdim=True)
    succ_w = torch.gather(gathered[..., 4], 0, owner)[0]
    x = gathered[0, :, :, 5]
    fx = (x - pred) / (succ - pred)
    inner = tot * (1.0 - fx) + (tot + succ_w) * fx
    K = inner.shape[0]
    wn = torch.zeros((K, GRID_KNOTS), dtype=torch.float64, device=inner.device)
    wn[:, 1:-1] = inner
    wn[:, -1] = 1.0
    return wn


def reference_fit_scale(g, rank):
    """Scale factor of rank `rank` from the gathered (m_r, s_r) pairs [world, 2]: density = e exp(m_rank - M) / S with
    M = max_r m_r, S = sum_r s_r exp(m_r - M) in rank order (what jp_scale_gathered_kernel computes)."""
    import torch
    M = g[:, 0].amax()
    S = torch.zeros((), dtype=torch.float64, device=g.device)
    for r in range(g.shape[0]):
        S = S + (g[r, 1] * torch.exp(g[r, 0] - M) if float(g[r, 1]) != 0.0 else 0.0)
    return torch.exp(g[rank, 0] - M) / S


def reference_combine(gm, gc):
    """(mu, sigma, value_nodes, weight_nodes) from the gathered moments [world, K, 4] and candidates [world, K, 98, 6]
    with torch ops (what jp_combine_gathered_kernel computes)."""
    import torch
    s = gm[:, :, :2].sum(dim=0)
    vmin = gm[:, :, 2].amin(dim=0)
    vmax = gm[:, :, 3].amax(dim=0)
    wn = reference_combine_knots(gc, vmin, vmax)
    vn = reference_knot_values(gc, vmin, vmax)
    mu = s[:, 0]
    sigma = torch.sqrt(s[:, 1] - mu * mu)
    return mu, sigma, vn, wn


def bulk_bytes_for(N, world):
    """Bytes of a communicator's bulk region that the coefficient rows of N observations may need on `world` ranks
    (12 rows of 4-byte coefficients per observation, slices of whole 128-observation tiles; include/jpcuda.h)."""
    tiles = -(-int(N) // 128)
    n_loc = -(-tiles // int(world)) * 128
    return int(world) * 12 * n_loc * 4


class Comm:
    """jp_comm: this rank's mailbox plus the mapped mailboxes of its peers (csrc/jp_comm.cu).  The 64-byte IPC handles are
    exchanged ONCE through `exchange(bytes) -> [bytes] * world` (default: torch.distributed.all_gather_object); every
    data-path exchange afterwards is a kernel of the library storing into peer memory over NVLink."""

    def __init__(self, ctx, rank, world, bulk_bytes=0, group=None, exchange=None):
        h = C.c_void_p()
        rc = lib().jp_comm_create(ctx.handle, C.c_int(rank), C.c_int(world), C.c_longlong(int(bulk_bytes)), C.byref(h))
        self.handle, self.ctx, self.rank, self.world, self.group = (h if rc == 0 else None), ctx, int(rank), int(world), group
        if world > 1:
            # a rank whose mailbox could not be created still takes part in the handle exchange (with an empty handle), so
            # that every rank of the group fails together instead of the others waiting in the collective
            buf = (C.c_ubyte * 64)()
            if rc == 0:
                rc = lib().jp_comm_ipc_handle(h, buf)
            if exchange is None:
                import torch.distributed as dist

                def exchange(mine):
                    out = [None] * world
                    dist.all_gather_object(out, mine, group=group)
                    return out
            handles = exchange(bytes(buf) if rc == 0 else b"")
            check(rc)
            blob = b"".join(handles)
            if len(blob) != 64 * world:
                self.destroy()
                raise ValueError("Comm: expected %d handles of 64 bytes (a peer rank could not create its mailbox)" % world)
            check(lib().jp_comm_connect_ipc(h, C.c_char_p(blob)))
        else:
            check(rc)

    @property
    def bulk_bytes(self):
        return int(lib().jp_comm_bulk_bytes(self.handle))

    def status(self):
        """Synchronise the stream and raise if a peer never reached a matching exchange."""
        check(lib().jp_comm_status(self.handle))

    def all_gather(self, t):
        """[world, n] from every rank's float64 device tensor of n elements (asynchronous on the context's stream)."""
        import torch
        t = t.contiguous().view(-1)
        out = torch.empty((self.world, t.numel()), dtype=torch.float64, device=t.device)
        check(lib().jp_comm_all_gather(self.handle, C.c_void_p(t.data_ptr()), C.c_int(t.numel()), C.c_void_p(out.data_ptr())))
        return out

    def destroy(self):
        if getattr(self, "handle", None):
            lib().jp_comm_destroy(self.handle)
            self.handle = None


def bulk_bytes_obs(nodes, world):
    """Bulk bytes of an observation-sharded fit: one (even, odd) partial sum per mirror pair of nodes and rank."""
    return int(world) * ((int(nodes) >> 1) + 1) * 16


def comm_for(ctx, N, group=None, min_bulk=0):
    """The context's communicator over `group`, created (collectively) on first use and re-created when a data set needs a
    larger bulk region.  None when torch.distributed is not initialised, the group has one rank, or JP_NO_P2P is set."""
    import os
    import torch.distributed as dist
    if os.environ.get("JP_NO_P2P") or not (dist.is_available() and dist.is_initialized()):
        return None
    world = dist.get_world_size(group)
    if world < 2 or world > 8:
        return None
    need = max(bulk_bytes_for(N, world), int(min_bulk))
    cache = ctx.__dict__.setdefault("_comms", {})
    key = id(group) if group is not None else 0
    c = cache.get(key)
    if c is None or c.bulk_bytes < need:
        if c is not None:
            c.destroy()
        c = Comm(ctx, dist.get_rank(group), world, need, group)
        cache[key] = c
    return c


class CudaLocal:
    """Local phases of one rank through the C ABI, on torch's current CUDA stream.  With a `comm` (jp_comm) the whole
    sharded fit and the global marginals are ONE library call each (exchanges over NVLink peer memory inside the library);
    without, the phase calls below are driven with torch.distributed collectives in between."""

    def __init__(self, jp, comm="auto", group=None):
        import torch
        self.jp = jp
        self.torch = torch
        self.dev = torch.device("cuda", jp.ctx.device)
        jp.ctx.use_stream(torch.cuda.current_stream(self.dev).cuda_stream)
        self.comm = comm_for(jp.ctx, jp.data.N, group) if isinstance(comm, str) else comm

    # ---- exchanges inside the library (jp_fit_p2p / jp_marginal_coords_p2p)
    def fit_p2p(self):
        check(lib().jp_fit_p2p(self.jp.handle, C.byref(self.jp._args), self.comm.handle))
        self.jp._theta = self.jp._density = None

    def marginals_p2p(self, coords):
        cs = np.ascontiguousarray(coords, dtype=np.int32)
        K = len(cs)
        mu, sg = np.zeros(K), np.zeros(K)
        vn, wn = np.zeros((K, GRID_KNOTS)), np.zeros((K, GRID_KNOTS))
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        args = (self.jp.handle, self.comm.handle, C.c_int(K), p(cs), p(mu), p(sg), p(vn), p(wn))
        st = lib().jp_marginal_coords_p2p(*args)
        if st == 6 and self.jp._args.path == 0:
            # the tensor-core bounds did not hold (every rank sees the same decision): refit on the FP64 path, once
            self.jp._args.path = 1
            self.fit_p2p()
            st = lib().jp_marginal_coords_p2p(*args)
        check(st)
        return mu, sg, vn, wn

    def _buf(self, *shape):
        return self.torch.empty(shape, dtype=self.torch.float64, device=self.dev)

    # ---- one-collective protocol (what fit_sharded / marginals_sharded drive)
    def fit_local_stats(self):
        out = self._buf(2)
        check(lib().jp_fit_local_stats(self.jp.handle, C.byref(self.jp._args), C.c_void_p(out.data_ptr())))
        return out

    # ---- observation-sharded prep of the tensor-core path (jp_fit_prep_*, include/jpcuda.h)
    def prep_worthwhile(self):
        """The sharded prep trades O(N (1 - 1/world)) replicated FP64 work (~0.3 ns per observation) for two more
        collectives (slice sums + bounds, then ONE all_gather of the chosen coefficient rows), the host's read of the bounds
        hidden under this rank's node prep: on whenever there are several ranks.  JP_SHARDED_PREP_MIN_MB (default 0) sets a
        record-size threshold below which the prep stays replicated (A/B aid)."""
        import os
        return self.jp.data.nbytes >= float(os.environ.get("JP_SHARDED_PREP_MIN_MB", "0")) * 1e6

    def fit_prep_local(self, rank, world):
        """Slice sums and bounds of this rank's observation slice [L], or None when the posterior is not on the
        tensor-core path (not a GLM, constrained coordinates, path forced to FP64)."""
        L = int(lib().jp_fit_prep_len(C.c_int(self.jp._args.d)))
        out = self._buf(L)
        st = lib().jp_fit_prep_local(self.jp.handle, C.byref(self.jp._args), C.c_int(rank), C.c_int(world),
                                     C.c_void_p(out.data_ptr()))
        if st == 6:            # JP_ERR_UNSUPPORTED
            return None
        check(st)
        return out

    def fit_prep_gathered(self, g, rank):
        """-> number of coefficient rows to exchange, or None when the series bounds are not met (use the FP64 kernel)."""
        g = g.contiguous()
        n_rows = C.c_int()
        st = lib().jp_fit_prep_gathered(self.jp.handle, C.byref(self.jp._args), C.c_void_p(g.data_ptr()), C.c_int(g.shape[0]),
                                        C.c_int(rank), C.byref(n_rows))
        if st == 6:
            return None
        check(st)
        return n_rows.value

    def fit_coef_slab(self, n_rows, world):
        """(mine [count], everybody's [world, count]) as float32 torch views of the library's buffers: this rank's first
        n_rows coefficient rows and the destination of the one all_gather that exchanges them."""
        pl, pa, cnt = C.c_void_p(), C.c_void_p(), C.c_longlong()
        check(lib().jp_fit_coef_slab(self.jp.handle, C.c_int(n_rows), C.byref(pl), C.byref(pa), C.byref(cnt)))
        mine = _device_view_f32(self.torch, pl.value, cnt.value, self.dev)
        return mine, _device_view_f32(self.torch, pa.value, world * cnt.value, self.dev).view(world, cnt.value)

    def fit_local_stats_prepared(self):
        out = self._buf(2)
        check(lib().jp_fit_local_stats_prepared(self.jp.handle, C.byref(self.jp._args), C.c_void_p(out.data_ptr())))
        return out

    def fit_normalise_gathered(self, g, rank):
        g = g.contiguous()
        check(lib().jp_fit_normalise_gathered(self.jp.handle, C.c_void_p(g.data_ptr()), C.c_int(g.shape[0]), C.c_int(rank)))

    def moments(self, coords):
        cs = np.ascontiguousarray(coords, dtype=np.int32)
        out = self._buf(len(cs), 4)
        check(lib().jp_marginal_local_moments(self.jp.handle, C.c_int(len(cs)), cs.ctypes.data_as(C.c_void_p), None,
                                              C.c_void_p(out.data_ptr())))
        return out

    def knots_gathered(self, coords, gm):
        """Knot candidates of the value columns of the preceding `moments` call, global extrema from the gathered moments."""
        gm = gm.contiguous()
        K = len(coords)
        out = self._buf(K, NK, 6)
        check(lib().jp_marginal_local_knots_gathered(self.jp.handle, C.c_int(K), None, None, C.c_void_p(gm.data_ptr()),
                                                     C.c_int(gm.shape[0]), C.c_void_p(out.data_ptr())))
        return out

    def combine_gathered(self, gm, gc):
        gm, gc = gm.contiguous(), gc.contiguous()
        world, K = gm.shape[0], gm.shape[1]
        mu, sg = np.zeros(K), np.zeros(K)
        vn, wn = np.zeros((K, GRID_KNOTS)), np.zeros((K, GRID_KNOTS))
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        check(lib().jp_marginal_combine_gathered(self.jp.handle, C.c_int(K), C.c_int(world), C.c_void_p(gm.data_ptr()),
                                                 C.c_void_p(gc.data_ptr()), p(mu), p(sg), p(vn), p(wn)))
        return mu, sg, vn, wn

    # ---- two-collective fit / explicit (min, max) knots: the phases of include/jpcuda.h one by one
    def fit_local_max(self):
        out = self._buf(1)
        check(lib().jp_fit_local(self.jp.handle, C.byref(self.jp._args), C.c_void_p(out.data_ptr())))
        return out

    def fit_local_sum(self, gmax):
        out = self._buf(1)
        check(lib().jp_fit_local_sum(self.jp.handle, C.c_void_p(gmax.data_ptr()), C.c_void_p(out.data_ptr())))
        return out

    def fit_normalise(self, gsum):
        check(lib().jp_fit_normalise(self.jp.handle, C.c_void_p(gsum.data_ptr())))

    def knots(self, coords, minmax):
        cs = np.ascontiguousarray(coords, dtype=np.int32)
        out = self._buf(len(cs), NK, 6)
        check(lib().jp_marginal_local_knots(self.jp.handle, C.c_int(len(cs)), cs.ctypes.data_as(C.c_void_p), None,
                                            C.c_void_p(minmax.data_ptr()), C.c_void_p(out.data_ptr())))
        return out


def _device_view_f32(torch, ptr, numel, dev):
    """A float32 torch tensor over `numel` elements of device memory owned by the library (no copy)."""
    class _Ext:
        pass
    e = _Ext()
    e.__cuda_array_interface__ = dict(shape=(int(numel),), typestr="<f4", data=(int(ptr), False), version=2)
    return torch.as_tensor(e, device=dev)


def _rank(group):
    import torch.distributed as dist
    return dist.get_rank(group)


def fit_sharded(local, group=None, gather=None, rank=None, world=None, gather_slab=None):
    """Normalise a node-sharded fit.  `gather(t, group)` defaults to torch.distributed all_gather; tests emulating
    several ranks inject their own.  Every combine runs in rank order on every rank: bit-identical scalars.

    When the local phase offers it (tensor-core GLM path), the O(N) prep is sharded by observation first: one tiny
    all_gather of (slice sums, bounds), then ONE all_gather of the chosen coefficient rows into the library's buffer
    (`gather_slab(mine, everybody, group)`, default: torch all_gather_into_tensor on the views)."""
    if getattr(local, "comm", None) is not None:        # exchanges inside the library: one call, no collective from here
        local.fit_p2p()
        local.last_prep = "sharded-p2p"
        return None
    gather = gather or _all_gather
    r = _rank(group) if rank is None else rank
    if world is None:
        import torch.distributed as dist
        world = dist.get_world_size(group)
    stats = None
    prep = getattr(local, "fit_prep_local", None)
    if prep is not None and world > 1 and local.prep_worthwhile():
        mine = prep(r, world)
        if mine is not None:
            n_rows = local.fit_prep_gathered(gather(mine, group), r)
            if n_rows is not None:
                slab, everybody = local.fit_coef_slab(n_rows, world)
                (gather_slab or _all_gather_slab)(slab, everybody, group)
                stats = local.fit_local_stats_prepared()
    local.last_prep = "sharded" if stats is not None else "replicated"      # which protocol this fit took (diagnostics)
    if stats is None:
        stats = local.fit_local_stats()
    g = gather(stats, group)      # [world, 2]
    local.fit_normalise_gathered(g, r)
    return g


def _all_gather_slab(mine, everybody, group):
    """mine: [count] (this rank's coefficient rows), everybody: [world, count]."""
    import torch.distributed as dist
    dist.all_gather_into_tensor(everybody.view(-1), mine, group=group)


def marginals_sharded(local, coords, group=None, gather=None):
    """Global (mu, sigma, value_nodes, weight_nodes) for K coordinate marginals of a node-sharded posterior:
    two all_gathers per batch, no global sort, no host arithmetic in between."""
    if getattr(local, "comm", None) is not None:
        return local.marginals_p2p(coords)
    gather = gather or _all_gather
    gm = gather(local.moments(coords), group)                   # [world, K, 4] = (sum w v, sum w v^2, min, max)
    gc = gather(local.knots_gathered(coords, gm), group)        # [world, K, 98, 6]
    return local.combine_gathered(gm, gc)


# ----------------------------------------------------------------------------- observation sharding
def mode_p2p(M, ddata, comm, x0=None):
    """mode(M, data) (reference src/joint_posterior.jl:164-168) of a GLM whose rows are sharded over the ranks of `comm`:
    jp_mode_p2p, every rank returns the same bits."""
    from .model import deduce_scale
    d = M.d
    x = np.zeros(d) if x0 is None else np.array(x0, dtype=np.float64)
    H = np.zeros((d, d), order="F")
    neg_min, evals = C.c_double(), C.c_int()
    code = np.ascontiguousarray(M.transform, dtype=np.int32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    check(lib().jp_mode_p2p(ddata.ctx.handle, ddata.handle, comm.handle, C.c_int(d), p(code), p(x), p(H), C.byref(neg_min),
                            C.byref(evals)))
    from .model import _warn_unless_converged
    _warn_unless_converged("mode_p2p")
    return x, deduce_scale(M, M.hessian_scale * np.array(H)), neg_min.value


class ObsShardedPosterior:
    """Result of an OBSERVATION-sharded fit: this rank uploaded and holds only its rows, but -- the remainder sums of the
    tensor-core GLM path being additive over observations -- it ends up with the complete joint posterior (all nodes,
    bit-identical on every rank), so marginals, quantiles and downloads are those of a single-GPU JointPosterior."""

    def __init__(self, post, comm, rows):
        self.post, self.comm, self.rows, self.prep = post, comm, rows, "observation-sharded-p2p"
        self.M, self.mu_hat, self.U = post.M, post.mu_hat, post.U

    def refit(self):
        check(lib().jp_fit_p2p_obs(self.post.handle, C.byref(self.post._args), self.comm.handle))
        self.post._theta = self.post._density = None
        self.post._theta_gen += 1
        return self

    density = property(lambda self: self.post.density)
    Theta = property(lambda self: self.post.Theta)

    def marginals(self, coords):
        from .marginals import marginals
        return marginals(self.post, [int(c) for c in coords])

    def marginal(self, k):
        return self.marginals([k])[0]

    def free(self):
        self.post.free()


def glm_obs_shardable(M, data):
    from .data import FAM_LOGISTIC, FAM_POISSON
    return data.family in (FAM_LOGISTIC, FAM_POISSON) and bool(np.all(M.transform == 0)) and M.d <= 32


def fit_obs_sharded(M, data, n=None, group=None, mode_result=None, ddata=None, comm=None):
    """fit(M, data[, n]) with the OBSERVATIONS sharded over the ranks of `group` (GLM families on the tensor-core path):
    this rank uploads rows [b, e) only; mode, series-length decision and the exchange of the per-pair partial sums happen
    inside the library over NVLink peer memory.  Raises JPError(status 6) when the a-priori bounds of the tensor-core path
    fail at the first blocking call (use fit_distributed(..., shard="nodes") then)."""
    import torch
    import torch.distributed as dist
    from .model import DeviceData, JointPosterior, colmajor, default
    ctx = M.ctx
    dev = torch.device("cuda", ctx.device)
    ctx.use_stream(torch.cuda.current_stream(dev).cuda_stream)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if n is None:
        n = default(M.build)
    N = data.records()[0].shape[0] if ddata is None else None
    if ddata is None:
        b, e, _ = row_slice(N, rank, world)
        ddata = DeviceData(ctx, data, rows=(b, e))
        rows = (b, e)
    else:
        rows = None
    if comm is None:
        comm = comm_for(ctx, 0, group, min_bulk=1 << 22)
    if comm is None:
        raise RuntimeError("fit_obs_sharded needs an initialised torch.distributed group of 2..8 ranks (and JP_NO_P2P unset)")
    mu_hat, U, neg_min = mode_p2p(M, ddata, comm) if mode_result is None else mode_result
    U = colmajor(U)
    grid = ctx.grid(M.build.rule.rule_id, U.shape[1], int(n))
    Mtot = int(lib().jp_grid_size(grid))
    if comm.bulk_bytes < bulk_bytes_obs(Mtot, world):
        comm = comm_for(ctx, 0, group, min_bulk=bulk_bytes_obs(Mtot, world))
    post = JointPosterior(M, ddata, grid, mu_hat, U, neg_min)
    return ObsShardedPosterior(post, comm, rows).refit()


# ----------------------------------------------------------------------------- the public multi-GPU call
class ShardedPosterior:
    """Result of `fit_distributed`: this rank's block of the joint posterior (reference struct JointPosterior,
    src/joint_posterior.jl:2-8, restricted to the nodes [begin, end)) plus what is needed to answer `marginal` globally.
    Every rank holds bit-identical global scalars, moments, knots and quantiles."""

    def __init__(self, post, local, group, bounds, prep):
        self.post, self.local, self.group, self.bounds, self.prep = post, local, group, bounds, prep
        self.M, self.mu_hat, self.U = post.M, post.mu_hat, post.U

    @property
    def density(self):
        """This rank's block of the normalised weights (global normalisation)."""
        return self.post.density

    @property
    def Theta(self):
        return self.post.Theta

    def marginals(self, coords):
        """Global marginal(jp, f) of coordinate selectors: two all_gathers per batch, no global sort."""
        from .marginals import Grid, marginal_result
        coords = [int(c) for c in coords]
        mu, sg, vn, wn = marginals_sharded(self.local, coords, self.group)
        return [marginal_result(None, j, mu[j], sg[j], Grid(wn[j].copy(), vn[j].copy())) for j in range(len(coords))]

    def marginal(self, k):
        return self.marginals([k])[0]

    def free(self):
        self.post.free()


def fit_distributed(M, data, n=None, group=None, path=None, mode_result=None, shard="nodes"):
    """fit(M, data[, n]) (reference src/joint_posterior.jl:177-188) with the grid nodes sharded over the ranks of `group`
    (one process per GPU, torch.distributed initialised by the caller): row-sharded upload + NVLink all_gather of the
    records, the mode on every rank (deterministic: identical everywhere), this rank's node block through stages 2-3, the
    global normalisation through one all_gather (plus the two of the observation-sharded tensor-core prep)."""
    import torch
    import torch.distributed as dist
    from . import _lib
    from .model import DeviceData, JointPosterior, colmajor, default, mode
    ctx = M.ctx
    if shard == "obs" or (shard == "auto" and not isinstance(data, DeviceData) and glm_obs_shardable(M, data)
                          and comm_for(ctx, 0, group, min_bulk=1 << 22) is not None):
        return fit_obs_sharded(M, data, n, group, mode_result)
    dev = torch.device("cuda", ctx.device)
    ctx.use_stream(torch.cuda.current_stream(dev).cuda_stream)      # collectives and kernels on one stream
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    ddata = data if isinstance(data, DeviceData) else ctx.upload_sharded(data, group)
    if n is None:
        n = default(M.build)
    mu_hat, U, neg_min = mode(M, ddata) if mode_result is None else mode_result
    U = colmajor(U)
    grid = ctx.grid(M.build.rule.rule_id, U.shape[1], int(n))
    Mtot = int(lib().jp_grid_size(grid))
    b, e = shard_bounds(Mtot, rank, world)
    post = JointPosterior(M, ddata, grid, mu_hat, U, neg_min, path=_lib.PATH_AUTO if path is None else path, node_range=(b, e))
    local = CudaLocal(post, group=group)
    fit_sharded(local, group)
    return ShardedPosterior(post, local, group, (b, e), getattr(local, "last_prep", "replicated"))
