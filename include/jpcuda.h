/* jpcuda.h -- C ABI of libjpcuda.so: the B200 (sm_100a) implementation of the posterior-integration
 * hot path of chriselrod/JointPosteriors.jl.
 *
 * The reference has no FFI seam of its own (pure Julia, multiple dispatch).  The seam this library
 * replaces is the internal call pair
 *     eval_grid!(M.Grid, ldc, mu_hat, U, index(...), MP)      reference src/joint_posterior.jl:180,186
 *     weights_values(jp, f) + Grid(wv)                        reference src/marginal_posterior.jl:98-123,
 *                                                             src/interp.jl:448-457
 * plus the small host-side scale-matrix helpers around it (src/joint_posterior.jl:15-144).
 * INTEGRATION.md shows the Julia `ccall` shim a maintainer would add; the Python mirror in
 * jointposteriors.jl_b200/ binds exactly these symbols through ctypes.
 *
 * Conventions
 *   - every entry point returns a jp_status (0 = ok); jp_last_error() gives the message
 *     (thread-local), mirroring the reference's Bool/throw conventions
 *     (src/joint_posterior.jl:26, src/interp.jl:444).
 *   - matrices named U/H/S are COLUMN-major (Julia layout); `obs` is ROW-major N x ncols.
 *   - pointers prefixed h_ are host memory, d_ are device memory on the context's GPU.
 *   - one jp_ctx = one GPU + one CUDA stream; calls on one ctx are serialised by the caller.
 *     Entry points that write h_ outputs block until the data is on the host; entry points that
 *     only touch d_ buffers are asynchronous on the ctx stream.
 *   - there is no CPU fallback: without a CUDA device jp_ctx_create fails with JP_ERR_NO_DEVICE.
 */
#ifndef JPCUDA_H
#define JPCUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum jp_status {
  JP_OK = 0,
  JP_ERR_BAD_ARG = 1,
  JP_ERR_NOT_PD = 2,      /* Cholesky pivot <= 0: the `false` of try_chol!, src/joint_posterior.jl:26 */
  JP_ERR_CUDA = 3,
  JP_ERR_NO_DEVICE = 4,
  JP_ERR_ALLOC = 5,
  JP_ERR_UNSUPPORTED = 6,
  JP_ERR_COMM = 7         /* a peer rank never reached the matching exchange (jp_comm_status) */
} jp_status;

/* quadrature rule families (SparseQuadratureGrids.GenzKeister / KronrodPatterson, reference test/runtests.jl:42) */
enum { JP_RULE_GENZ_KEISTER = 0, JP_RULE_KRONROD_PATTERSON = 1 };
/* constraint transforms (ConstrainedParameters RealVector / PositiveVector / ProbabilityVector,
 * reference src/JointPosteriors.jl:22-26, README.md:32,247-248) -- one code per unconstrained coordinate */
enum { JP_T_REAL = 0, JP_T_POSITIVE = 1, JP_T_PROBABILITY = 2, JP_T_NONCENTRED = 3, JP_T_SIMPLEX = 4, JP_T_COVMAT = 5 };
/* JP_T_NONCENTRED couples coordinate k to two EARLIER constrained coordinates: theta_k = theta_loc + theta_scale * x_k,
 * log|J| += log(theta_scale); loc and scale travel in the code word (hierarchical models whose centred
 * parameterisation has no joint mode, e.g. eight schools).  Transforms are applied in coordinate order. */
#define JP_T_NONCENTRED_CODE(loc, scale) (JP_T_NONCENTRED | ((loc) << 8) | ((scale) << 16))
/* JP_T_SIMPLEX (ConstrainedParameters Simplex, reference src/JointPosteriors.jl:26): the n - 1 coordinates
 * [first, first + len) are the first n - 1 components of a point of the n-simplex, the last one is implied:
 *   theta_k = e^{x_k} / (1 + sum_j e^{x_j}),  theta_n = 1 / (1 + sum_j e^{x_j}),  log|J| = sum_{k=1..n} log theta_k
 * (additive log-ratio map; det(diag(theta) - theta theta') = prod of all n components).  Every coordinate of the block
 * carries the same code word. */
#define JP_T_SIMPLEX_CODE(first, len) (JP_T_SIMPLEX | ((first) << 8) | ((len) << 16))
/* JP_T_COVMAT (ConstrainedParameters CovarianceMatrix, reference src/JointPosteriors.jl:22): the p (p + 1) / 2 coordinates
 * [first, first + len) hold the lower triangle of a p x p matrix row by row, (0,0), (1,0), (1,1), (2,0), ..  Unconstrained:
 * the log-Cholesky factor (L_ii = exp(x_ii), L_ij = x_ij for i > j); constrained: the lower triangle of Sigma = L L' in the
 * same order; log|J| = p log 2 + sum_{i=0}^{p-1} (p - i + 1) x_ii.  Every coordinate of the block carries the same code word. */
#define JP_T_COVMAT_CODE(first, len) (JP_T_COVMAT | ((first) << 8) | ((len) << 16))
#define JP_T_KIND(code) ((code) & 0xFF)
#define JP_T_LOC(code) (((code) >> 8) & 0xFF)
#define JP_T_SCALE(code) (((code) >> 16) & 0xFF)
/* likelihood families registered in the library (device-function plugins, csrc/jp_family.cuh) */
enum {
  JP_FAM_BINOMIAL_MIXTURE = 0, /* README Example 1, reference README.md:62-72 */
  JP_FAM_LOGISTIC = 1,         /* logistic regression, N(0, s^2) prior */
  JP_FAM_POISSON = 2,          /* Poisson regression (log link), N(0, s^2) prior */
  JP_FAM_HIER_NORMAL = 3,      /* hierarchical normal ("eight schools") */
  JP_FAM_NORMAL_LINEAR = 4,    /* README Example 2 "HiWorld", reference README.md:245-258 */
  JP_FAM_MULTINOMIAL = 5,      /* category counts with a symmetric Dirichlet prior on a Simplex block */
  JP_FAM_MVN_COV = 6,          /* zero-mean multivariate normal, inverse-Wishart prior, on a CovarianceMatrix block */
  JP_FAM_ANOVA2 = 7            /* balanced two-factor random-effects ANOVA, README Example 3 (reference README.md:416-470) */
};
/* log-density evaluation path of jp_fit */
enum {
  JP_PATH_AUTO = 0,
  JP_PATH_FP64 = 1,   /* generic plugin kernel, FP64 throughout: every family */
  JP_PATH_TC = 2      /* GLM families only: tcgen05 3xTF32 X*dTheta' contraction + centred epilogue */
};

typedef struct jp_ctx jp_ctx;
typedef struct jp_grid jp_grid;
typedef struct jp_data jp_data;
typedef struct jp_posterior jp_posterior;
typedef struct jp_comm jp_comm;

const char* jp_last_error(void);
int jp_version(void);

/* ---------------------------------------------------------------------------------------------
 * Host-side scale matrix helpers (no GPU needed).  d x d column-major.
 * --------------------------------------------------------------------------------------------- */
/* chol!      reference src/joint_posterior.jl:30-43  */
int jp_chol(double* h_U, const double* h_S, int d);
/* try_chol!  reference src/joint_posterior.jl:15-29; JP_ERR_NOT_PD when a pivot is not positive */
int jp_try_chol(double* h_U, const double* h_S, int d);
/* inv!       reference src/joint_posterior.jl:56-68 (in place, upper triangular) */
int jp_inv_upper(double* h_U, int d);
/* inv_chol!  reference src/joint_posterior.jl:72-76 : U = chol(H)^-1 (lower triangle zeroed) */
int jp_inv_chol(double* h_U, const double* h_H, int d);
/* reduce_dimensions!  reference src/joint_posterior.jl:98-110 (max_rank = 0) and :120-134
 * (FixedRank{p}: max_rank = p).  h_out is d x d storage; *rank receives the kept column count. */
int jp_reduce_dimensions(const double* h_H, int d, int max_rank, double* h_out, int* rank);
/* reduce_dimensions!(M, H, LDR{g})  reference src/joint_posterior.jl:78-95,111-119: the leading admissible
 * eigen directions (largest 1/lambda first) that carry the fraction g in (0,1) of the total variance.  The
 * reference's `count` uses `total_energy` uninitialised (:84); zero-initialised here. */
int jp_reduce_dimensions_ldr(const double* h_H, int d, double g, double* h_out, int* rank);
/* deduce_scale!(..., Dynamic)  reference src/joint_posterior.jl:136-138 : Cholesky if H is positive
 * definite, else the eigen fallback.  *rank receives p; h_U is d x p column-major in d x d storage. */
int jp_deduce_scale_dynamic(const double* h_H, int d, double* h_U, int* rank);

/* ---------------------------------------------------------------------------------------------
 * Context
 * --------------------------------------------------------------------------------------------- */
int jp_ctx_create(int device, jp_ctx** out);
int jp_ctx_destroy(jp_ctx* ctx);
/* run all subsequent work of this ctx on an externally owned cudaStream_t (e.g. torch's current
 * stream, so that torch.distributed collectives order with the library's kernels). */
int jp_ctx_set_stream(jp_ctx* ctx, void* cuda_stream);
int jp_ctx_sync(jp_ctx* ctx);
/* number of kernels this ctx has launched since creation (for bench.py's gpu_launches) */
long long jp_ctx_launch_count(const jp_ctx* ctx);
/* Stage tracing: jp_ctx_trace(ctx, 1) clears the trace and makes the entry points record CUDA events at their phase
 * boundaries (on the stream the phase runs on); jp_ctx_trace_dump synchronises and writes one "name<TAB>microseconds since
 * the first mark" line per event into buf.  jp_ctx_trace(ctx, 0) switches it off.  Diagnostic; not thread-safe. */
int jp_ctx_trace(jp_ctx* ctx, int on);
int jp_ctx_trace_dump(jp_ctx* ctx, char* buf, int len);
/* device time (CUDA events on the ctx stream) of the most recent launch of the dominant kernel of the path:
 * the node x observation log-density kernel of jp_fit (FP64 plugin kernel or tcgen05 GLM kernel).  Blocking. */
int jp_ctx_last_kernel_ms(jp_ctx* ctx, float* ms);

/* ---------------------------------------------------------------------------------------------
 * Stage 1 -- Smolyak sparse grid, built on the GPU and cached per ctx by (rule, d_eff, level),
 * the analogue of the reference's grid cache key index(U, R, data, n), src/joint_posterior.jl:157-162.
 * Nodes are integer keys (one 1-D master-node index per dimension); merged nodes are in ascending
 * lexicographic key order.
 * --------------------------------------------------------------------------------------------- */
int jp_grid_get(jp_ctx* ctx, int rule, int d_eff, int level, jp_grid** out);
long long jp_grid_size(const jp_grid* g);
int jp_grid_dim(const jp_grid* g);
/* pre-merge statistics of the build: number of multi-indices with non-zero combination coefficient
 * and number of tensor-product points before the duplicate merge */
int jp_grid_build_stats(const jp_grid* g, long long* n_multi, long long* n_premerge);
/* h_idx: M x d_eff row-major uint8 keys; h_w: M weights.  Either may be NULL. */
int jp_grid_download(const jp_grid* g, uint8_t* h_idx, double* h_w);
/* 1-D rule tables (master z-nodes, per-level weights BY MASTER INDEX, 0 for nodes a level does not use) as compiled
 * into the library.  Genz-Keister: 1, 3, 9, 19, 35 points (nested) and the published 37-, 41-, 43-point members, which
 * extend the 19-point rule; Kronrod-Patterson: 1 .. 63 points (nested).  (The rule family lives in the absent
 * SparseQuadratureGrids package, reference test/runtests.jl:42.) */
int jp_rule_info(int rule, int* levels, int* nmax, int* h_npts, double* h_nodes, double* h_weights);
/* master indices of the nodes of 1-D level `level` (1-based) in generation order (h_index may be NULL); returns their
 * number, or -1 for an unknown rule / level */
int jp_rule_level_nodes(int rule, int level, int* h_index);
/* highest 1-D level the build of this grid used: min(level, levels in the rule table).  A grid asked for a level past the
 * table end is the combination formula over the capped index set, and this is how a caller finds out. */
int jp_grid_level_cap(const jp_grid* g);

/* ---------------------------------------------------------------------------------------------
 * Observations
 * --------------------------------------------------------------------------------------------- */
/* h_obs: N x ncols row-major doubles, family-specific columns:
 *   BINOMIAL_MIXTURE (X, freq, NmX)            hyper = (a_m-1, b_m-1, a_p-1, b_p-1, a_tau-1, b_tau-1)
 *   LOGISTIC/POISSON (x_0..x_{d-1}, y)         hyper = (prior sd)
 *   HIER_NORMAL      (y_j, s_j)                hyper = (half-Cauchy scale of tau)
 *   NORMAL_LINEAR    (x_0..x_{p-1}, y)         hyper = (sd of beta prior, sd of sigma prior)
 *   MULTINOMIAL      (count_k), one row per category, N = d + 1      hyper = (alpha - 1 of the Dirichlet prior) */
int jp_data_upload(jp_ctx* ctx, int family, long long N, int ncols, const double* h_obs, const double* h_hyper,
                   int n_hyper, jp_data** out);
/* Same, for records that are ALREADY in the memory of ctx's GPU (e.g. row slices uploaded by each rank of a box and
 * all-gathered over NVLink, SURVEY 8(e): "X broadcast GPU->GPU once per dataset").  The buffer is borrowed: it must
 * stay valid until jp_data_free, work on it must be ordered on ctx's stream, and the library never frees it. */
int jp_data_adopt_device(jp_ctx* ctx, int family, long long N, int ncols, const double* d_obs, const double* h_hyper,
                         int n_hyper, jp_data** out);
int jp_data_free(jp_data* data);

/* GLM score and observed information at beta on the GPU (mode finding, upstream of the five
 * stages: reference src/joint_posterior.jl:164-168).  h_g: d; h_Hneg: d x d column-major =
 * -Hessian of the log posterior; *h_logpost: log posterior at beta. */
int jp_glm_grad_hess(jp_ctx* ctx, const jp_data* data, int d, const double* h_beta, double* h_g, double* h_Hneg,
                     double* h_logpost);

/* mode(M, data)  reference src/joint_posterior.jl:164-168 (optBFGS! + ForwardDiff Hessian there): the unconstrained
 * posterior mode by Newton iterations driven from native host code, every function / derivative evaluation one
 * batched GPU call.  glm != 0 (LOGISTIC / POISSON with unconstrained coefficients): analytic score and information
 * (jp_glm_grad_hess); glm == 0: saddle-free Newton on Richardson finite differences of jp_log_density_points.
 * Small models (d <= 16, records <= 32 KB: the reference's own examples) run the whole glm == 0 search inside ONE launch of
 * a thread-block cluster (csrc/jp_mode_dev.cu) and fall back to the host-driven iteration if that does not converge;
 * environment: JP_MODE_HOST=1 forces the host-driven iteration, JP_MODE_TRACE=1 prints the kernel's cycle account.
 * h_x: in = start, out = mode; h_H: d x d Hessian of the NEGATIVE log-density at the mode (what `mode` doubles and
 * hands to deduce_scale!, :167); *neg_min: the minimised objective (:166); *evals: GPU evaluations used. */
int jp_mode(jp_ctx* ctx, const jp_data* data, int d, const int* h_transform, int glm, double* h_x, double* h_H,
            double* neg_min, int* evals);
/* How the last jp_mode / jp_mode_p2p call of the calling thread ended: infinity norm of the gradient of the negative
 * log-density at the returned point, iterations used, converged = 1 when the search stopped on its own criterion, 0 at the
 * iteration cap or after a failed line search (the point is still returned: a grid centred beside the mode is a worse
 * quadrature, not an error -- the host decides whether to warn). */
int jp_mode_report(double* grad_inf_norm, int* iterations, int* converged);

/* Unconstrained log-density  log_density(transform(x), data) + log|J(x)|  at K arbitrary points: the
 * objective `mode` minimises (reference src/joint_posterior.jl:164-168, sign flipped, no neg_min),
 * evaluated by the same family plugins as the grid path.  h_x: K x d row-major; h_ld: K.  Blocking. */
int jp_log_density_points(jp_ctx* ctx, const jp_data* data, int d, const int* h_transform, long long K,
                          const double* h_x, double* h_ld);

/* ---------------------------------------------------------------------------------------------
 * Stages 2-4 -- fit: theta = transform(mu_hat + U z), log-density per node, normalised weights.
 * Replaces eval_grid!, reference src/joint_posterior.jl:180,186.
 * --------------------------------------------------------------------------------------------- */
typedef struct jp_fit_args {
  int d;                        /* number of unconstrained coordinates */
  int p;                        /* columns of U (= grid dimension d_eff), p <= d */
  const int* h_transform;       /* d transform codes */
  const double* h_mu_hat;       /* d: unconstrained mode, `M.diff_buffer.state.x` (:167) */
  const double* h_U;            /* d x p column-major scale matrix from deduce_scale! */
  double neg_min;               /* minimised objective, added to every log-density (:149,:153) */
  int path;                     /* JP_PATH_* */
  long long node_begin;         /* node shard [node_begin, node_end) of the merged grid owned by */
  long long node_end;           /*   this ctx; (0, -1) = all nodes */
  int raw;                      /* 1: RawBuild (src/joint_posterior.jl:183-188): the result keeps the UNCONSTRAINED node
                                 *    coordinates x = mu_hat + U z (the reference's grid.cache, jp_get_cache); the constrained
                                 *    parameters are constructed on the device whenever a marginal or jp_get_theta needs them
                                 *    (update!(Theta), src/marginal_posterior.jl:86-90,106-115).  0: CacheBuild (Theta stored) */
} jp_fit_args;

/* allocate the device-resident result (Theta SoA [d][M_local], density, work buffers) */
int jp_posterior_create(jp_ctx* ctx, const jp_grid* g, const jp_data* data, const jp_fit_args* args,
                        jp_posterior** out);
int jp_posterior_free(jp_posterior* post);
long long jp_posterior_size(const jp_posterior* post); /* local node count */

/* single-GPU fit: all of stages 2-4, asynchronous on the ctx stream. */
int jp_fit(jp_posterior* post, const jp_fit_args* args);

/* multi-GPU fit in phases; the host performs the (tiny) collectives on the d_ buffers between them:
 *   jp_fit_local(post, args, d_local_max)        stages 2-3 on the shard; writes max_m a_m (1 double)
 *   jp_fit_local_sum(post, d_global_max, d_sum)  e_m = w_m exp(a_m - gmax); writes sum_m e_m (1 double)
 *   jp_fit_normalise(post, d_global_sum)         density_m = e_m / gsum                               */
int jp_fit_local(jp_posterior* post, const jp_fit_args* args, double* d_local_max);
int jp_fit_local_sum(jp_posterior* post, const double* d_global_max, double* d_local_sum);
int jp_fit_normalise(jp_posterior* post, const double* d_global_sum);

/* The same normalisation with ONE collective per fit (what the sharded path uses):
 *   jp_fit_local_stats(post, args, d_stats)      stages 2-3 on the shard; d_stats[0] = m_r = max_m a_m,
 *                                                e_m = w_m exp(a_m - m_r), d_stats[1] = s_r = sum_m e_m
 *   -- all_gather of the (m_r, s_r) pairs into d_gathered[world][2] --
 *   jp_fit_normalise_gathered(post, d_gathered, world, rank)
 *                                                M = max_r m_r, S = sum_r s_r exp(m_r - M) (rank order, identical on
 *                                                every rank), density_m = e_m exp(m_rank - M) / S                      */
int jp_fit_local_stats(jp_posterior* post, const jp_fit_args* args, double* d_stats);
int jp_fit_normalise_gathered(jp_posterior* post, const double* d_gathered, int world, int rank);

/* Node-sharded fit on the tensor-core path with its O(N) FP64 prep (X'WX sums, per-observation series coefficients)
 * sharded by OBSERVATION: rank r prepares observation slice r only, so the work replicated on every rank stays
 * O(N / world).  Sequence per fit (device buffers; collectives by the caller, on ctx's stream):
 *   jp_fit_prep_local(post, args, rank, world, d_out[L])          L = jp_fit_prep_len(d); asynchronous
 *   all_gather -> g[world][L]
 *   jp_fit_prep_gathered(post, args, g, world, rank, &n_rows)     sums and bounds combined in rank order ON THE DEVICE, this
 *                                                                 rank's node operand queued, then the host reads the 22
 *                                                                 combined bounds (the device is busy meanwhile), decides
 *                                                                 the series length and folds this rank's slice
 *   jp_fit_coef_slab(post, n_rows, &d_local, &d_all, &count)      ONE all_gather: every rank contributes `count` floats at
 *                                                                 d_local (its first n_rows coefficient rows) and receives
 *                                                                 float d_all[world][count]
 *   jp_fit_local_stats_prepared(post, args, d_out[2])             = jp_fit_local_stats without the prep
 * JP_ERR_UNSUPPORTED from the first two calls (not a GLM / bounds not met) means: use jp_fit_local_stats. */
int jp_fit_prep_len(int d);
int jp_fit_prep_local(jp_posterior* post, const jp_fit_args* args, int rank, int world, double* d_out);
int jp_fit_prep_gathered(jp_posterior* post, const jp_fit_args* args, const double* d_gathered, int world, int rank,
                         int* n_rows);
int jp_fit_coef_slab(jp_posterior* post, int n_rows, void** d_local, void** d_all, long long* count);
int jp_fit_local_stats_prepared(jp_posterior* post, const jp_fit_args* args, double* d_stats);

/* ---- The node-sharded path with its exchanges INSIDE the library, over NVLink peer memory (csrc/jp_comm.cu).
 * SURVEY 8b sketched jp_ctx_create(n_gpus); the sharding keeps one process (and one jp_ctx) per GPU, and a jp_comm joins
 * the ranks of one box.  Every rank owns a mailbox in its GPU's memory; peers map it through CUDA IPC and the library's
 * kernels store their payloads and a sequence flag straight into it (no NCCL call, no host synchronisation between the
 * phases).  The host moves ONE 64-byte handle per rank, once, by whatever it has (MPI, sockets, a file, torch.distributed):
 *   jp_comm_create(ctx, rank, world, bulk_bytes, &comm)   bulk_bytes: room for the coefficient rows of the tensor-core
 *                                                         path, 4 bytes x 12 rows x observations rounded up to 128 x world
 *                                                         (0: GLM fits then keep their O(N) prep replicated)
 *   jp_comm_ipc_handle(comm, h[64])  --  all ranks exchange their handles  --  jp_comm_connect_ipc(comm, handles[world][64])
 *   (one process driving several contexts, e.g. tests: jp_comm_connect_local(comm, comms[world]))
 * Every rank must then issue the same sequence of jp_*_p2p calls.  jp_comm_status synchronises the stream and reports a
 * peer that never arrived (JP_ERR_COMM after JP_COMM_TIMEOUT_S seconds, default 60). */
int jp_comm_create(jp_ctx* ctx, int rank, int world, long long bulk_bytes, jp_comm** out);
int jp_comm_ipc_handle(jp_comm* comm, void* handle64);
int jp_comm_connect_ipc(jp_comm* comm, const void* handles);
int jp_comm_connect_local(jp_comm* comm, jp_comm* const* comms);
long long jp_comm_bulk_bytes(const jp_comm* comm);
int jp_comm_status(jp_comm* comm);
int jp_comm_destroy(jp_comm* comm);
/* all_gather of n doubles per rank (device buffers, asynchronous): d_out[world][n] */
int jp_comm_all_gather(jp_comm* comm, const double* d_src, int n, double* d_out);
/* jp_fit on this rank's node block with a global normalisation: stages 2-4 of eval_grid! (reference
 * src/joint_posterior.jl:180,186) as ONE asynchronous queue per rank.  Tensor-core GLM path: observation-sharded FP64 prep,
 * exchange of (sums, bounds), series length decided ON THE DEVICE (every instantiation of the kernel is launched and all
 * but the chosen one exit at once), coefficient rows pushed into the peers' bulk regions, kernel, local (max, sum),
 * exchange, scale.  Other families / JP_PATH_FP64: plugin kernel, (max, sum), exchange, scale.  When the a-priori bounds of
 * the tensor-core path do not hold, the first blocking call on the posterior (jp_get_*, jp_marginal_coords_p2p,
 * jp_fit_p2p_check) returns JP_ERR_UNSUPPORTED on EVERY rank (they see the same bits): call again with JP_PATH_FP64. */
int jp_fit_p2p(jp_posterior* post, const jp_fit_args* args, jp_comm* comm);
int jp_fit_p2p_check(jp_posterior* post);
/* OBSERVATION sharding (SURVEY 8e "alternative worth benchmarking"; the natural decomposition when N grows with the GPU
 * count): every rank uploads and keeps ITS rows only (jp_data of N / world observations, no replication of X), the
 * posterior handle covers ALL grid nodes.  Tensor-core GLM path only: the ranks exchange the slice sums and bounds, decide
 * the series length on the device, run the kernel over (all nodes) x (own observations), push their per-pair partial sums to
 * every peer and finish in rank order -- every rank then holds the complete, bit-identical posterior, so jp_marginal_coords
 * and the jp_get_* downloads work on it as after jp_fit and stage 5 needs no exchange.  bulk_bytes of the communicator:
 * 16 bytes x ceil(nodes / 2 + 1) x world.  JP_ERR_UNSUPPORTED: not a GLM / constrained coordinates -- shard the nodes.
 * jp_mode_p2p: the Newton iteration of jp_mode on row-sharded GLM data, its sums added over the ranks inside the library. */
int jp_fit_p2p_obs(jp_posterior* post, const jp_fit_args* args, jp_comm* comm);
int jp_mode_p2p(jp_ctx* ctx, const jp_data* data, jp_comm* comm, int d, const int* h_transform, double* h_x, double* h_H,
                double* neg_min, int* evals);
/* marginal(jp, f) of K coordinates of the node-sharded posterior, globally: moments, exchange, knot candidates, exchange,
 * combine in rank order (every rank gets identical bits); results to the host (blocking).  Reference
 * src/marginal_posterior.jl:117-123, src/interp.jl:448-457. */
int jp_marginal_coords_p2p(jp_posterior* post, jp_comm* comm, int K, const int* h_coords, double* h_mu, double* h_sigma,
                           double* h_value_nodes, double* h_weight_nodes);

/* results to the host (blocking).  h_theta: d x M_local row-major (coordinate k of node m at
 * [k*M_local + m]); h_logdens: log-density + neg_min per node; h_density: normalised weights. */
int jp_get_theta(jp_posterior* post, double* h_theta);
/* RawBuild results only: the d x M_local unconstrained node matrix (row-major like h_theta), the `cache` field of the
 * reference's raw grid (src/marginal_posterior.jl:69,107) */
int jp_get_cache(jp_posterior* post, double* h_x);
int jp_get_logdens(jp_posterior* post, double* h_logdens);
int jp_get_density(jp_posterior* post, double* h_density);
/* device views for zero-copy consumers (valid until jp_posterior_free) */
const double* jp_dev_theta(const jp_posterior* post);
const double* jp_dev_density(const jp_posterior* post);
/* which path the last jp_fit took (JP_PATH_FP64 or JP_PATH_TC) */
int jp_fit_path_used(const jp_posterior* post);
/* a-priori error figures of the tensor-core path for the last fit that tried it (csrc/jp_glm_tc.cu,
 * jp_tc_choose_order): h_out8[0] = max |Delta eta| over (node, observation) pairs, [1] = truncation bound of the
 * link-remainder series at |z| <= 6, [2] = statistical rounding estimate, [3] = series coefficients used
 * (0 = bounds not met, FP64 kernel used), [4] = worst-case rounding bound, [5] = 1 when the coefficients are the
 * economised ones (next odd / even order folded into the kept orders, tools/gen_fold.py); [6..7] reserved */
int jp_fit_diagnostics(const jp_posterior* post, double* h_out8);

/* ---------------------------------------------------------------------------------------------
 * Stage 5 -- marginal(jp, f): weighted mean / sigma and the 100-knot Grid CDF.
 * Replaces weights_values + marginal + Grid, reference src/marginal_posterior.jl:98-123,
 * src/interp.jl:21-31,448-457.  Batched over K functions f.
 *   coordinate selectors f(Theta) = Theta[coord] are evaluated on the device (zero-copy views of
 *   the Theta SoA); arbitrary host closures are supported by uploading their values f(Theta_m).
 * Outputs per marginal k: h_mu[k], h_sigma[k], h_value_nodes[k*100 ..], h_weight_nodes[k*100 ..].
 * --------------------------------------------------------------------------------------------- */
#define JP_GRID_KNOTS 100
int jp_marginal_coords(jp_posterior* post, int K, const int* h_coords, double* h_mu, double* h_sigma,
                       double* h_value_nodes, double* h_weight_nodes);
int jp_marginal_values(jp_posterior* post, int K, const double* h_values /* K x M_local */, double* h_mu,
                       double* h_sigma, double* h_value_nodes, double* h_weight_nodes);
/* sorted (values, weights) of marginal k of the last jp_marginal_* call: the `wv` field of the
 * reference's marginal struct after simultaneous_sort!, src/interp.jl:21-26.  Blocking. */
int jp_marginal_sorted(jp_posterior* post, int k, double* h_sorted_values, double* h_sorted_weights,
                       double* h_cum_weights);

/* update_MarginalBuffer! / Vandermonde! of the smooth-CDF path, reference src/marginal_posterior.jl:10-67 (the part
 * of `marginal(jp, f, Normal)` that touches all M nodes; jp_marginal_smooth below runs the 9-parameter fit on it): for marginal k of the last jp_marginal_coords / _values call on an UNSHARDED
 * posterior, the stable sort permutation (0-based node indices), the cumulative weights in sorted order, and the
 * 10 x M column-major design matrix with column i = (1, z, z^2, .., z^9), z = (v_sorted[i] - mu) / sigma.  Blocking. */
int jp_marginal_buffer(jp_posterior* post, int k, long long* h_ind, double* h_cum_weights, double* h_V, double* h_mu,
                       double* h_sigma);

/* The 100-knot Grid of the first K marginals of the last jp_marginal_* call recomputed from the explicit stable
 * sort + cumulative sum, exactly as the reference does it (src/interp.jl:21-31,448-457).  The default path above
 * gets the same knots from one binning pass without sorting; this entry point exists to cross-check it. */
int jp_marginal_knots_from_sort(jp_posterior* post, int K, double* h_value_nodes, double* h_weight_nodes);

/* multi-GPU marginal in phases (sort-free splitter histogram):
 *   jp_marginal_local_moments: d_out[k*4 + {0,1,2,3}] = (sum w v, sum w v^2, min v, max v) on the shard
 *   jp_marginal_local_knots:   given the GLOBAL (min, max) per marginal in d_minmax[k*2 + {0,1}], for each
 *     interior knot i = 1..98 writes d_out[(k*98 + i-1)*6 + {0..5}] =
 *       (S = sum of w over v <= x_i,  pred = max v <= x_i (-inf if none),
 *        succ = min v > x_i (+inf if none), global index of the lowest-index element attaining succ,
 *        its weight, the knot value x_i)
 *   the host combines over ranks (jointposteriors.jl_b200.distributed) into the 100-knot Grid.
 * h_coords may be NULL when d_values (K x M_local, device) is given instead. */
int jp_marginal_local_moments(jp_posterior* post, int K, const int* h_coords, const double* d_values, double* d_out);
int jp_marginal_local_knots(jp_posterior* post, int K, const int* h_coords, const double* d_values,
                            const double* d_minmax, double* d_out);
/* The same two phases driven by the gathered buffers themselves, so that the host does no arithmetic between the
 * collectives (two all_gathers and four launches per batch of K marginals):
 *   jp_marginal_local_knots_gathered: the global (min, max) are taken from the all_gathered moments
 *     d_gathered_moments[world][K][4]
 *   jp_marginal_combine_gathered:     combines d_gathered_moments and the all_gathered candidates
 *     d_gathered_cands[world][K][98][6] over the ranks (fixed rank order: every rank gets identical bits) into
 *     mu, sigma and the 100-knot Grid of each marginal, exactly like jp_marginal_coords (reference
 *     src/marginal_posterior.jl:118-122, src/interp.jl:448-457); results to the host (blocking). */
int jp_marginal_local_knots_gathered(jp_posterior* post, int K, const int* h_coords, const double* d_values,
                                     const double* d_gathered_moments, int world, double* d_out);
int jp_marginal_combine_gathered(jp_posterior* post, int K, int world, const double* d_gathered_moments,
                                 const double* d_gathered_cands, double* h_mu, double* h_sigma, double* h_value_nodes,
                                 double* h_weight_nodes);

/* quantile(::Grid, p) and cdf(::Grid, x): reference src/interp.jl:458-481 (host, no GPU).
 * Field order as in the reference: Grid(weights, values). */
double jp_quantile(const double* h_weight_nodes, const double* h_value_nodes, int n, double p);
double jp_cdf(const double* h_weight_nodes, const double* h_value_nodes, int n, double x);

/* Smooth CDF of a marginal: marginal(jp, f, Normal) -> NestedPolyGLM, reference src/marginal_posterior.jl:124-129 and
 * src/interp.jl:13-17 (struct: beta, theta, d = Normal(mu, sigma)), :377-384 (the fit).
 *   F(x) = Phi(P(Q(z))), z = (x - mu) / sigma, Q(z) = z^3 + l z^2 + m z + n, P(y) = a y^3 + b y^2 + c y + d;
 *   beta[0..9]  coefficients of P(Q(z)) in ascending powers of z        (update_ab!, src/interp.jl:203-235)
 *   theta[0..6] = (a, c, b, d, m, l, n)                                 (update_bt!, src/interp.jl:193-199)
 *   phi[0..8]   the unconstrained parameters the fit stopped at, objective = ntl_likelihood! there (src/interp.jl:108-111) */
typedef struct jp_smooth_cdf {
  double beta[10];
  double theta[7];
  double phi[9];
  double mu, sigma;
  double objective, grad_inf_norm;
  int iterations, evaluations, converged;
} jp_smooth_cdf;
/* NestedPolyGLM(m, Normal(mu, sigma)) for marginal k of the last jp_marginal_coords / _values call on an UNSHARDED
 * posterior: sort + design matrix on the device (as jp_marginal_buffer), then the minimisation of ntl_likelihood!
 * (src/interp.jl:81-111): a short BFGS phase with a backtracking line search (the optimiser the reference asks of Optim,
 * :380; one kernel launch over all M nodes per objective / score evaluation) followed by a saddle-free trust-region Newton
 * iteration whose Hessian comes from ONE batched launch (the score at phi and at its nine forward-difference neighbours).
 * phi_init: 9 starting values or NULL (zeros; the reference's MarginalBuffer.init is in the absent LogDensities package);
 * max_iter 0 = 300 iterations in all, g_tol 0 = 1e-8 (Optim's default).  converged: 1 = the score's infinity norm is below
 * g_tol; 2 = the decrease a full Newton step predicts is below 1e-8 max(1, |objective|) (the objective's curvatures span ten
 * orders of magnitude, the infinity norm of the score does not reach 1e-8 with any optimiser); 0 = stopped at
 * the iteration cap (not an error).  JP_SMOOTH_BFGS=1 in the environment: BFGS only, max_iter 0 = 1000.  Blocking. */
int jp_marginal_smooth(jp_posterior* post, int k, const double* phi_init, int max_iter, double g_tol, jp_smooth_cdf* out);
/* The same with the per-function buffer cache of the reference (`get!(() -> MarginalBuffer(n), M.MarginalBuffers, f)`,
 * src/marginal_posterior.jl:10,71): `key` names the marginal function (the host keeps one key per f); the sorted design matrix
 * and cumulative weights built for a key since the last fit of `post` are kept on the device (up to 4 functions) and a
 * repeated call reuses them -- no sort, no Vandermonde pass (*cache_hit = 1).  k < 0 only looks the key up: on a miss nothing
 * is computed and *cache_hit = 0; k >= 0 names the marginal of the last jp_marginal_* call that provides the values. */
int jp_marginal_smooth_keyed(jp_posterior* post, int k, long long key, const double* phi_init, int max_iter, double g_tol,
                             jp_smooth_cdf* out, int* cache_hit);
/* ntl_likelihood! and ntscore! at a given phi (src/interp.jl:81-111) for marginal k; grad9 / beta10 / theta7 may be NULL. */
int jp_smooth_objective(jp_posterior* post, int k, const double* phi, double* f, double* grad9, double* beta10,
                        double* theta7);
/* cdf / pdf / quantile of a fitted NestedPolyGLM: reference src/interp.jl:365-374 (host, no GPU). */
double jp_smooth_cdf_eval(const jp_smooth_cdf* s, double x);
double jp_smooth_pdf_eval(const jp_smooth_cdf* s, double x);
double jp_smooth_quantile_eval(const jp_smooth_cdf* s, double p);

#ifdef __cplusplus
}
#endif
#endif /* JPCUDA_H */
