// jp_api.cu -- handles and plumbing of the C ABI (include/jpcuda.h): context, grid cache,
// observation upload, posterior buffers, host getters.
#include <algorithm>
#include <cstring>
#include <new>
#include "jp_common.cuh"
#include "jp_family.cuh"

// queue the download of the normalised weights on the context's `side` stream, behind everything queued on the main stream so far
int jp_density_prefetch(jp_posterior* p) {
  jp_ctx* ctx = p->ctx;
  const size_t n = (size_t)p->M;
  if (n == 0 || n > ((size_t)1 << 24)) return JP_OK;                         // (128 MB of pinned memory at most)
  if (n > ctx->h_density_cap) {
    if (ctx->h_density) {
      JP_CUDA(cudaEventSynchronize(ctx->ev_density));
      JP_CUDA(cudaFreeHost(ctx->h_density));
      ctx->h_density = nullptr;
      ctx->h_density_cap = 0;
    }
    JP_CUDA(cudaMallocHost(&ctx->h_density, n * 8));
    ctx->h_density_cap = n;
  }
  ctx->h_density_owner = nullptr;
  JP_CUDA(cudaEventRecord(ctx->ev_stage4, ctx->stream));
  JP_CUDA(cudaStreamWaitEvent(ctx->side, ctx->ev_stage4, 0));
  JP_CUDA(cudaMemcpyAsync(ctx->h_density, p->d_density, n * 8, cudaMemcpyDeviceToHost, ctx->side));
  JP_CUDA(cudaEventRecord(ctx->ev_density, ctx->side));
  ctx->h_density_owner = p;
  ctx->h_density_gen = p->fit_gen;
  return JP_OK;
}


extern "C" {

int jp_ctx_create(int device, jp_ctx** out) {
  JP_REQUIRE(out, "jp_ctx_create: null output");
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    jp_set_error("jp_ctx_create: no CUDA device (%s); libjpcuda has no CPU fallback",
                 e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    cudaGetLastError();
    return JP_ERR_NO_DEVICE;
  }
  JP_REQUIRE(device >= 0 && device < count, "jp_ctx_create: device %d out of range [0,%d)", device, count);
  JP_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  JP_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    jp_set_error("jp_ctx_create: device %d is sm_%d%d; this library is built for sm_100a (B200) only", device,
                 prop.major, prop.minor);
    return JP_ERR_UNSUPPORTED;
  }
  jp_ctx* ctx = new (std::nothrow) jp_ctx();
  if (!ctx) return JP_ERR_ALLOC;
  ctx->device = device;
  ctx->sm_count = prop.multiProcessorCount;
  JP_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
  ctx->own_stream = true;
  {
    // keep freed blocks cached in the pool instead of returning them to the driver at every sync
    cudaMemPool_t pool;
    JP_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
    unsigned long long keep = ~0ull;
    JP_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
  }
  JP_CUDA(jp_dmalloc(ctx, &ctx->d_scratch, sizeof(double) * JP_SCRATCH_DOUBLES));
  JP_CUDA(jp_dmalloc(ctx, &ctx->d_counters, sizeof(unsigned int) * JP_COUNTERS));
  JP_CUDA(cudaMemsetAsync(ctx->d_counters, 0, sizeof(unsigned int) * JP_COUNTERS, ctx->stream));
  JP_CUDA(jp_dmalloc(ctx, &ctx->d_bpart, sizeof(double) * JP_BPART_DOUBLES));
  JP_CUDA(cudaMallocHost(&ctx->h_pinned, sizeof(double) * JP_PINNED_DOUBLES));
  JP_CUDA(cudaStreamCreateWithFlags(&ctx->side, cudaStreamNonBlocking));
  JP_CUDA(cudaStreamCreateWithFlags(&ctx->side2, cudaStreamNonBlocking));
  JP_CUDA(cudaEventCreateWithFlags(&ctx->ev_join2, cudaEventDisableTiming));
  JP_CUDA(cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
  JP_CUDA(cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming));
  JP_CUDA(cudaEventCreateWithFlags(&ctx->ev_pinned, cudaEventDisableTiming));
  JP_CUDA(cudaEventCreateWithFlags(&ctx->ev_density, cudaEventDisableTiming));
  JP_CUDA(cudaEventCreateWithFlags(&ctx->ev_stage4, cudaEventDisableTiming));
  JP_CUDA(cudaEventCreate(&ctx->ev_k0));
  JP_CUDA(cudaEventCreate(&ctx->ev_k1));
  *out = ctx;
  return JP_OK;
}

static void free_grid(jp_ctx* ctx, jp_grid* g) {
  if (!g) return;
  jp_dfree(ctx, g->d_idx);
  jp_dfree(ctx, g->d_w);
  jp_dfree(ctx, g->d_hzz);
  delete g;
}

int jp_ctx_destroy(jp_ctx* ctx) {
  if (!ctx) return JP_OK;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  for (auto& kv : ctx->grids) free_grid(ctx, kv.second);
  jp_dfree(ctx, ctx->d_scratch); jp_dfree(ctx, ctx->d_counters); jp_dfree(ctx, ctx->d_bpart);
  cudaStreamSynchronize(ctx->stream);
  cudaFreeHost(ctx->h_pinned);
  if (ctx->h_density) cudaFreeHost(ctx->h_density);
  cudaEventDestroy(ctx->ev_density);
  cudaEventDestroy(ctx->ev_stage4);
  cudaFree(ctx->d_rule_nodes[0]); cudaFree(ctx->d_rule_nodes[1]);
  cudaStreamSynchronize(ctx->side);
  cudaStreamDestroy(ctx->side);
  cudaStreamSynchronize(ctx->side2);
  cudaStreamDestroy(ctx->side2);
  cudaEventDestroy(ctx->ev_join2);
  cudaEventDestroy(ctx->ev_fork);
  cudaEventDestroy(ctx->ev_join);
  cudaEventDestroy(ctx->ev_pinned);
  cudaEventDestroy(ctx->ev_k0);
  cudaEventDestroy(ctx->ev_k1);
  if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
  return JP_OK;
}

int jp_ctx_set_stream(jp_ctx* ctx, void* cuda_stream) {
  JP_REQUIRE(ctx, "jp_ctx_set_stream: null ctx");
  JP_CUDA(cudaStreamSynchronize(ctx->stream));
  if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
  ctx->stream = (cudaStream_t)cuda_stream;
  ctx->own_stream = false;
  return JP_OK;
}

int jp_ctx_sync(jp_ctx* ctx) {
  JP_REQUIRE(ctx, "jp_ctx_sync: null ctx");
  JP_CUDA(cudaStreamSynchronize(ctx->stream));
  return JP_OK;
}

long long jp_ctx_launch_count(const jp_ctx* ctx) { return ctx ? ctx->launches : 0; }

int jp_ctx_trace(jp_ctx* ctx, int on) {
  JP_REQUIRE(ctx, "jp_ctx_trace: null ctx");
  for (auto& m : ctx->trace.marks) cudaEventDestroy(m.second);
  ctx->trace.marks.clear();
  ctx->trace.on = on != 0;
  return JP_OK;
}

int jp_ctx_trace_dump(jp_ctx* ctx, char* buf, int len) {
  JP_REQUIRE(ctx && buf && len > 0, "jp_ctx_trace_dump: bad argument");
  JP_CUDA(cudaStreamSynchronize(ctx->stream));
  JP_CUDA(cudaStreamSynchronize(ctx->side));
  JP_CUDA(cudaStreamSynchronize(ctx->side2));
  int pos = 0;
  buf[0] = 0;
  for (size_t i = 0; i < ctx->trace.marks.size(); ++i) {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, ctx->trace.marks[0].second, ctx->trace.marks[i].second) != cudaSuccess) {
      cudaGetLastError();
      continue;
    }
    int n = snprintf(buf + pos, (size_t)(len - pos), "%s\t%.1f\n", ctx->trace.marks[i].first, ms * 1e3);
    if (n < 0 || n >= len - pos) break;
    pos += n;
  }
  return JP_OK;
}

int jp_ctx_last_kernel_ms(jp_ctx* ctx, float* ms) {
  JP_REQUIRE(ctx && ms, "jp_ctx_last_kernel_ms: null argument");
  JP_REQUIRE(ctx->ev_valid, "jp_ctx_last_kernel_ms: no log-density kernel has been launched on this context");
  JP_CUDA(cudaEventSynchronize(ctx->ev_k1));
  JP_CUDA(cudaEventElapsedTime(ms, ctx->ev_k0, ctx->ev_k1));
  return JP_OK;
}

// ------------------------------------------------------------------------------------ stage 1
int jp_grid_get(jp_ctx* ctx, int rule, int d_eff, int level, jp_grid** out) {
  JP_REQUIRE(ctx && out, "jp_grid_get: null argument");
  JP_CUDA(cudaSetDevice(ctx->device));
  auto key = std::make_tuple(rule, d_eff, level);
  auto it = ctx->grids.find(key);
  if (it != ctx->grids.end()) {
    *out = it->second;
    return JP_OK;
  }
  jp_grid* g = new (std::nothrow) jp_grid();
  if (!g) return JP_ERR_ALLOC;
  int st = jp_grid_build(ctx, rule, d_eff, level, g);
  if (st != JP_OK) {
    free_grid(ctx, g);
    return st;
  }
  ctx->grids[key] = g;
  *out = g;
  return JP_OK;
}

long long jp_grid_size(const jp_grid* g) { return g ? g->M : -1; }
int jp_grid_dim(const jp_grid* g) { return g ? g->d : -1; }

int jp_grid_build_stats(const jp_grid* g, long long* n_multi, long long* n_premerge) {
  JP_REQUIRE(g, "jp_grid_build_stats: null grid");
  if (n_multi) *n_multi = g->n_multi;
  if (n_premerge) *n_premerge = g->n_premerge;
  return JP_OK;
}

int jp_grid_download(const jp_grid* g, uint8_t* h_idx, double* h_w) {
  JP_REQUIRE(g, "jp_grid_download: null grid");
  JP_CUDA(cudaSetDevice(g->ctx->device));
  JP_CUDA(cudaStreamSynchronize(g->ctx->stream));
  if (h_idx) {
    std::vector<uint8_t> soa((size_t)g->M * g->d);
    JP_CUDA(cudaMemcpy(soa.data(), g->d_idx, soa.size(), cudaMemcpyDeviceToHost));
    for (long long m = 0; m < g->M; ++m)
      for (int k = 0; k < g->d; ++k) h_idx[(size_t)m * g->d + k] = soa[(size_t)k * g->M + m];
  }
  if (h_w) JP_CUDA(cudaMemcpy(h_w, g->d_w, (size_t)g->M * 8, cudaMemcpyDeviceToHost));
  return JP_OK;
}

int jp_rule_info(int rule, int* levels, int* nmax, int* h_npts, double* h_nodes, double* h_weights) {
  JP_REQUIRE(rule == 0 || rule == 1, "jp_rule_info: unknown rule %d", rule);
  JpRule R = jp_get_rule(rule);
  if (levels) *levels = R.levels;
  if (nmax) *nmax = R.nmax;
  if (h_npts) std::memcpy(h_npts, R.npts, sizeof(int) * R.levels);
  if (h_nodes) std::memcpy(h_nodes, R.nodes, sizeof(double) * R.nmax);
  if (h_weights) std::memcpy(h_weights, R.weights, sizeof(double) * R.levels * R.nmax);
  return JP_OK;
}

int jp_rule_level_nodes(int rule, int level, int* h_index) {
  if (rule != 0 && rule != 1) return -1;
  JpRule R = jp_get_rule(rule);
  if (level < 1 || level > R.levels) return -1;
  const int n = R.npts[level - 1];
  if (h_index)
    for (int j = 0; j < n; ++j) h_index[j] = R.index[(size_t)(level - 1) * R.nmax + j];
  return n;
}

int jp_grid_level_cap(const jp_grid* g) { return g ? std::min(g->level, jp_get_rule(g->rule).levels) : -1; }

// ------------------------------------------------------------------------------------ data
int jp_data_upload(jp_ctx* ctx, int family, long long N, int ncols, const double* h_obs, const double* h_hyper,
                   int n_hyper, jp_data** out) {
  JP_REQUIRE(ctx && h_obs && out, "jp_data_upload: null argument");
  JP_REQUIRE(N >= 1 && ncols >= 1, "jp_data_upload: empty data (N=%lld, ncols=%d)", N, ncols);
  JP_REQUIRE(n_hyper >= 0 && n_hyper <= JP_MAX_HYPER, "jp_data_upload: n_hyper=%d out of range", n_hyper);
  JP_REQUIRE(jp_find_family(family) != nullptr, "jp_data_upload: family %d is not registered", family);
  JP_CUDA(cudaSetDevice(ctx->device));
  jp_data* dt = new (std::nothrow) jp_data();
  if (!dt) return JP_ERR_ALLOC;
  dt->ctx = ctx; dt->family = family; dt->N = N; dt->ncols = ncols; dt->n_hyper = n_hyper;
  for (int i = 0; i < n_hyper; ++i) dt->hyper[i] = h_hyper[i];
  size_t bytes = (size_t)N * ncols * sizeof(double);
  cudaError_t e = jp_dmalloc(ctx, &dt->d_obs, bytes);
  if (e == cudaSuccess) e = cudaMemcpyAsync(dt->d_obs, h_obs, bytes, cudaMemcpyHostToDevice, ctx->stream);
  if (e != cudaSuccess) {
    jp_set_error("jp_data_upload: %s", cudaGetErrorString(e));
    jp_dfree(ctx, dt->d_obs);
    delete dt;
    return JP_ERR_CUDA;
  }
  *out = dt;
  return JP_OK;
}

int jp_data_adopt_device(jp_ctx* ctx, int family, long long N, int ncols, const double* d_obs, const double* h_hyper,
                         int n_hyper, jp_data** out) {
  JP_REQUIRE(ctx && d_obs && out, "jp_data_adopt_device: null argument");
  JP_REQUIRE(N >= 1 && ncols >= 1, "jp_data_adopt_device: empty data (N=%lld, ncols=%d)", N, ncols);
  JP_REQUIRE(n_hyper >= 0 && n_hyper <= JP_MAX_HYPER, "jp_data_adopt_device: n_hyper=%d out of range", n_hyper);
  JP_REQUIRE(jp_find_family(family) != nullptr, "jp_data_adopt_device: family %d is not registered", family);
  cudaPointerAttributes attr;
  JP_CUDA(cudaSetDevice(ctx->device));
  cudaError_t e = cudaPointerGetAttributes(&attr, d_obs);
  if (e != cudaSuccess || attr.type != cudaMemoryTypeDevice || attr.device != ctx->device) {
    cudaGetLastError();
    jp_set_error("jp_data_adopt_device: d_obs is not device memory of GPU %d", ctx->device);
    return JP_ERR_BAD_ARG;
  }
  jp_data* dt = new (std::nothrow) jp_data();
  if (!dt) return JP_ERR_ALLOC;
  dt->ctx = ctx; dt->family = family; dt->N = N; dt->ncols = ncols; dt->n_hyper = n_hyper;
  for (int i = 0; i < n_hyper; ++i) dt->hyper[i] = h_hyper[i];
  dt->d_obs = const_cast<double*>(d_obs);
  dt->owns_obs = false;
  *out = dt;
  return JP_OK;
}

int jp_data_free(jp_data* data) {
  if (!data) return JP_OK;
  cudaSetDevice(data->ctx->device);
  jp_tc_data_free(data);
  if (data->owns_obs) jp_dfree(data->ctx, data->d_obs);
  delete data;
  return JP_OK;
}

// ------------------------------------------------------------------------------------ posterior
int jp_posterior_create(jp_ctx* ctx, const jp_grid* g, const jp_data* data, const jp_fit_args* args,
                        jp_posterior** out) {
  JP_REQUIRE(ctx && g && data && args && out, "jp_posterior_create: null argument");
  JP_REQUIRE(args->d >= 1 && args->d <= JP_MAX_D, "jp_posterior_create: d=%d out of range [1,%d]", args->d, JP_MAX_D);
  JP_REQUIRE(args->p >= 1 && args->p <= args->d, "jp_posterior_create: p=%d must be in [1,d=%d]", args->p, args->d);
  JP_REQUIRE(args->p == g->d, "jp_posterior_create: U has %d columns but the grid has dimension %d", args->p, g->d);
  long long m0 = args->node_begin, m1 = args->node_end;
  if (m1 < 0) m1 = g->M;
  JP_REQUIRE(m0 >= 0 && m0 < m1 && m1 <= g->M, "jp_posterior_create: node shard [%lld,%lld) not inside [0,%lld)", m0, m1,
             g->M);
  JP_CUDA(cudaSetDevice(ctx->device));
  jp_posterior* p = new (std::nothrow) jp_posterior();
  if (!p) return JP_ERR_ALLOC;
  p->ctx = ctx; p->grid = g; p->data = data; p->d = args->d; p->p = args->p;
  p->m0 = m0; p->m1 = m1; p->M = m1 - m0;
  size_t M = (size_t)p->M;
  cudaError_t e = cudaSuccess;
  auto A = [&](void** ptr, size_t bytes) { if (e == cudaSuccess) e = jp_dmalloc(ctx, ptr, bytes); };
  A((void**)&p->d_theta, M * p->d * 8);
  A((void**)&p->d_a, M * 8);
  A((void**)&p->d_logdens, M * 8);
  A((void**)&p->d_density, M * 8);
  A((void**)&p->d_part, M * (JP_POST_PART_SPLITS + 1) * 8);
  A((void**)&p->d_stats, 16 * 8);
  // the per-fit constants (mu_hat d, U d x p, transform codes d) share ONE block: one host-to-device copy per fit
  A((void**)&p->d_mu, (size_t)(p->d + p->d * p->p) * 8 + (size_t)p->d * 4);
  if (e == cudaSuccess) {
    p->d_U = p->d_mu + p->d;
    p->d_tcode = reinterpret_cast<int*>(p->d_U + (size_t)p->d * p->p);
  }
  if (e != cudaSuccess) {
    jp_set_error("jp_posterior_create: %s", cudaGetErrorString(e));
    jp_posterior_free(p);
    return JP_ERR_CUDA;
  }
  *out = p;
  return JP_OK;
}

int jp_posterior_free(jp_posterior* p) {
  if (!p) return JP_OK;
  cudaSetDevice(p->ctx->device);
  jp_ctx* c = p->ctx;   // stream-ordered frees: work already queued on the ctx stream still sees the buffers
  if (c->h_density_owner == p) {      // a queued download of this posterior's weights must finish before its buffer is reused
    c->h_density_owner = nullptr;
    cudaStreamWaitEvent(c->stream, c->ev_density, 0);
  }
  jp_dfree(c, p->d_theta); jp_dfree(c, p->d_a); jp_dfree(c, p->d_logdens); jp_dfree(c, p->d_density); jp_dfree(c, p->d_part);
  jp_dfree(c, p->d_stats); jp_dfree(c, p->d_mu); jp_tc_post_free(p);      // d_U, d_tcode live in d_mu's block
  jp_dfree(c, p->d_vals); jp_dfree(c, p->d_bins); jp_dfree(c, (void*)p->d_vptr); jp_dfree(c, p->d_perm_a); jp_dfree(c, p->d_perm_b); jp_dfree(c, p->d_hist);
  jp_dfree(c, p->d_sv); jp_dfree(c, p->d_sw); jp_dfree(c, p->d_cw); jp_dfree(c, p->d_mout);
  jp_dfree(c, p->d_cmom); jp_dfree(c, p->d_coords); jp_dfree(c, p->d_cand);
  for (auto& e : p->design_cache) { jp_dfree(c, e.d_V); jp_dfree(c, e.d_cw); }
  delete p;
  return JP_OK;
}

long long jp_posterior_size(const jp_posterior* p) { return p ? p->M : -1; }
const double* jp_dev_theta(const jp_posterior* p) { return p ? p->d_theta : nullptr; }
const double* jp_dev_density(const jp_posterior* p) { return p ? p->d_density : nullptr; }
int jp_fit_path_used(const jp_posterior* p) { return p ? p->path_used : 0; }
int jp_fit_diagnostics(const jp_posterior* p, double* h_out8) {
  JP_REQUIRE(p && h_out8, "jp_fit_diagnostics: null argument");
  jp_fit_tc_verify(const_cast<jp_posterior*>(p));      // a device-side decision not yet read back (blocks once)
  for (int i = 0; i < 8; ++i) h_out8[i] = p->tc_bounds[i];
  return JP_OK;
}

static int download(jp_posterior* p, const double* d_src, double* h_dst, size_t n) {
  JP_REQUIRE(p && h_dst, "jp_get_*: null argument");
  // The caller's array is pageable (a Julia Vector / numpy array): results that fit the context's pinned buffer are
  // DMA-ed there at PCIe speed and copied out by the host; larger ones take the driver's staged pageable path.
  JP_CUDA(cudaSetDevice(p->ctx->device));
  if (n <= JP_PINNED_DOUBLES - JP_PINNED_TAIL_DOUBLES) {
    JP_CUDA(jp_pinned_acquire(p->ctx));      // no earlier host-to-device copy may still be reading the staging buffer
    JP_TRY(jp_fit_tc_verify_prefetch(p));
    JP_CUDA(cudaMemcpyAsync(p->ctx->h_pinned, d_src, n * 8, cudaMemcpyDeviceToHost, p->ctx->stream));
    JP_CUDA(cudaStreamSynchronize(p->ctx->stream));
    std::memcpy(h_dst, p->ctx->h_pinned, n * 8);
    return jp_fit_tc_verify(p);
  }
  JP_CUDA(cudaMemcpyAsync(h_dst, d_src, n * 8, cudaMemcpyDeviceToHost, p->ctx->stream));
  JP_CUDA(cudaStreamSynchronize(p->ctx->stream));
  return jp_fit_tc_verify(p);
}
int jp_get_theta(jp_posterior* p, double* h) {
  JP_REQUIRE(p && h, "jp_get_theta: null argument");
  if (!p->raw) return download(p, p->d_theta, h, (size_t)p->M * p->d);
  // RawBuild: construct every coordinate from the unconstrained cache into a temporary, then download
  JP_CUDA(cudaSetDevice(p->ctx->device));
  double* tmp = nullptr;
  JP_CUDA(jp_dmalloc(p->ctx, &tmp, (size_t)p->M * p->d * 8));
  int st = jp_construct_columns(p, p->d, nullptr, tmp);
  if (st == JP_OK) st = download(p, tmp, h, (size_t)p->M * p->d);
  jp_dfree(p->ctx, tmp);
  return st;
}
int jp_get_cache(jp_posterior* p, double* h) {
  JP_REQUIRE(p && h, "jp_get_cache: null argument");
  JP_REQUIRE(p->raw, "jp_get_cache: the posterior was not fitted as a RawBuild (jp_fit_args.raw)");
  return download(p, p->d_theta, h, (size_t)p->M * p->d);
}
int jp_get_logdens(jp_posterior* p, double* h) { return download(p, p ? p->d_logdens : nullptr, h, p ? (size_t)p->M : 0); }
int jp_get_density(jp_posterior* p, double* h) {
  JP_REQUIRE(p && h, "jp_get_density: null argument");
  jp_ctx* ctx = p->ctx;
  if (ctx->h_density_owner == p && ctx->h_density_gen == p->fit_gen) {      // queued by jp_fit behind stage 4
    JP_CUDA(cudaSetDevice(ctx->device));
    JP_CUDA(cudaEventSynchronize(ctx->ev_density));
    std::memcpy(h, ctx->h_density, (size_t)p->M * 8);
    return jp_fit_tc_verify(p);
  }
  return download(p, p->d_density, h, (size_t)p->M);
}

}  // extern "C"
