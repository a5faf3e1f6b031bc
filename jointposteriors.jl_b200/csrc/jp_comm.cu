// jp_comm.cu -- the exchanges of the node-sharded path inside the library, over NVLink peer memory.
//
// The reference is single-process; sharding its eval_grid! loop (reference src/joint_posterior.jl:180,186)
// and its marginal (reference src/marginal_posterior.jl:117-123, src/interp.jl:448-457) over the GPUs of one box needs
// four tiny all_gathers per fit + marginal batch (slice sums and bounds, (max, sum), moments, knot candidates) and one
// larger one (the coefficient rows of the observation-sharded tensor-core prep).  Here they are kernels of this library
// that store straight into the peers' memory (CUDA IPC mapping, NVLink / NVSwitch) and spin on flags in their own: no
// NCCL call, no host synchronisation between the phases, and a host language without an NCCL binding (the Julia shim
// of INTEGRATION.md) only has to move one 64-byte handle per rank once, by any means it has.
// Layout and protocol: jp_common.cuh (struct jp_comm, JpCommDev).
#include <cstdlib>
#include <cstring>
#include <new>
#include "jp_common.cuh"

// per-rank payload capacities (doubles) of the small channels
static size_t chan_rank_cap(int chan) {
  switch (chan) {
    case JP_CH_PREP: return 640;                                    // d + d (d + 1) / 2 + 1 + bounds at d <= 32
    case JP_CH_STATS: return 8;
    case JP_CH_MOM: return 4 * (size_t)JP_COMM_KMAX;
    case JP_CH_KNOTS: return (size_t)JP_COMM_KMAX * (JP_GRID_KNOTS - 2) * 6;
    case JP_CH_USER: return 4096;
    default: return 0;                                              // JP_CH_BULK: flags only
  }
}

JpCommDev jp_comm_dev(const jp_comm* c) {
  JpCommDev v;
  v.peer = c->d_peer;
  v.self = c->mailbox;
  v.rank = c->rank;
  v.world = c->world;
  v.timeout_ns = c->timeout_ns;
  return v;
}

// Block p serves peer p.  n doubles from src into slot `rank` of the (chan, parity) region of peer p's mailbox, the flag,
// then the wait for p's own contribution to this rank's mailbox.
__global__ void __launch_bounds__(512) jp_comm_exchange_kernel(const JpCommDev c, int chan, int parity, unsigned long long seq,
                                                               const double* __restrict__ src, int n, size_t data_off) {
  const int p = blockIdx.x;
  double* dst = reinterpret_cast<double*>(c.peer[p] + data_off) + (size_t)c.rank * n;
  for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    jp_st_release_sys(jp_comm_flag(c.peer[p], chan, parity, c.rank), seq);
    jp_comm_wait_flag(c, chan, parity, p, seq);
  }
}

// lane s waits for sender s
__global__ void jp_comm_wait_kernel(const JpCommDev c, int chan, int parity, unsigned long long seq) {
  if ((int)threadIdx.x < c.world) jp_comm_wait_flag(c, chan, parity, (int)threadIdx.x, seq);
}

int jp_comm_exchange(jp_comm* c, int chan, const double* d_src, int n, const double** d_gathered) {
  JP_REQUIRE(c && c->connected, "jp_comm: the mailboxes are not connected (jp_comm_connect_ipc / jp_comm_connect_local)");
  JP_REQUIRE(chan >= 0 && chan < JP_COMM_NCHAN && chan != JP_CH_BULK, "jp_comm: bad channel %d", chan);
  JP_REQUIRE(n >= 1 && (size_t)n * c->world <= c->cap_doubles[chan], "jp_comm: %d doubles per rank exceed channel %d", n, chan);
  const unsigned long long seq = ++c->seq[chan];
  const int parity = (int)(seq & 1ull);
  jp_comm_exchange_kernel<<<c->world, n >= 2048 ? 512 : 128, 0, c->ctx->stream>>>(jp_comm_dev(c), chan, parity, seq, d_src, n,
                                                                                   c->data_off[chan][parity]);
  JP_CHECK_LAUNCH(c->ctx);
  if (d_gathered) *d_gathered = reinterpret_cast<const double*>(c->mailbox + c->data_off[chan][parity]);
  return JP_OK;
}

int jp_comm_bulk_begin(jp_comm* c, unsigned long long* seq) {
  JP_REQUIRE(c && c->connected && seq, "jp_comm: the mailboxes are not connected");
  *seq = ++c->seq[JP_CH_BULK];
  return JP_OK;
}

int jp_comm_wait(jp_comm* c, int chan, int parity, unsigned long long seq) {
  jp_comm_wait_kernel<<<1, 32, 0, c->ctx->stream>>>(jp_comm_dev(c), chan, parity, seq);
  JP_CHECK_LAUNCH(c->ctx);
  return JP_OK;
}

extern "C" {

int jp_comm_create(jp_ctx* ctx, int rank, int world, long long bulk_bytes, jp_comm** out) {
  JP_REQUIRE(ctx && out, "jp_comm_create: null argument");
  JP_REQUIRE(world >= 1 && world <= JP_COMM_MAX_WORLD && rank >= 0 && rank < world, "jp_comm_create: rank %d of %d (at most %d ranks)", rank,
             world, JP_COMM_MAX_WORLD);
  JP_REQUIRE(bulk_bytes >= 0, "jp_comm_create: negative bulk size");
  JP_ENTER_CTX(ctx);
  jp_comm* c = new (std::nothrow) jp_comm();
  if (!c) return JP_ERR_ALLOC;
  c->ctx = ctx; c->rank = rank; c->world = world;
  if (const char* t = getenv("JP_COMM_TIMEOUT_S")) c->timeout_ns = (long long)(atof(t) * 1e9);
  size_t off = JP_COMM_HEADER_BYTES;
  static_assert(64 + JP_COMM_NCHAN * 2 * JP_COMM_MAX_WORLD * 8 <= JP_COMM_HEADER_BYTES, "mailbox header too small");
  for (int ch = 0; ch < JP_COMM_NCHAN; ++ch) {
    c->cap_doubles[ch] = chan_rank_cap(ch) * JP_COMM_MAX_WORLD;
    for (int par = 0; par < 2; ++par) {
      c->data_off[ch][par] = off;
      off += c->cap_doubles[ch] * 8;
    }
  }
  off = (off + ((size_t)1 << 20) - 1) & ~(((size_t)1 << 20) - 1);
  c->bulk_off = off;
  c->bulk_bytes = ((size_t)bulk_bytes + 255) & ~(size_t)255;
  c->bytes = off + c->bulk_bytes;
  // plain cudaMalloc: memory of the stream-ordered pool cannot be exported through cudaIpcGetMemHandle
  cudaError_t e = cudaMalloc((void**)&c->mailbox, c->bytes);
  if (e == cudaSuccess) e = cudaMemset(c->mailbox, 0, JP_COMM_HEADER_BYTES);
  if (e == cudaSuccess) e = cudaMalloc((void**)&c->d_peer, sizeof(unsigned char*) * JP_COMM_MAX_WORLD);
  if (e == cudaSuccess) e = cudaMalloc((void**)&c->d_counter, 64);
  if (e == cudaSuccess) e = cudaMemset(c->d_counter, 0, 64);
  if (e != cudaSuccess) {
    jp_set_error("jp_comm_create: %s (mailbox of %zu bytes)", cudaGetErrorString(e), c->bytes);
    cudaFree(c->mailbox); cudaFree(c->d_peer); cudaFree(c->d_counter);
    delete c;
    return JP_ERR_CUDA;
  }
  c->peer_h[rank] = c->mailbox;
  if (world == 1) {      // nothing to connect
    JP_CUDA(cudaMemcpy(c->d_peer, c->peer_h, sizeof(unsigned char*) * JP_COMM_MAX_WORLD, cudaMemcpyHostToDevice));
    c->connected = true;
  }
  *out = c;
  return JP_OK;
}

long long jp_comm_bulk_bytes(const jp_comm* c) { return c ? (long long)c->bulk_bytes : -1; }

int jp_comm_ipc_handle(jp_comm* c, void* handle64) {
  JP_REQUIRE(c && handle64, "jp_comm_ipc_handle: null argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is expected to be 64 bytes");
  JP_ENTER_CTX(c->ctx);
  cudaIpcMemHandle_t h;
  JP_CUDA(cudaIpcGetMemHandle(&h, c->mailbox));
  std::memcpy(handle64, &h, 64);
  return JP_OK;
}

static int publish_peers(jp_comm* c) {
  JP_CUDA(cudaMemcpy(c->d_peer, c->peer_h, sizeof(unsigned char*) * JP_COMM_MAX_WORLD, cudaMemcpyHostToDevice));
  c->connected = true;
  return JP_OK;
}

int jp_comm_connect_ipc(jp_comm* c, const void* handles) {
  JP_REQUIRE(c && handles, "jp_comm_connect_ipc: null argument");
  JP_REQUIRE(!c->connected || c->world == 1, "jp_comm_connect_ipc: already connected");
  JP_ENTER_CTX(c->ctx);
  for (int r = 0; r < c->world; ++r) {
    if (r == c->rank) continue;
    cudaIpcMemHandle_t h;
    std::memcpy(&h, static_cast<const unsigned char*>(handles) + (size_t)r * 64, 64);
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      cudaGetLastError();
      jp_set_error("jp_comm_connect_ipc: cannot map the mailbox of rank %d (%s); the GPUs of the ranks need peer access", r,
                   cudaGetErrorString(e));
      return JP_ERR_CUDA;
    }
    c->peer_h[r] = static_cast<unsigned char*>(p);
    c->ipc_opened[r] = true;
  }
  return publish_peers(c);
}

int jp_comm_connect_local(jp_comm* c, jp_comm* const* peers) {
  JP_REQUIRE(c && peers, "jp_comm_connect_local: null argument");
  JP_ENTER_CTX(c->ctx);
  for (int r = 0; r < c->world; ++r) {
    JP_REQUIRE(peers[r] && peers[r]->world == c->world && peers[r]->rank == r && peers[r]->bytes == c->bytes,
               "jp_comm_connect_local: entry %d is not rank %d of the same %d-rank communicator", r, r, c->world);
    if (peers[r]->ctx->device != c->ctx->device) {
      int ok = 0;
      JP_CUDA(cudaDeviceCanAccessPeer(&ok, c->ctx->device, peers[r]->ctx->device));
      JP_REQUIRE(ok, "jp_comm_connect_local: GPU %d cannot access GPU %d", c->ctx->device, peers[r]->ctx->device);
      cudaError_t e = cudaDeviceEnablePeerAccess(peers[r]->ctx->device, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) JP_CUDA(e);
      cudaGetLastError();
    }
    c->peer_h[r] = peers[r]->mailbox;
  }
  return publish_peers(c);
}

int jp_comm_all_gather(jp_comm* c, const double* d_src, int n, double* d_out) {
  JP_REQUIRE(c && d_src && d_out, "jp_comm_all_gather: null argument");
  JP_ENTER_CTX(c->ctx);
  const double* g = nullptr;
  JP_TRY(jp_comm_exchange(c, JP_CH_USER, d_src, n, &g));
  JP_CUDA(cudaMemcpyAsync(d_out, g, (size_t)n * c->world * 8, cudaMemcpyDeviceToDevice, c->ctx->stream));
  return JP_OK;
}

int jp_comm_status(jp_comm* c) {
  JP_REQUIRE(c, "jp_comm_status: null communicator");
  JP_ENTER_CTX(c->ctx);
  JP_CUDA(cudaStreamSynchronize(c->ctx->stream));
  int err = 0;
  JP_CUDA(cudaMemcpy(&err, c->mailbox, sizeof(int), cudaMemcpyDeviceToHost));
  if (err != 0) {
    const int code = err - 1;
    jp_set_error("jp_comm: rank %d timed out after %.0f s waiting for rank %d on channel %d (a peer never reached the matching call)",
                 c->rank, c->timeout_ns * 1e-9, code / 16, code % 16);
    cudaMemset(c->mailbox, 0, sizeof(int));
    return JP_ERR_COMM;
  }
  return JP_OK;
}

int jp_comm_destroy(jp_comm* c) {
  if (!c) return JP_OK;
  cudaSetDevice(c->ctx->device);
  cudaStreamSynchronize(c->ctx->stream);
  for (int r = 0; r < c->world; ++r)
    if (c->ipc_opened[r]) cudaIpcCloseMemHandle(c->peer_h[r]);
  cudaFree(c->mailbox);
  cudaFree(c->d_peer);
  cudaFree(c->d_counter);
  delete c;
  return JP_OK;
}

}  // extern "C"
