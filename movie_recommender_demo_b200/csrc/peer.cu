// peer.cu — the shard-result exchange over NVLink PEER MEMORY (SURVEY.md §8e, §5 "peer-memory fused all-gather").
//
// The reference has no sharding; round 1 exchanged the per-rank top-k lists with NCCL all-gathers.  For the
// latency-bound batches (Q <= 256: 0.1-1 MB per rank) the collective's launch + protocol latency is what is left
// beside the shard scan, so this file replaces it with plain P2P stores:
//
//   every rank owns ONE cudaMalloc'd buffer, exported with cudaIpcGetMemHandle and mapped by every peer:
//       [control block 4 KB: ready[64], ack[64]] [receive area: P blocks of Q x (2k+1) packed int32 words]
//   b2r_peer_allgather   peer_push_kernel : waits until every peer has consumed the previous contents (ack flags
//                                           in MY control block, written remotely by the peers), copies my packed
//                                           list into slot [my rank] of EVERY rank's receive area (16-byte stores
//                                           over NVLink / NVSwitch), fences at system scope; the last block
//                                           publishes ready[my rank] = epoch in every rank's control block
//                        peer_wait_kernel : spins (bounded) until ready[r] >= epoch for all r
//   (the caller merges the receive area with b2r_topk_merge_packed, as after the NCCL all-gather)
//   b2r_peer_ack         peer_ack_kernel  : tells every peer that my receive area has been read
//
// The epoch counter lives in device memory and every pointer is fixed, so the whole step (local search -> pack ->
// push -> wait -> merge -> ack) can be captured once in a CUDA graph and replayed - no NCCL call inside.
#include <string.h>

#include "internal.h"
#include "ptx.cuh"

namespace {
constexpr int kPeerMax = 64;
constexpr size_t kPeerCtrlBytes = 4096;
}  // namespace

struct b2r_peer {
  int rank = 0, world = 1, device = 0;
  size_t cap_bytes = 0;            // receive area
  uint8_t* local = nullptr;        // control block + receive area (IPC exported)
  uint8_t* peers[kPeerMax] = {};   // mapped bases (peers[rank] == local)
  bool opened[kPeerMax] = {};
  uint32_t* epoch = nullptr;       // device: pushes completed by this rank
  uint32_t* counter = nullptr;     // device: block counter of the push kernel
};

namespace b2r {
namespace {

struct PeerPtrs {
  uint8_t* base[kPeerMax];
};

__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// bounded spin: a protocol bug (or a dead peer) traps instead of hanging the GPU
__device__ __forceinline__ void spin_until_ge(const uint32_t* p, uint32_t want, int tag) {
  long long t0 = 0;
  for (uint32_t it = 0;; ++it) {
    if ((int32_t)(ld_acquire_sys(p) - want) >= 0) return;
    if ((it & 1023u) == 1023u) {
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > (1ll << 32)) {   // ~2 s
        printf("b2r: peer exchange timeout tag=%d thread=%d want=%u have=%u\n", tag, (int)threadIdx.x, want, ld_acquire_sys(p));
        __trap();
      }
    }
  }
}

constexpr int kPushThreads = 256;

__global__ void __launch_bounds__(kPushThreads)
peer_push_kernel(PeerPtrs peers, int rank, int world, const int4* __restrict__ send, size_t vec16, size_t block_bytes,
                 uint32_t* __restrict__ epoch, uint32_t* __restrict__ counter) {
  const uint32_t e = *epoch;      // pushes completed so far; every block reads it before the last block bumps it
  uint32_t* ctrl = reinterpret_cast<uint32_t*>(peers.base[rank]);
  // 1. every peer must have read what my previous push left in ITS receive area
  if ((int)threadIdx.x < world) spin_until_ge(ctrl + kPeerMax + threadIdx.x, e, 1);
  __syncthreads();
  // 2. my packed list -> slot [rank] of every rank's receive area
  for (int p = 0; p < world; ++p) {
    int4* dst = reinterpret_cast<int4*>(peers.base[(rank + p) % world] + kPeerCtrlBytes + (size_t)rank * block_bytes);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < vec16; i += (size_t)gridDim.x * blockDim.x)
      dst[i] = send[i];
  }
  __threadfence_system();
  __syncthreads();
  // 3. the last block to finish publishes the new epoch in every rank's control block
  __shared__ bool last;
  if (threadIdx.x == 0) last = atomicAdd(counter, 1u) == gridDim.x - 1;
  __syncthreads();
  if (last) {
    __threadfence_system();
    if ((int)threadIdx.x < world)
      st_release_sys(reinterpret_cast<uint32_t*>(peers.base[threadIdx.x]) + rank, e + 1);
    if (threadIdx.x == 0) {
      *counter = 0;
      *epoch = e + 1;
    }
  }
}

__global__ void peer_wait_kernel(const uint32_t* __restrict__ ctrl, int world, const uint32_t* __restrict__ epoch) {
  const uint32_t e = *epoch;      // already bumped by the push kernel in front of this launch
  if ((int)threadIdx.x < world) spin_until_ge(ctrl + threadIdx.x, e, 2);
}

__global__ void peer_ack_kernel(PeerPtrs peers, int rank, int world, const uint32_t* __restrict__ epoch) {
  const uint32_t e = *epoch;
  __threadfence_system();
  if ((int)threadIdx.x < world)
    st_release_sys(reinterpret_cast<uint32_t*>(peers.base[threadIdx.x]) + kPeerMax + rank, e);
}

}  // namespace
}  // namespace b2r

using namespace b2r;

extern "C" {

int b2r_peer_destroy(b2r_peer* c) {
  if (!c) return B2R_OK;
  DeviceGuard guard(c->device);
  for (int p = 0; p < c->world; ++p)
    if (c->opened[p] && c->peers[p]) cudaIpcCloseMemHandle(c->peers[p]);
  cudaFree(c->local);
  cudaFree(c->epoch);
  cudaFree(c->counter);
  delete c;
  return B2R_OK;
}

int b2r_peer_create(b2r_peer** out, int rank, int world, size_t cap_bytes, int device) {
  if (!out) return fail(B2R_EINVAL, "peer_create: NULL argument");
  *out = nullptr;
  if (world < 1 || world > kPeerMax || rank < 0 || rank >= world) return fail(B2R_EINVAL, "peer_create: bad rank / world");
  DeviceGuard guard(device);
  if (!guard.ok) return fail(B2R_ECUDA, "peer_create: cudaSetDevice failed");
  b2r_peer* c = new b2r_peer();
  c->rank = rank;
  c->world = world;
  c->device = device;
  c->cap_bytes = align_up(cap_bytes, 256);
  if (cudaMalloc(&c->local, kPeerCtrlBytes + c->cap_bytes) != cudaSuccess || cudaMalloc(&c->epoch, 256) != cudaSuccess ||
      cudaMalloc(&c->counter, 256) != cudaSuccess) {
    cudaGetLastError();
    b2r_peer_destroy(c);
    return fail(B2R_ENOMEM, "peer_create: cudaMalloc failed");
  }
  cudaMemset(c->local, 0, kPeerCtrlBytes);
  cudaMemset(c->epoch, 0, 256);
  cudaMemset(c->counter, 0, 256);
  if (cudaDeviceSynchronize() != cudaSuccess) {
    b2r_peer_destroy(c);
    return fail(B2R_ECUDA, "peer_create: initialisation failed");
  }
  c->peers[rank] = c->local;
  *out = c;
  return B2R_OK;
}

/* 64-byte cudaIpcMemHandle of this rank's buffer (host memory). */
int b2r_peer_handle(b2r_peer* c, void* handle64) {
  if (!c || !handle64) return fail(B2R_EINVAL, "peer_handle: NULL argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is expected to be 64 bytes");
  DeviceGuard guard(c->device);
  cudaIpcMemHandle_t h;
  B2R_CUDA(cudaIpcGetMemHandle(&h, c->local));
  memcpy(handle64, &h, 64);
  return B2R_OK;
}

/* handles: [world][64] bytes, rank-major (the all-gathered b2r_peer_handle outputs). */
int b2r_peer_connect(b2r_peer* c, const void* handles) {
  if (!c || !handles) return fail(B2R_EINVAL, "peer_connect: NULL argument");
  DeviceGuard guard(c->device);
  for (int p = 0; p < c->world; ++p) {
    if (p == c->rank) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, reinterpret_cast<const uint8_t*>(handles) + (size_t)p * 64, 64);
    void* ptr = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      cudaGetLastError();
      return fail(B2R_ECUDA, std::string("peer_connect: cudaIpcOpenMemHandle of rank ") + std::to_string(p) + ": " +
                                 cudaGetErrorString(e));
    }
    c->peers[p] = reinterpret_cast<uint8_t*>(ptr);
    c->opened[p] = true;
  }
  return B2R_OK;
}

/* All-gather `bytes` (multiple of 16, 16-byte aligned) from every rank: on return (stream order) rank r's data sits at
 * *recv + r * bytes on every rank.  Collective: every rank calls it with the same `bytes`. */
int b2r_peer_allgather(b2r_peer* c, const void* send, size_t bytes, void** recv, void* stream_) {
  if (!c || !send || !recv) return fail(B2R_EINVAL, "peer_allgather: NULL argument");
  if (bytes % 16 != 0 || (reinterpret_cast<uintptr_t>(send) & 15) != 0) return fail(B2R_EINVAL, "peer_allgather: 16-byte granularity");
  if (bytes * (size_t)c->world > c->cap_bytes) return fail(B2R_ENOMEM, "peer_allgather: receive area too small");
  for (int p = 0; p < c->world; ++p)
    if (!c->peers[p]) return fail(B2R_ESTATE, "peer_allgather: not connected");
  cudaStream_t stream = (cudaStream_t)stream_;
  DeviceGuard guard(c->device);
  PeerPtrs pp;
  for (int p = 0; p < kPeerMax; ++p) pp.base[p] = p < c->world ? c->peers[p] : nullptr;
  const size_t vec16 = bytes / 16;
  int blocks = (int)ceil_div((int64_t)vec16, kPushThreads * 4);
  if (blocks < 1) blocks = 1;
  if (blocks > 64) blocks = 64;
  peer_push_kernel<<<blocks, kPushThreads, 0, stream>>>(pp, c->rank, c->world, reinterpret_cast<const int4*>(send), vec16,
                                                       bytes, c->epoch, c->counter);
  B2R_CHECK_LAUNCH("peer_push_kernel");
  peer_wait_kernel<<<1, 64, 0, stream>>>(reinterpret_cast<const uint32_t*>(c->local), c->world, c->epoch);
  B2R_CHECK_LAUNCH("peer_wait_kernel");
  *recv = c->local + kPeerCtrlBytes;
  return B2R_OK;
}

/* Tell the peers that this rank has finished reading its receive area (enqueue after the consumer kernel). */
int b2r_peer_ack(b2r_peer* c, void* stream_) {
  if (!c) return fail(B2R_EINVAL, "peer_ack: NULL argument");
  DeviceGuard guard(c->device);
  PeerPtrs pp;
  for (int p = 0; p < kPeerMax; ++p) pp.base[p] = p < c->world ? c->peers[p] : nullptr;
  peer_ack_kernel<<<1, 64, 0, (cudaStream_t)stream_>>>(pp, c->rank, c->world, c->epoch);
  B2R_CHECK_LAUNCH("peer_ack_kernel");
  return B2R_OK;
}

}  // extern "C"
