// Data-parallel gradient exchange over NVLink 5 / NVSwitch peer memory: ONE kernel per optimiser update that sum-all-reduces
// a range of the flat fp32 gradient buffer in place on every rank of one box (2..8 GPUs, one process each).
//
// Why not ncclAllReduce here: the train step is ~1.5 ms and exchanges three ranges of 17-20 MB; measured on 2 x B200
// (profiles/r4d_nccl_knobs.log) the NCCL ring costs ~48 us per range fully exposed (~200 GB/s), limiting its channel count
// makes it proportionally slower (the time IS transfer time), and overlapping it with the backward pass costs more than it
// hides (its CTAs displace the one-wave tcgen05 launches).  NVSwitch gives every GPU 900 GB/s to every peer, so a two-shot
// exchange by plain peer loads needs 2 * (N-1)/N * bytes / ~0.7 TB/s: ~25 us for 17 MB at N = 2, ~45 us at N = 8.
//
// Algorithm (the gradient buffers of all ranks are mapped into every process: gg_ipc_export / gg_ipc_import):
//   barrier 0   every rank has entered the kernel, i.e. its gradient kernels have completed (stream order on that rank);
//   phase 1     reduce-scatter by pull: rank r sums chunk r over the ranks IN RANK ORDER (the result does not depend on
//               timing, and every rank ends up with bit-identical gradients) and writes it into its own gradient buffer
//               AND into its (peer-mapped) staging buffer;
//   barrier 1   (after a system-scope fence) every chunk is reduced -- and nobody reads a peer's GRADIENTS any more;
//   phase 2     all-gather by pull: rank r copies chunk p from rank p's staging buffer, for every p != r.
// There is no closing barrier: a rank's staging buffer is next written in phase 1 of the NEXT launch, i.e. after that launch's
// barrier 0, which a peer only reaches once it has left this launch; and the gradient buffer may be rewritten by the rank's
// next kernels (the following update's zero fill) as soon as this kernel ends, because peers read it before barrier 1 only.
// Barriers are per thread block: block b of every rank only ever reads what block b of a peer wrote (same grid, same
// grid-stride element mapping on every rank), so block b synchronises with the peers' block b alone -- no grid-wide barrier.
// A flag is the launch sequence number (monotonic, kept in the signal buffer): replays of a captured CUDA graph need no
// host-side state.  Spins are bounded: a protocol bug or a missing peer traps (launch error) instead of hanging the box.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <mutex>
#include <string>

#include "common.cuh"

namespace gg {

constexpr int DP_MAX_RANKS = 8;
constexpr int DP_BLOCKS = 128;          // capacity of the signal buffer; the launch uses GG_DP_BLOCKS (default 64) of them
constexpr int DP_THREADS = 512;

// signal buffer of one rank (zero-filled by the caller): counter[b] = launches block b has completed;
// flag[k][b][src] = sequence number last written by rank src's block b at barrier k
struct DpSignals {
  unsigned int counter[DP_BLOCKS];
  unsigned int flag[2][DP_BLOCKS][DP_MAX_RANKS];
  long long prof[8];                  // block 0 of the last launch: cycles to [barrier 0 passed, phase 1 done, barrier 1 passed, phase 2 done]
};

struct DpPeers {
  float* grads[DP_MAX_RANKS];         // the same flat gradient buffer on every rank (peer-mapped)
  float* stage[DP_MAX_RANKS];         // staging of the chunk a rank reduced: ceil(numel / 4 / world) float4 (peer-mapped)
  DpSignals* sig[DP_MAX_RANKS];
};

__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// peer data is read exactly once per launch and must come from the owner's memory, never from a stale local line
__device__ __forceinline__ float4 ld_peer4(const float4* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float ld_peer1(const float* p) {
  float v;
  asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ void st_relaxed_sys(unsigned int* p, unsigned int v) {
  asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint2 ld_peer2(const uint2* p) {
  uint2 v;
  asm volatile("ld.relaxed.sys.global.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint2 pack_bf16x4(float4 v) {
  __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
  return make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
}
__device__ __forceinline__ float4 unpack_bf16x4(uint2 u) {
  return make_float4(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xffff0000u), __uint_as_float(u.y << 16), __uint_as_float(u.y & 0xffff0000u));
}
__device__ __forceinline__ unsigned int ld_relaxed_sys(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// All threads of the block.  The block's earlier global writes are ordered before the flag by CTA barrier + release store
// (release is cumulative over what the storing thread observed through bar.sync): no per-thread system fence -- 512 threads
// x 64 blocks of membar.sys made a barrier cost ~9 us (profiles/r4f_*).  The spin polls with relaxed loads (the flag lives in
// this rank's own memory; peers push into it) and acquires once at the end.
// `release` = false for the opening barrier: nothing written by THIS kernel has to be published (the gradients were written by
// earlier kernels, complete at the kernel boundary), and a system-scope release at kernel entry cost ~6 us (profiles/r4g_*).
__device__ __forceinline__ void dp_barrier(const DpPeers& P, int k, int rank, int world, unsigned int seq, bool release) {
  __syncthreads();
  if ((int)threadIdx.x < world) {
    const int t = threadIdx.x;
    if (release) st_release_sys(&P.sig[t]->flag[k][blockIdx.x][rank], seq);
    else st_relaxed_sys(&P.sig[t]->flag[k][blockIdx.x][rank], seq);
    const unsigned int* mine = &P.sig[rank]->flag[k][blockIdx.x][t];
    if ((int)(ld_relaxed_sys(mine) - seq) < 0) {
      const long long t0 = clock64();
      while ((int)(ld_relaxed_sys(mine) - seq) < 0) {
        if (clock64() - t0 > 6000000000ll) __trap();      // ~3 s: a peer never arrived
      }
    }
    (void)ld_acquire_sys(mine);
  }
  __syncthreads();
}

// WIRE16: the reduced chunks cross NVLink as bf16 in phase 2 (half the all-gather bytes); every rank -- the owner of a chunk
// included -- then holds the SAME bf16-rounded sums, so the replicas stay bit-identical.  The sums themselves are fp32 over
// the ranks' fp32 gradients (phase 1 is unchanged).
template <int WORLD, bool WIRE16>
__global__ void __launch_bounds__(DP_THREADS, 1)
dp_allreduce_kernel(const __grid_constant__ DpPeers P, int rank, long long lo, long long n) {
  pdl_grid_sync();                                   // this rank's gradient kernels have completed and flushed
  const long long t_in = clock64();
  const unsigned int seq = P.sig[rank]->counter[blockIdx.x] + 1u;
  dp_barrier(P, 0, rank, WORLD, seq, false);
  const long long t_b0 = clock64();
  // float4 units; `lo` is 16-byte aligned (optimiser groups are); the (n % 4)-element tail is summed by every rank itself
  const long long nv = n >> 2;
  const long long per = (nv + WORLD - 1) / WORLD;
  const long long stride = (long long)gridDim.x * DP_THREADS;
  const long long first = (long long)blockIdx.x * DP_THREADS + threadIdx.x;
  float tail = 0.f;
  const bool has_tail = blockIdx.x == 0 && (int)threadIdx.x < (int)(n & 3);
  {
    const long long c0 = min(nv, (long long)rank * per), c1 = min(nv, c0 + per);
    const float4* src[WORLD];
#pragma unroll
    for (int p = 0; p < WORLD; ++p) src[p] = reinterpret_cast<const float4*>(P.grads[p] + lo) + c0;
    float4* dst = reinterpret_cast<float4*>(P.grads[rank] + lo) + c0;
    float4* stg = reinterpret_cast<float4*>(P.stage[rank]);
    const long long len = c1 - c0;
    long long i = first;
    constexpr int U = WORLD <= 2 ? 4 : (WORLD == 3 ? 3 : 2);      // elements per trip: >= 8 independent 16-byte loads in flight per thread
    for (; i + (U - 1) * stride < len; i += U * stride) {
      float4 a[U][WORLD];
#pragma unroll
      for (int u = 0; u < U; ++u)
#pragma unroll
        for (int p = 0; p < WORLD; ++p) a[u][p] = ld_peer4(src[p] + i + u * stride);
#pragma unroll
      for (int u = 0; u < U; ++u) {
        float4 sa = a[u][0];
#pragma unroll
        for (int p = 1; p < WORLD; ++p) { sa.x += a[u][p].x; sa.y += a[u][p].y; sa.z += a[u][p].z; sa.w += a[u][p].w; }
        if (WIRE16) {
          const uint2 h = pack_bf16x4(sa);
          reinterpret_cast<uint2*>(stg)[i + u * stride] = h;
          dst[i + u * stride] = unpack_bf16x4(h);
        } else {
          dst[i + u * stride] = sa;
          stg[i + u * stride] = sa;
        }
      }
    }
    for (; i < len; i += stride) {
      float4 s = ld_peer4(src[0] + i);
#pragma unroll
      for (int p = 1; p < WORLD; ++p) { const float4 a = ld_peer4(src[p] + i); s.x += a.x; s.y += a.y; s.z += a.z; s.w += a.w; }
      if (WIRE16) {
        const uint2 h = pack_bf16x4(s);
        reinterpret_cast<uint2*>(stg)[i] = h;
        dst[i] = unpack_bf16x4(h);
      } else {
        dst[i] = s;
        stg[i] = s;
      }
    }
    if (has_tail) {                                    // every rank: the same rank-order sum
      const long long e = lo + (nv << 2) + threadIdx.x;
      tail = ld_peer1(P.grads[0] + e);
#pragma unroll
      for (int p = 1; p < WORLD; ++p) tail += ld_peer1(P.grads[p] + e);
      if (WIRE16) tail = __bfloat162float(__float2bfloat16_rn(tail));
    }
  }
  const long long t_p1 = clock64();
  dp_barrier(P, 1, rank, WORLD, seq, true);
  const long long t_b1 = clock64();
  if (has_tail) P.grads[rank][lo + (nv << 2) + threadIdx.x] = tail;      // peers have read the raw value (barrier 1)
  if (!WIRE16) {
    // every trip reads element i of ALL the other ranks' chunks (WORLD - 1 independent loads to WORLD - 1 different GPUs), two
    // elements per trip: a rank that walked its peers one after the other got 430 GB/s at N = 8 (profiles/r4h_*)
    const float4* src[WORLD];
    float4* dst[WORLD];
    long long len[WORLD];
#pragma unroll
    for (int q = 1; q < WORLD; ++q) {
      const int p = (rank + q) % WORLD;
      const long long c0 = min(nv, (long long)p * per), c1 = min(nv, c0 + per);
      src[q] = reinterpret_cast<const float4*>(P.stage[p]);
      dst[q] = reinterpret_cast<float4*>(P.grads[rank] + lo) + c0;
      len[q] = c1 - c0;
    }
    long long common = len[1];
#pragma unroll
    for (int q = 2; q < WORLD; ++q) common = min(common, len[q]);
    long long i = first;
    constexpr int U2 = WORLD <= 2 ? 4 : (WORLD <= 4 ? 2 : 1);       // elements per trip: 4-7 independent peer loads in flight per thread
    for (; i + (U2 - 1) * stride < common; i += U2 * stride) {
      float4 a[U2][WORLD];
#pragma unroll
      for (int u = 0; u < U2; ++u)
#pragma unroll
        for (int q = 1; q < WORLD; ++q) a[u][q] = ld_peer4(src[q] + i + u * stride);
#pragma unroll
      for (int u = 0; u < U2; ++u)
#pragma unroll
        for (int q = 1; q < WORLD; ++q) dst[q][i + u * stride] = a[u][q];
    }
#pragma unroll
    for (int q = 1; q < WORLD; ++q)
      for (long long k = i; k < len[q]; k += stride) dst[q][k] = ld_peer4(src[q] + k);
  } else {
#pragma unroll 1
    for (int q = 1; q < WORLD; ++q) {
      const int p = (rank + q) % WORLD;                  // every rank starts with a different peer
      const long long c0 = min(nv, (long long)p * per), c1 = min(nv, c0 + per);
      float4* dst = reinterpret_cast<float4*>(P.grads[rank] + lo) + c0;
      const long long len = c1 - c0;
      long long i = first;
      const uint2* src16 = reinterpret_cast<const uint2*>(P.stage[p]);
      for (; i + 7 * stride < len; i += 8 * stride) {
        uint2 h[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) h[u] = ld_peer2(src16 + i + u * stride);
#pragma unroll
        for (int u = 0; u < 8; ++u) dst[i + u * stride] = unpack_bf16x4(h[u]);
      }
      for (; i < len; i += stride) dst[i] = unpack_bf16x4(ld_peer2(src16 + i));
    }
  }
  if (threadIdx.x == 0) {
    P.sig[rank]->counter[blockIdx.x] = seq;
    if (blockIdx.x == 0) {
      long long* pr = P.sig[rank]->prof;
      pr[0] = t_b0 - t_in; pr[1] = t_p1 - t_in; pr[2] = t_b1 - t_in; pr[3] = clock64() - t_in;
    }
  }
}

// ---- CUDA IPC (legacy handles: the allocation a pointer lies in, plus the offset of the pointer inside it) ----
typedef CUresult (*GetRangeFn)(CUdeviceptr*, size_t*, CUdeviceptr);
static GetRangeFn get_range_fn() {
  static GetRangeFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess) fn = (GetRangeFn)p;
  });
  return fn;
}

int dp_ipc_export(const void* dev_ptr, void* handle64, uint64_t* offset) {
  GetRangeFn fn = get_range_fn();
  GG_REQUIRE(fn != nullptr, GG_ERR_CUDA, "gg_ipc_export: cuMemGetAddressRange entry point not available");
  CUdeviceptr base = 0;
  size_t size = 0;
  CUresult r = fn(&base, &size, (CUdeviceptr)dev_ptr);
  GG_REQUIRE(r == CUDA_SUCCESS, GG_ERR_CUDA, "gg_ipc_export: cuMemGetAddressRange failed (%d)", (int)r);
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, (void*)base);
  if (e != cudaSuccess) (void)cudaGetLastError();
  GG_REQUIRE(e == cudaSuccess, GG_ERR_CUDA, "gg_ipc_export: cudaIpcGetMemHandle failed: %s (allocations of an expandable-segments / VMM pool cannot be exported)",
             cudaGetErrorString(e));
  static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
  memcpy(handle64, &h, 64);
  *offset = (uint64_t)((CUdeviceptr)dev_ptr - base);
  return GG_OK;
}

int dp_ipc_import(const void* handle64, uint64_t offset, void** dev_ptr) {
  static std::map<std::string, void*> opened;      // an allocation may be opened once per process
  static std::mutex mu;
  std::lock_guard<std::mutex> lock(mu);
  const std::string key((const char*)handle64, 64);
  auto it = opened.find(key);
  void* base = nullptr;
  if (it != opened.end()) {
    base = it->second;
  } else {
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    cudaError_t e = cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) (void)cudaGetLastError();
    GG_REQUIRE(e == cudaSuccess, GG_ERR_CUDA, "gg_ipc_import: cudaIpcOpenMemHandle failed: %s", cudaGetErrorString(e));
    opened[key] = base;
  }
  *dev_ptr = (char*)base + offset;
  return GG_OK;
}

size_t dp_signal_bytes() { return sizeof(DpSignals); }

int dp_allreduce(void* const* grads, void* const* stage, void* const* signals, int rank, int world, int64_t lo, int64_t n, int wire_bf16,
                 cudaStream_t st) {
  GG_REQUIRE(world >= 2 && world <= DP_MAX_RANKS && rank >= 0 && rank < world, GG_ERR_INVALID, "gg_dp_allreduce: world size %d (2..8), rank %d", world, rank);
  GG_REQUIRE(n > 0 && lo >= 0, GG_ERR_INVALID, "gg_dp_allreduce: empty range");
  DpPeers P;
  memset(&P, 0, sizeof(P));
  for (int p = 0; p < world; ++p) {
    GG_REQUIRE(grads[p] != nullptr && stage[p] != nullptr && signals[p] != nullptr, GG_ERR_INVALID, "gg_dp_allreduce: rank %d is not mapped", p);
    P.grads[p] = (float*)grads[p];
    P.stage[p] = (float*)stage[p];
    P.sig[p] = (DpSignals*)signals[p];
  }
  GG_REQUIRE(((uintptr_t)(P.grads[rank] + lo) % 16) == 0 && ((uintptr_t)P.stage[rank] % 16) == 0, GG_ERR_INVALID,
             "gg_dp_allreduce: the range and the staging buffer must start 16-byte aligned");
  static const int nblocks = [] {
    const char* v = getenv("GG_DP_BLOCKS");
    const int b = (v && *v) ? atoi(v) : 64;
    return b < 1 ? 1 : (b > DP_BLOCKS ? DP_BLOCKS : b);
  }();
#define GG_DP(W)                                                                                                          \
  case W:                                                                                                                 \
    if (wire_bf16) Launch(nblocks, DP_THREADS, 0, st)(dp_allreduce_kernel<W, true>, P, rank, (long long)lo, (long long)n);     \
    else Launch(nblocks, DP_THREADS, 0, st)(dp_allreduce_kernel<W, false>, P, rank, (long long)lo, (long long)n);              \
    break
  switch (world) {
    GG_DP(2); GG_DP(3); GG_DP(4); GG_DP(5); GG_DP(6); GG_DP(7); GG_DP(8);
  }
#undef GG_DP
  return check_launch("dp_allreduce");
}

}  // namespace gg
