// Shared helpers for libgifgan.so (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "../../include/gifgan.h"

namespace gg {

void set_error(const char* fmt, ...);
extern std::atomic<uint64_t> g_launches;

inline int check_launch(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    set_error("%s: launch failed: %s", what, cudaGetErrorString(e));
    (void)cudaGetLastError();
    return GG_ERR_CUDA;
  }
  return GG_OK;
}

// ---- programmatic dependent launch (PDL) -----------------------------------------------------------
// Every kernel of the library is launched with cudaLaunchAttributeProgrammaticStreamSerialization and begins with
// pdl_grid_sync(): the next kernel's CTAs may be scheduled (and run their prologue: barrier init, TMEM allocation,
// descriptor prefetch) while the previous kernel drains, but touch global memory only after the previous grid has
// completed and flushed.  Inside a captured CUDA graph the attribute becomes a programmatic dependency edge; the
// step has ~200 short kernels, so the launch-to-launch bubble is a first-order cost.  GG_PDL=0 disables it.
bool pdl_enabled();
// Experiment switch (GG_CARVEOUT=1): ask for the same shared-memory carveout (maximum shared memory) in every kernel
// so that SMs are not reconfigured when a tcgen05 kernel (200 KB of stages) follows a streaming kernel.  Measured on
// B200: no gain, so the driver default stays.
void prefer_max_shared(const void* kernel);
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_grid_sync() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
#endif

struct Launch {
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attr[2];
  // cluster_x > 1: thread-block cluster (cluster_x, 1, 1); grid.x must be a multiple of it
  Launch(dim3 grid, dim3 block, size_t smem, cudaStream_t st, int cluster_x = 1) {
    cfg = cudaLaunchConfig_t();
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    cfg.attrs = attr; cfg.numAttrs = 1;
    if (cluster_x > 1) {
      attr[1].id = cudaLaunchAttributeClusterDimension;
      attr[1].val.clusterDim.x = (unsigned)cluster_x; attr[1].val.clusterDim.y = 1; attr[1].val.clusterDim.z = 1;
      cfg.numAttrs = 2;
    }
  }
  template <typename... KArgs, typename... Args>
  void operator()(void (*kernel)(KArgs...), Args&&... args) {
    prefer_max_shared(reinterpret_cast<const void*>(kernel));
    (void)cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);   // errors surface in check_launch()
  }
};

#define GG_REQUIRE(cond, code, ...) \
  do {                              \
    if (!(cond)) {                  \
      gg::set_error(__VA_ARGS__);   \
      return (code);                \
    }                               \
  } while (0)

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

typedef __nv_bfloat16 bf16;

// ---- element load/store with conversion to/from fp32 ------------------------------
__device__ __forceinline__ float ldf(const float* p) { return __ldg(p); }
__device__ __forceinline__ float ldf(const bf16* p) { return __bfloat162float(*p); }
__device__ __forceinline__ void stf(float* p, float v) { *p = v; }
__device__ __forceinline__ void stf(bf16* p, float v) { *p = __float2bfloat16_rn(v); }

// 4 consecutive elements (address must be 16 B / 8 B aligned)
__device__ __forceinline__ float4 ld4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 ld4(const bf16* p) {
  uint2 u = __ldg(reinterpret_cast<const uint2*>(p));
  __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&u.x);
  __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&u.y);
  float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
  return make_float4(fa.x, fa.y, fb.x, fb.y);
}
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void st4(bf16* p, float4 v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
  uint2 u;
  u.x = *reinterpret_cast<uint32_t*>(&a);
  u.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = u;
}

// ---- activations (include/gifgan.h gg_act) ----------------------------------------
// NOTE: no `switch` on the runtime activation code inside per-element code: nvcc lowers it to an indirect
// jump-table branch (BRX) per element, which cost ~20k cycles per tile in the tcgen05 epilogue (profiles/).
// none / relu / lrelu share one branch-free form y = max(u, slope*u) with slope = 1 / 0 / leak; the
// transcendental ones sit behind a warp-uniform if-chain.
__device__ __forceinline__ float act_slope(int act, float a) { return act == GG_ACT_NONE ? 1.f : (act == GG_ACT_RELU ? 0.f : a); }

__device__ __forceinline__ float act_fwd(float u, int act, float a) {
  if (act <= GG_ACT_LRELU) return fmaxf(u, u * act_slope(act, a));
  if (act == GG_ACT_TANH) return tanhf(u);
  if (act == GG_ACT_SIGMOID) return 1.f / (1.f + expf(-u));
  return 0.5f * (tanhf(u) + 1.f);   // GG_ACT_TANH01
}
// derivative expressed through the OUTPUT y = act(u)
__device__ __forceinline__ float act_grad_from_out(float y, int act, float a) {
  if (act <= GG_ACT_LRELU) {
    // tf.nn.relu: 0 at 0;  tf.maximum(x, a*x): 1 at 0 (0 < a < 1);  identity: 1
    const float at_zero = act == GG_ACT_RELU ? 0.f : 1.f;
    return y > 0.f ? 1.f : (y == 0.f ? at_zero : act_slope(act, a));
  }
  if (act == GG_ACT_TANH) return 1.f - y * y;
  if (act == GG_ACT_SIGMOID) return y * (1.f - y);
  const float t = 2.f * y - 1.f;
  return 0.5f * (1.f - t * t);
}
// derivative expressed through the pre-activation u
__device__ __forceinline__ float act_grad_from_pre(float u, int act, float a) {
  if (act <= GG_ACT_LRELU) {
    const float at_zero = act == GG_ACT_RELU ? 0.f : 1.f;
    return u > 0.f ? 1.f : (u == 0.f ? at_zero : act_slope(act, a));
  }
  if (act == GG_ACT_TANH) { const float t = tanhf(u); return 1.f - t * t; }
  if (act == GG_ACT_SIGMOID) { const float s = 1.f / (1.f + expf(-u)); return s * (1.f - s); }
  const float t = tanhf(u);
  return 0.5f * (1.f - t * t);
}

// Vector forms: ONE warp-uniform branch per N elements, every loop fully unrolled.  (A `#pragma unroll 1` loop
// over the register array indexes it dynamically, which moves the whole array -- on EVERY path -- to local memory:
// ptxas reported a 128-byte stack frame for the tcgen05 epilogue until the slow paths were unrolled as well.)
template <int N>
__device__ __forceinline__ void act_fwd_vec(float (&x)[N], int act, float a) {
  if (act <= GG_ACT_LRELU) {
    const float slope = act_slope(act, a);
#pragma unroll
    for (int i = 0; i < N; ++i) x[i] = fmaxf(x[i], x[i] * slope);
  } else if (act == GG_ACT_TANH) {
#pragma unroll
    for (int i = 0; i < N; ++i) x[i] = tanhf(x[i]);
  } else if (act == GG_ACT_SIGMOID) {
#pragma unroll
    for (int i = 0; i < N; ++i) x[i] = 1.f / (1.f + expf(-x[i]));
  } else {
#pragma unroll
    for (int i = 0; i < N; ++i) x[i] = 0.5f * (tanhf(x[i]) + 1.f);
  }
}
// g[i] *= act'(u[i])  (derivative through the pre-activation)
template <int N>
__device__ __forceinline__ void act_bwd_pre_vec(float (&g)[N], const float (&u)[N], int act, float a) {
  if (act <= GG_ACT_LRELU) {
    const float slope = act_slope(act, a), at_zero = act == GG_ACT_RELU ? 0.f : 1.f;
#pragma unroll
    for (int i = 0; i < N; ++i) g[i] *= (u[i] > 0.f ? 1.f : (u[i] == 0.f ? at_zero : slope));
  } else if (act == GG_ACT_SIGMOID) {
#pragma unroll
    for (int i = 0; i < N; ++i) { const float s = 1.f / (1.f + expf(-u[i])); g[i] *= s * (1.f - s); }
  } else {
    const float sc = act == GG_ACT_TANH ? 1.f : 0.5f;
#pragma unroll
    for (int i = 0; i < N; ++i) { const float t = tanhf(u[i]); g[i] *= sc * (1.f - t * t); }
  }
}
// g[i] *= act'(.) evaluated from the output y[i]
template <int N>
__device__ __forceinline__ void act_bwd_out_vec(float (&g)[N], const float (&y)[N], int act, float a) {
  if (act <= GG_ACT_LRELU) {
    const float slope = act_slope(act, a), at_zero = act == GG_ACT_RELU ? 0.f : 1.f;
#pragma unroll
    for (int i = 0; i < N; ++i) g[i] *= (y[i] > 0.f ? 1.f : (y[i] == 0.f ? at_zero : slope));
  } else if (act == GG_ACT_TANH) {
#pragma unroll
    for (int i = 0; i < N; ++i) g[i] *= 1.f - y[i] * y[i];
  } else if (act == GG_ACT_SIGMOID) {
#pragma unroll
    for (int i = 0; i < N; ++i) g[i] *= y[i] * (1.f - y[i]);
  } else {
#pragma unroll
    for (int i = 0; i < N; ++i) { const float t = 2.f * y[i] - 1.f; g[i] *= 0.5f * (1.f - t * t); }
  }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Column sums of a 32 x 32 tile held one ROW per lane (a[j] = element (lane, j)): returns, in lane L, the sum over the 32
// lanes of a[L].  Butterfly transpose-reduce: each round a lane keeps the half of its values whose column bit equals its
// own lane bit and adds the partner's copy of that half -- 16 + 8 + 4 + 2 + 1 = 31 shuffles instead of 32 x 5.
// `a` is clobbered.
__device__ __forceinline__ float warp_transpose_sum32(float (&a)[32], int lane) {
#pragma unroll
  for (int half = 16; half >= 1; half >>= 1) {
    const bool up = (lane & half) != 0;
#pragma unroll
    for (int i = 0; i < half; ++i) {
      const float send = up ? a[i] : a[i + half];
      const float keep = up ? a[i + half] : a[i];
      a[i] = keep + __shfl_xor_sync(0xffffffffu, send, half);
    }
  }
  return a[0];
}

// ---- replicated fp64 batch-norm accumulators ---------------------------------------------------------
// Statistics / backward reductions are accumulated with one fp64 atomic per channel per CTA (300-500 CTAs per address).
// The accumulators can be replicated R times, sums[R][groups][2][C]: a CTA adds to replica (its index mod R), readers
// add the R copies.
// Measured on B200 (DCGAN-64 step, batch 64): replication did NOT pay -- 1.775 ms/step with R = 8/4 against 1.682 ms with
// R = 1 (the readers' extra L2 round trips and the larger zero-fill outweigh the shorter atomic tails), so R is 1; the
// layout and the helpers stay so that the experiment is one line away.
__host__ __device__ inline int bn_replicas(int C, int groups) {
  (void)C; (void)groups;
  return 1;
}
__host__ __device__ inline int64_t bn_sum_index(int rep, int groups, int grp, int which, int C, int c) {
  return ((((int64_t)rep * groups + grp) * 2 + which) * C) + c;
}
#ifdef __CUDACC__
__device__ __forceinline__ double bn_sum_read(const double* __restrict__ sums, int R, int groups, int grp, int which, int C, int c) {
  if (R == 1) return sums[bn_sum_index(0, groups, grp, which, C, c)];
  // all replica loads are issued before the first add (an `a += load` loop serialises R L2 round trips per channel)
  double v[8];
#pragma unroll
  for (int r = 0; r < 8; ++r) v[r] = (r < R) ? sums[bn_sum_index(r, groups, grp, which, C, c)] : 0.0;
  return ((v[0] + v[1]) + (v[2] + v[3])) + ((v[4] + v[5]) + (v[6] + v[7]));
}
#endif

inline size_t dtype_size(int dt) { return dt == GG_BF16 ? 2 : 4; }

// dispatch a lambda-like macro over (dtype) -> type
#define GG_DISPATCH_DTYPE(dt, T, ...)                   \
  do {                                                  \
    if ((dt) == GG_F32) {                               \
      typedef float T;                                  \
      __VA_ARGS__;                                      \
    } else {                                            \
      typedef gg::bf16 T;                               \
      __VA_ARGS__;                                      \
    }                                                   \
  } while (0)

}  // namespace gg
