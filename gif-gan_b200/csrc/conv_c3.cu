// Dedicated kernels for the image-side layers of the DCGAN: the 5x5 stride-2 SAME relation between a
// LARGE grid with 3 channels (the fp32 image, or its gradient) and a SMALL grid with a multiple of 64
// channels:   d_h0_conv (model.py:273: conv2d 3 -> 64) and g_h4 (model.py:321: deconv2d 64 -> 3).
// With K = 75 (or N = 3) these are not tensor-core GEMMs (arithmetic intensity ~31 flop/B);
// they are FFMA/shared-memory bound, so the kernels stage the image patch, the activation tile and the whole
// 19 KB filter in shared memory and keep 16 accumulators per thread fed by broadcast 128-bit filter loads.
//
//   c3_down : small[n,p,q,k] = act(sum_{r,s,c} large[n,2p+r-1,2q+s-1,c] w[r,s,c,k] + b[k])   (d_h0 fwd, g_h4 dgrad)
//   c3_up   : large[n,i,j,c] = act(sum_{r,s,k} small[n,p,q,k]  w[r,s,c,k] + b[c]),  i = 2p+r-1  (g_h4 fwd, d_h0 dgrad)
//   c3_wgrad: dw[r,s,c,k]   += sum_{n,p,q} large[n,2p+r-1,2q+s-1,c] small[n,p,q,k]            (both filter gradients)
#include <algorithm>

#include "common.cuh"

namespace gg {

constexpr int C3 = 3, KT = 5, NTAP = 25, RED = 75;   // channels, filter size, taps, taps*channels
constexpr int TS_ = 8;                                // 8x8 small-grid tile
constexpr int PATCH = 2 * TS_ + 3;                    // 19 large-grid rows/cols feed an 8x8 small tile
constexpr int PROW = PATCH * C3;                      // 57 floats per patch row

__device__ __forceinline__ void load_patch(float* sp, const float* __restrict__ large, int n, int p0, int q0, int H, int W, int tid, int nthreads) {
  // patch rows i = 2*p0 - 1 + a, cols j = 2*q0 - 1 + b, a,b in [0,19); zero outside the image (SAME padding)
  const int i0 = 2 * p0 - 1, j0 = 2 * q0 - 1;
  for (int e = tid; e < PATCH * PROW; e += nthreads) {
    const int a = e / PROW, rem = e - a * PROW, b = rem / C3, c = rem - b * C3;
    const int i = i0 + a, j = j0 + b;
    float v = 0.f;
    if (i >= 0 && i < H && j >= 0 && j < W) v = __ldg(large + (((int64_t)n * H + i) * W + j) * C3 + c);
    sp[e] = v;
  }
}

// ---------------------------------------------------------------------------------------------
template <typename TSM>
__global__ void __launch_bounds__(256)
c3_down_kernel(const float* __restrict__ large, const float* __restrict__ w, const float* __restrict__ bias, TSM* __restrict__ small,
               int H, int W, int Ho, int Wo, int K, int act, float act_param) {
  pdl_grid_sync();
  __shared__ __align__(16) float sw[RED * 64];
  __shared__ float sp[PATCH * PROW];
  const int tid = threadIdx.x;
  const int tiles_w = (Wo + TS_ - 1) / TS_;
  const int n = blockIdx.z, p0 = (blockIdx.x / tiles_w) * TS_, q0 = (blockIdx.x % tiles_w) * TS_;
  const int k0 = blockIdx.y * 64;
  for (int e = tid; e < RED * 64; e += 256) sw[e] = __ldg(w + (int64_t)(e >> 6) * K + k0 + (e & 63));
  load_patch(sp, large, n, p0, q0, H, W, tid, 256);
  __syncthreads();
  const int pix = tid & 63, kg = tid >> 6;
  const int py = pix >> 3, px = pix & 7;
  float acc[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) acc[j] = 0.f;
  const float* xp = sp + (2 * py) * PROW + (2 * px) * C3;
#pragma unroll 1
  for (int r = 0; r < KT; ++r) {
#pragma unroll
    for (int sc = 0; sc < KT * C3; ++sc) {
      const float xv = xp[r * PROW + sc];
      const float4* wr = reinterpret_cast<const float4*>(sw + (r * KT * C3 + sc) * 64 + kg * 16);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 wv = wr[q];
        acc[q * 4 + 0] = fmaf(xv, wv.x, acc[q * 4 + 0]);
        acc[q * 4 + 1] = fmaf(xv, wv.y, acc[q * 4 + 1]);
        acc[q * 4 + 2] = fmaf(xv, wv.z, acc[q * 4 + 2]);
        acc[q * 4 + 3] = fmaf(xv, wv.w, acc[q * 4 + 3]);
      }
    }
  }
  const int p = p0 + py, q = q0 + px;
  if (p < Ho && q < Wo) {
    TSM* dst = small + (((int64_t)n * Ho + p) * Wo + q) * K + k0 + kg * 16;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      float v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = acc[g * 4 + j] + (bias ? __ldg(bias + k0 + kg * 16 + g * 4 + j) : 0.f);
      act_fwd_vec<4>(v, act, act_param);
      st4(dst + g * 4, make_float4(v[0], v[1], v[2], v[3]));
    }
  }
}

// ---------------------------------------------------------------------------------------------
// c3_up: CTA = 16x16 tile of the large grid for one image; warp-uniform output parity class.
constexpr int UP_T = 16;                 // large tile edge
constexpr int UP_S = UP_T / 2 + 2;       // 10 small rows/cols: p in [i0/2 - 1, i0/2 + 8]
constexpr int UP_STRIDE = 68;            // floats per small pixel in smem (64 + 4: spreads 128-bit loads over banks)

template <typename TSM>
__global__ void __launch_bounds__(256)
c3_up_kernel(const TSM* __restrict__ small, const float* __restrict__ w, const float* __restrict__ bias, float* __restrict__ large,
             int H, int W, int Ho, int Wo, int K, int act, float act_param) {
  pdl_grid_sync();
  extern __shared__ __align__(16) float smem_up[];
  float* sw = smem_up;                         // [25][3][64] for the current 64-channel block
  float* sx = smem_up + NTAP * C3 * 64;        // [10*10][UP_STRIDE]
  const int tid = threadIdx.x;
  const int tiles_w = (W + UP_T - 1) / UP_T;
  const int n = blockIdx.z, i0 = (blockIdx.x / tiles_w) * UP_T, j0 = (blockIdx.x % tiles_w) * UP_T;
  const int pb = i0 / 2 - 1, qb = j0 / 2 - 1;  // small-grid origin of the staged patch
  // thread -> (class, position): warps 0-1 class (0,0), 2-3 (0,1), 4-5 (1,0), 6-7 (1,1)
  const int cls = tid >> 6, pos = tid & 63;
  const int ah = cls >> 1, aw = cls & 1;
  const int mi = pos >> 3, mj = pos & 7;
  const int i = i0 + 2 * mi + ah, j = j0 + 2 * mj + aw;
  float acc[C3] = {0.f, 0.f, 0.f};
  for (int kb = 0; kb < K; kb += 64) {
    __syncthreads();
    for (int e = tid; e < NTAP * C3 * 64; e += 256) sw[e] = __ldg(w + (int64_t)(e >> 6) * K + kb + (e & 63));
    for (int e = tid; e < UP_S * UP_S * 16; e += 256) {          // 16 groups of 4 channels per small pixel
      const int sp = e >> 4, c4 = (e & 15) * 4;
      const int p = pb + sp / UP_S, q = qb + sp % UP_S;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (p >= 0 && p < Ho && q >= 0 && q < Wo) v = ld4(small + (((int64_t)n * Ho + p) * Wo + q) * K + kb + c4);
      *reinterpret_cast<float4*>(sx + sp * UP_STRIDE + c4) = v;
    }
    __syncthreads();
    // taps of this parity class: r with (ah + 1 - r) even -> p = (i + 1 - r) / 2
#pragma unroll 1
    for (int r = (ah + 1) & 1; r < KT; r += 2) {
      const int pl = (2 * mi + ah + 1 - r) / 2 + 1;   // row inside the staged patch (pb = i0/2 - 1); exact division
#pragma unroll 1
      for (int s = (aw + 1) & 1; s < KT; s += 2) {
        const int ql = (2 * mj + aw + 1 - s) / 2 + 1;
        const float4* xr = reinterpret_cast<const float4*>(sx + (pl * UP_S + ql) * UP_STRIDE);
        const float4* w0 = reinterpret_cast<const float4*>(sw + ((r * KT + s) * C3 + 0) * 64);
        const float4* w1 = w0 + 16;
        const float4* w2 = w0 + 32;
#pragma unroll
        for (int k4 = 0; k4 < 16; ++k4) {
          const float4 x = xr[k4], a = w0[k4], b = w1[k4], c = w2[k4];
          acc[0] = fmaf(x.x, a.x, fmaf(x.y, a.y, fmaf(x.z, a.z, fmaf(x.w, a.w, acc[0]))));
          acc[1] = fmaf(x.x, b.x, fmaf(x.y, b.y, fmaf(x.z, b.z, fmaf(x.w, b.w, acc[1]))));
          acc[2] = fmaf(x.x, c.x, fmaf(x.y, c.y, fmaf(x.z, c.z, fmaf(x.w, c.w, acc[2]))));
        }
      }
    }
  }
  if (i < H && j < W) {
    float* dst = large + (((int64_t)n * H + i) * W + j) * C3;
#pragma unroll
    for (int c = 0; c < C3; ++c) acc[c] += (bias ? __ldg(bias + c) : 0.f);
    act_fwd_vec<C3>(acc, act, act_param);
#pragma unroll
    for (int c = 0; c < C3; ++c) dst[c] = acc[c];
  }
}

// ---------------------------------------------------------------------------------------------
// c3_wgrad: persistent CTAs loop over 8x8 small tiles; thread = (tap*channel t in [0,75), 16-channel group)
constexpr int WG_THREADS = 320;   // 4 groups x 80 (75 used)

template <typename TSM>
__global__ void __launch_bounds__(WG_THREADS)
c3_wgrad_kernel(const float* __restrict__ large, const TSM* __restrict__ small, float* __restrict__ dw, int N, int H, int W, int Ho, int Wo,
                int K, int kblocks) {
  pdl_grid_sync();
  __shared__ float sp[PATCH * PROW];
  __shared__ __align__(16) float sy[64 * 64];
  const int tid = threadIdx.x;
  const int kg = tid / 80, t = tid % 80;
  const bool active = t < RED;
  const int r = t / (KT * C3), sc = t % (KT * C3);
  const int tiles_w = (Wo + TS_ - 1) / TS_, tiles_h = (Ho + TS_ - 1) / TS_;
  const int64_t ntiles = (int64_t)N * tiles_h * tiles_w;
  const int kb = (blockIdx.x % kblocks) * 64;
  float acc[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) acc[j] = 0.f;
  for (int64_t tile = blockIdx.x / kblocks; tile < ntiles; tile += gridDim.x / kblocks) {
    const int n = (int)(tile / (tiles_h * tiles_w));
    const int rem = (int)(tile - (int64_t)n * tiles_h * tiles_w);
    const int p0 = (rem / tiles_w) * TS_, q0 = (rem % tiles_w) * TS_;
    __syncthreads();
    load_patch(sp, large, n, p0, q0, H, W, tid, WG_THREADS);
    for (int e = tid; e < 64 * 16; e += WG_THREADS) {
      const int pix = e >> 4, c4 = (e & 15) * 4;
      const int p = p0 + (pix >> 3), q = q0 + (pix & 7);
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (p < Ho && q < Wo) v = ld4(small + (((int64_t)n * Ho + p) * Wo + q) * K + kb + c4);
      *reinterpret_cast<float4*>(sy + pix * 64 + c4) = v;
    }
    __syncthreads();
    if (active) {
      const float* xp = sp + r * PROW + sc;
#pragma unroll 4
      for (int pix = 0; pix < 64; ++pix) {
        const float xv = xp[(2 * (pix >> 3)) * PROW + (2 * (pix & 7)) * C3];
        const float4* yr = reinterpret_cast<const float4*>(sy + pix * 64 + kg * 16);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 yv = yr[q];
          acc[q * 4 + 0] = fmaf(xv, yv.x, acc[q * 4 + 0]);
          acc[q * 4 + 1] = fmaf(xv, yv.y, acc[q * 4 + 1]);
          acc[q * 4 + 2] = fmaf(xv, yv.z, acc[q * 4 + 2]);
          acc[q * 4 + 3] = fmaf(xv, yv.w, acc[q * 4 + 3]);
        }
      }
    }
  }
  if (active) {
    float* dst = dw + (int64_t)t * K + kb + kg * 16;     // dw[(r*5+s)*3+c][k]: t enumerates (r,s,c) in that order
#pragma unroll
    for (int j = 0; j < 16; ++j) atomicAdd(dst + j, acc[j]);
  }
}

// ---------------------------------------------------------------------------------------------
bool c3_applicable(const gg_conv_desc* d) {
  return d->C == C3 && d->K % 64 == 0 && d->large_dtype == GG_F32 && d->D == 1 && d->Do == 1 && d->kd == 1 && d->kh == KT && d->kw == KT &&
         d->sd == 1 && d->sh == 2 && d->sw == 2 && d->pd == 0 && d->ph == 1 && d->pw == 1 && d->H % 2 == 0 && d->W % 2 == 0 &&
         d->Ho == d->H / 2 && d->Wo == d->W / 2;
}

int c3_conv_down(const gg_conv_desc* d, const float* large, const float* w, const float* bias, void* small, cudaStream_t st) {
  dim3 grid(ceil_div(d->Ho, TS_) * ceil_div(d->Wo, TS_), d->K / 64, d->N);
  if (d->small_dtype == GG_F32)
    Launch(grid, 256, 0, st)(c3_down_kernel<float>, large, w, bias, (float*)small, d->H, d->W, d->Ho, d->Wo, d->K, d->act, d->act_param);
  else
    Launch(grid, 256, 0, st)(c3_down_kernel<bf16>, large, w, bias, (bf16*)small, d->H, d->W, d->Ho, d->Wo, d->K, d->act, d->act_param);
  return check_launch("c3_down");
}

int c3_conv_up(const gg_conv_desc* d, const void* small, const float* w, const float* bias, float* large, cudaStream_t st) {
  dim3 grid(ceil_div(d->H, UP_T) * ceil_div(d->W, UP_T), 1, d->N);
  const size_t smem = (NTAP * C3 * 64 + UP_S * UP_S * UP_STRIDE) * sizeof(float);
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(c3_up_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(c3_up_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr_done = true;
  }
  if (d->small_dtype == GG_F32)
    Launch(grid, 256, smem, st)(c3_up_kernel<float>, (const float*)small, w, bias, large, d->H, d->W, d->Ho, d->Wo, d->K, d->act, d->act_param);
  else
    Launch(grid, 256, smem, st)(c3_up_kernel<bf16>, (const bf16*)small, w, bias, large, d->H, d->W, d->Ho, d->Wo, d->K, d->act, d->act_param);
  return check_launch("c3_up");
}

int c3_conv_wgrad(const gg_conv_desc* d, const float* large, const void* small, float* dw, cudaStream_t st) {
  const int kblocks = d->K / 64;
  const int64_t ntiles = (int64_t)d->N * ceil_div(d->Ho, TS_) * ceil_div(d->Wo, TS_);
  const int per_k = (int)std::max<int64_t>(1, std::min<int64_t>(ntiles, (148 * 3) / kblocks));
  const int grid = per_k * kblocks;
  if (d->small_dtype == GG_F32)
    Launch(grid, WG_THREADS, 0, st)(c3_wgrad_kernel<float>, large, (const float*)small, dw, d->N, d->H, d->W, d->Ho, d->Wo, d->K, kblocks);
  else
    Launch(grid, WG_THREADS, 0, st)(c3_wgrad_kernel<bf16>, large, (const bf16*)small, dw, d->N, d->H, d->W, d->Ho, d->Wo, d->K, kblocks);
  return check_launch("c3_wgrad");
}

}  // namespace gg
