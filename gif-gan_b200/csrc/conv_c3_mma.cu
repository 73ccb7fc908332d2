// bf16-mode kernels for the image-side layers of the DCGAN (3 channels <-> 64k channels, 5x5, stride 2, SAME):
//   d_h0_conv (model.py:273) and g_h4 (model.py:321) -- forward, dgrad and wgrad.
// With C = 3 the reduction of the forward conv is K = 75 and the deconv has N = 3: there is no 64-channel row for a
// TMA box / tcgen05 K-chunk, and at ~31 flop/B the layers sit below the tensor ridge.  What bounded the fp32 SIMT
// versions (conv_c3.cu) was FFMA + shared-memory issue, not HBM, so the bf16 mode runs the same sums on warp-level
// tensor-core MMAs (mma.sync.m16n8k16, bf16 x bf16 -> fp32) with operands gathered from a staged image patch:
//   c3m_down : small[n,p,q,k] = act(sum_{r,s,c} large[n,2p+r-1,2q+s-1,c] w[r,s,c,k] + b[k])
//              GEMM  M = pixels, N = 64 channels, K = 5 kernel rows x 16 (15 used: (s,c) pairs + one zero column)
//   c3m_up   : large[n,2m+a,2l+b,c] = act(sum_{dp,dq,k} small[n,m+dp,l+dq,k] w[a+1-2dp, b+1-2dq, c, k] + bias[c])
//              GEMM  M = small-grid positions, N = 16 (4 output parities x 3 channels, 12 used), K = 9 neighbours x 64
//   c3m_wgrad: dw[r,s,c,k] += sum_{n,p,q} large[n,2p+r-1,2q+s-1,c] small[n,p,q,k]
//              GEMM  M = 15 blocks of 8 (kernel row, column pair, channel padded to 4), N = 64, K = pixels;
//              both operands are pixel-major in shared memory and enter through ldmatrix.trans.
// The large tensor (image / image gradient) is fp32 in HBM and rounded to bf16 while staging; accumulation is fp32.
#include <algorithm>

#include "common.cuh"

namespace gg {

namespace {

constexpr int KT5 = 5;

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  __nv_bfloat162 p = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&p);
}

// ------------------------------------------------------------------------------------------------
// c3m_down.  CTA = 8 warps; tile = 8 rows x 16 columns of the small grid; warp w owns row w (one m16 tile).
// K index kk = r*16 + sc with sc = s*3 + c (sc = 15 is a zero-weight column): one kernel row per k16 step, and the
// 16 values of a step are CONTIGUOUS in the staged patch row, so an A fragment is four 32-bit shared loads.
// Output-channel permutation: column e of n-tile j is channel (e>>1)*16 + 2j + (e&1), so that a thread ends up
// with 16 consecutive channels of a pixel (32 B of bf16) and a quad writes the pixel's whole 128-byte row.
constexpr int DN_TH = 8, DN_TW = 16;
constexpr int DN_PR = 2 * DN_TH + 3;            // 19 patch rows
constexpr int DN_PC = 2 * DN_TW + 3;            // 35 patch columns
constexpr int DN_PROW = DN_PC * 3 + 1;          // 106 bf16 per patch row (105 + the column read by the zero weight)
constexpr int DN_WROW = 88;                     // bf16 per channel row of the transposed filter (80 + 8: bank spread)

// Fused batch-norm BACKWARD reductions (BNB): when this launch is the dgrad of g_h4 (model.py:321), its output IS dy of g_bn3
// (train mode, ReLU): per channel sum g and sum g*xhat with g = bf16(dy) * act'(gamma*xhat + beta), xhat = (pre - mean) * rstd --
// what a separate colsum pass over (pre fp32, dy bf16) = 25 MB computed in ~20 us.  Here the thread that holds 16 channels
// of a pixel loads the matching 64 bytes of `pre` (a quad reads the pixel's whole 256-byte row), per-tile partials are
// reduced over the 8 pixel lanes by shuffles and meet in shared memory; ONE pair of fp64 atomics per channel per CTA.
struct C3Bnb {
  const float *pre, *mean, *rstd, *gamma, *beta;
  double* sums;               // [2][K], zeroed by the caller
  int act;
  float act_param;
};

template <typename TSM, bool BNB>
__global__ void __launch_bounds__(256, 2)
c3m_down_kernel(const float* __restrict__ large, const float* __restrict__ w, const float* __restrict__ bias, TSM* __restrict__ small,
                int N, int H, int W, int Ho, int Wo, int K, int kblocks, int act, float act_param, const C3Bnb bnb) {
  pdl_grid_sync();
  __shared__ float4 bconst[BNB ? 64 : 1];                   // {rstd, -mean*rstd, gamma, beta} of this CTA's 64 channels
  __shared__ float bsum[2][BNB ? 64 : 1];
  __shared__ __align__(16) bf16 swt[64 * DN_WROW];          // [channel][kk]
  __shared__ __align__(16) bf16 sp[DN_PR * DN_PROW];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int kb = (blockIdx.x % kblocks) * 64;
  const int tiles_w = (Wo + DN_TW - 1) / DN_TW, tiles_h = (Ho + DN_TH - 1) / DN_TH;
  const int ntiles = N * tiles_h * tiles_w;

  // ---- filter block -> shared (transposed, bf16, zero pad columns) -> B fragments in registers
  // (all 19 filter loads of a thread are issued before anything waits on one: as a load -> convert -> store loop the staging
  //  was a chain of 19 L2 round trips, ~half of the launch at batch 64)
  constexpr int NWL = (75 * 64 + 255) / 256;
  float wl[NWL];
#pragma unroll
  for (int k = 0; k < NWL; ++k) {
    const int e = tid + 256 * k;
    wl[k] = (e < 75 * 64) ? __ldg(w + (int64_t)(e >> 6) * K + kb + (e & 63)) : 0.f;
  }
  for (int e = tid; e < 64 * DN_WROW; e += 256) swt[e] = __float2bfloat16_rn(0.f);
  __syncthreads();
#pragma unroll
  for (int k = 0; k < NWL; ++k) {
    const int e = tid + 256 * k;
    if (e < 75 * 64) {
      const int tc = e >> 6, ch = e & 63;
      const int kk = (tc / 15) * 16 + (tc % 15);
      swt[ch * DN_WROW + kk] = __float2bfloat16_rn(wl[k]);
    }
  }
  for (int e = tid; e < DN_PR; e += 256) sp[e * DN_PROW + DN_PROW - 1] = __float2bfloat16_rn(0.f);
  __syncthreads();
  // B fragments in fragment order, swf[r][j][lane] = (b0, b1): one conflict-free 64-bit shared load per (kernel row,
  // n-tile) in the main loop.  (Holding all 80 fragment registers per thread capped the kernel at 8 warps per SM, and
  // ncu showed it latency-bound at 13 % warp occupancy.)
  __shared__ __align__(16) uint2 swf[KT5 * 8 * 32];
  {
    const uint32_t* swt32 = reinterpret_cast<const uint32_t*>(swt);
    for (int e = tid; e < KT5 * 8 * 32; e += 256) {
      const int ln = e & 31, j = (e >> 5) & 7, r = e >> 8;
      const int gg = ln >> 2, tt = ln & 3;
      const int ch = (gg >> 1) * 16 + 2 * j + (gg & 1);
      swf[e] = make_uint2(swt32[(ch * DN_WROW + r * 16 + 2 * tt) >> 1], swt32[(ch * DN_WROW + r * 16 + 2 * tt + 8) >> 1]);
    }
  }
  if (BNB && tid < 64) {
    const float mu = __ldg(bnb.mean + kb + tid), rs = __ldg(bnb.rstd + kb + tid);
    bconst[tid] = make_float4(rs, -mu * rs, bnb.gamma ? __ldg(bnb.gamma + kb + tid) : 1.f, bnb.beta ? __ldg(bnb.beta + kb + tid) : 0.f);
    bsum[0][tid] = 0.f;
    bsum[1][tid] = 0.f;
  }
  __syncthreads();
  float bv[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) bv[i] = (!BNB && bias) ? __ldg(bias + kb + t * 16 + i) : 0.f;     // (a dgrad launch has no bias: frees 16 registers for BNB)

  // Patch elements of this thread: e = tid + 256 k -> (row a, element x in the row, column x / 3), fixed for all tiles.
  constexpr int PK = (DN_PR * DN_PC * 3 + 255) / 256;           // 8 elements per thread
  int pdesc[PK];                                                // a | x << 8 | (x / 3) << 16, or -1
#pragma unroll
  for (int k = 0; k < PK; ++k) {
    const int e = tid + 256 * k;
    const int a = e / (DN_PC * 3), x = e - a * (DN_PC * 3);
    pdesc[k] = (e < DN_PR * DN_PC * 3) ? (a | (x << 8) | ((x / 3) << 16)) : -1;
  }
  float pv[PK];
  auto prefetch = [&](int tile) {                               // global -> registers (loads stay in flight)
    const int n = tile / (tiles_h * tiles_w);
    const int rem = tile - n * tiles_h * tiles_w;
    const int i0 = 2 * ((rem / tiles_w) * DN_TH) - 1, j0 = 2 * ((rem % tiles_w) * DN_TW) - 1;
    const float* img = large + (int64_t)n * H * W * 3;
#pragma unroll
    for (int k = 0; k < PK; ++k) {
      const int dsc = pdesc[k];
      const int i = i0 + (dsc & 0xff), j = j0 + (dsc >> 16);
      const bool ok = dsc >= 0 && i >= 0 && i < H && j >= 0 && j < W;
      pv[k] = ok ? __ldg(img + ((int64_t)i * W + j0) * 3 + ((dsc >> 8) & 0xff)) : 0.f;
    }
  };
  const int tstride = gridDim.x / kblocks;
  int tile = blockIdx.x / kblocks;
  if (tile < ntiles) prefetch(tile);
  for (; tile < ntiles; tile += tstride) {
    const int n = tile / (tiles_h * tiles_w);
    const int rem = tile - n * tiles_h * tiles_w;
    const int p0 = (rem / tiles_w) * DN_TH, q0 = (rem % tiles_w) * DN_TW;
    __syncthreads();                                  // previous tile's fragments are consumed
#pragma unroll
    for (int k = 0; k < PK; ++k)
      if (pdesc[k] >= 0) sp[(pdesc[k] & 0xff) * DN_PROW + ((pdesc[k] >> 8) & 0xff)] = __float2bfloat16_rn(pv[k]);
    __syncthreads();
    if (tile + tstride < ntiles) prefetch(tile + tstride);      // next patch streams in while this one is multiplied
    // BNB: this thread's 16 channels of the pre-norm tensor at its two pixels, requested BEFORE the MMAs (the epilogue would
    // otherwise expose one more memory round trip per tile: +18 us per launch measured)
    float xpre[BNB ? 2 : 1][BNB ? 16 : 1];
    if (BNB) {
      const int pp = p0 + warp;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int qq = q0 + g + 8 * half;
        const bool ok = pp < Ho && qq < Wo;
        const float4* xp = reinterpret_cast<const float4*>(bnb.pre + (((int64_t)n * Ho + (ok ? pp : 0)) * Wo + (ok ? qq : 0)) * K + kb + t * 16);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 u = ok ? __ldg(xp + i) : make_float4(0.f, 0.f, 0.f, 0.f);
          xpre[half][4 * i] = u.x; xpre[half][4 * i + 1] = u.y; xpre[half][4 * i + 2] = u.z; xpre[half][4 * i + 3] = u.w;
        }
      }
    }
    float acc[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) { acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f; }
    const uint32_t* sp32 = reinterpret_cast<const uint32_t*>(sp);
#pragma unroll
    for (int r = 0; r < KT5; ++r) {
      const uint32_t* row = sp32 + (2 * warp + r) * (DN_PROW / 2);
      uint32_t a[4];
      a[0] = row[3 * g + t];
      a[1] = row[3 * (g + 8) + t];
      a[2] = row[3 * g + t + 4];
      a[3] = row[3 * (g + 8) + t + 4];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint2 b = swf[(r * 8 + j) * 32 + lane];
        mma16816(acc[j], a, b.x, b.y);
      }
    }
    const int p = p0 + warp;
    float tg[BNB ? 16 : 1], tgx[BNB ? 16 : 1];                  // this tile's partial reductions of the thread's 16 channels
    if (BNB) {
#pragma unroll
      for (int i = 0; i < 16; ++i) { tg[i] = 0.f; tgx[i] = 0.f; }
    }
    if (p < Ho) {
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int q = q0 + g + 8 * half;
        if (q >= Wo) continue;
        float v[16];
#pragma unroll
        for (int j = 0; j < 8; ++j) { v[2 * j] = acc[j][2 * half] + bv[2 * j]; v[2 * j + 1] = acc[j][2 * half + 1] + bv[2 * j + 1]; }
        act_fwd_vec<16>(v, act, act_param);
        if (BNB) {
          const float* x = xpre[BNB ? half : 0];                // (register array: every index below is a compile-time constant)
          const float slope = act_slope(bnb.act, bnb.act_param), at_zero = bnb.act == GG_ACT_RELU ? 0.f : 1.f;
          const bool simple = bnb.act <= GG_ACT_LRELU;
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float4 k = bconst[t * 16 + i];
            const float xh = fmaf(x[i], k.x, k.y);
            const float u = fmaf(k.z, xh, k.w);
            float gv = sizeof(TSM) == 2 ? __bfloat162float(__float2bfloat16_rn(v[i])) : v[i];      // what the apply kernel reads back
            gv *= simple ? (u > 0.f ? 1.f : (u == 0.f ? at_zero : slope)) : act_grad_from_pre(u, bnb.act, bnb.act_param);
            tg[i] += gv;
            tgx[i] = fmaf(gv, xh, tgx[i]);
          }
        }
        TSM* dst = small + (((int64_t)n * Ho + p) * Wo + q) * K + kb + t * 16;
        if (sizeof(TSM) == 2) {                          // 16 bf16 = two 16-byte stores; the quad covers the pixel's 128-byte row
          uint4* d4 = reinterpret_cast<uint4*>(dst);
          d4[0] = make_uint4(pack2(v[0], v[1]), pack2(v[2], v[3]), pack2(v[4], v[5]), pack2(v[6], v[7]));
          d4[1] = make_uint4(pack2(v[8], v[9]), pack2(v[10], v[11]), pack2(v[12], v[13]), pack2(v[14], v[15]));
        } else {
#pragma unroll
          for (int i = 0; i < 16; i += 4) st4(dst + i, make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]));
        }
      }
    }
    if (BNB) {       // (all 32 lanes: rows beyond Ho contribute zeros) reduce over the 8 pixel lanes g (lane bits 2..4), then shared atomics
#pragma unroll
      for (int i = 0; i < 16; ++i) {
#pragma unroll
        for (int m = 4; m <= 16; m <<= 1) {
          tg[i] += __shfl_xor_sync(0xffffffffu, tg[i], m);
          tgx[i] += __shfl_xor_sync(0xffffffffu, tgx[i], m);
        }
      }
      if (g == 0) {
#pragma unroll
        for (int i = 0; i < 16; ++i) { atomicAdd(&bsum[0][t * 16 + i], tg[i]); atomicAdd(&bsum[1][t * 16 + i], tgx[i]); }
      }
    }
  }
  if (BNB) {
    __syncthreads();
    if (tid < 128) {
      const int which = tid >> 6, ch = tid & 63;
      atomicAdd(bnb.sums + bn_sum_index(0, 1, 0, which, K, kb + ch), (double)bsum[which][ch]);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// c3m_up.  CTA = 8 warps; tile = 8 rows x 16 columns of small-grid POSITIONS (= 16 x 32 output pixels);
// warp w owns position row w.  n = (a*2 + b)*3 + c  (a, b: output row / column parity).
constexpr int UP_TH = 8, UP_TW = 16;
constexpr int UP_PR = UP_TH + 2, UP_PC = UP_TW + 2;       // 10 x 18 staged small pixels (halo 1)
constexpr int UP_PIX = 72;                                 // bf16 per staged pixel (64 + 8: conflict-free ldmatrix rows)
constexpr int UP_WROW = 9 * 64 + 8;                        // 584 bf16 per n row of the expanded filter

template <typename TSM>
__global__ void __launch_bounds__(256, 3)
c3m_up_kernel(const TSM* __restrict__ small, const float* __restrict__ w, const float* __restrict__ bias, float* __restrict__ large,
              int N, int H, int W, int Ho, int Wo, int K, int act, float act_param) {
  pdl_grid_sync();
  __shared__ __align__(16) bf16 sx[UP_PR * UP_PC * UP_PIX];     // 25,920 B
  __shared__ __align__(16) bf16 swu[16 * UP_WROW];              // 18,688 B
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int tiles_w = (Wo + UP_TW - 1) / UP_TW, tiles_h = (Ho + UP_TH - 1) / UP_TH;
  const int ntiles = N * tiles_h * tiles_w;
  // Work items = (tile, 64-channel block); the next item's staged pixels are loaded into registers while the current
  // item is multiplied (the kernel was latency-bound on these loads).
  constexpr int PXC = (UP_PR * UP_PC * 8 + 255) / 256;          // 6 16-byte chunks per thread
  int pdesc[PXC];                                               // patch row | patch col << 8 | chunk << 16, -1 = none
#pragma unroll
  for (int k = 0; k < PXC; ++k) {
    const int e = tid + 256 * k, pix = e >> 3;
    pdesc[k] = (e < UP_PR * UP_PC * 8) ? ((pix / UP_PC) | ((pix % UP_PC) << 8) | ((e & 7) << 16)) : -1;
  }
  uint4 xv[PXC];
  auto prefetch = [&](int tile, int kb) {
    const int n = tile / (tiles_h * tiles_w);
    const int rem = tile - n * tiles_h * tiles_w;
    const int p0 = (rem / tiles_w) * UP_TH, q0 = (rem % tiles_w) * UP_TW;
#pragma unroll
    for (int k = 0; k < PXC; ++k) {
      const int p = p0 - 1 + (pdesc[k] & 0xff), q = q0 - 1 + ((pdesc[k] >> 8) & 0xff), c8 = (pdesc[k] >> 16) * 8;
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if (pdesc[k] >= 0 && p >= 0 && p < Ho && q >= 0 && q < Wo) {
        const TSM* src = small + (((int64_t)n * Ho + p) * Wo + q) * K + kb + c8;
        if (sizeof(TSM) == 2) {
          v = __ldg(reinterpret_cast<const uint4*>(src));
        } else {
          const float4 f0 = ld4(reinterpret_cast<const float*>(src)), f1 = ld4(reinterpret_cast<const float*>(src) + 4);
          v = make_uint4(pack2(f0.x, f0.y), pack2(f0.z, f0.w), pack2(f1.x, f1.y), pack2(f1.z, f1.w));
        }
      }
      xv[k] = v;
    }
  };
  if ((int)blockIdx.x < ntiles) prefetch(blockIdx.x, 0);
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int n = tile / (tiles_h * tiles_w);
    const int rem = tile - n * tiles_h * tiles_w;
    const int p0 = (rem / tiles_w) * UP_TH, q0 = (rem % tiles_w) * UP_TW;
    float acc[2][4];
#pragma unroll
    for (int j = 0; j < 2; ++j) { acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f; }
    for (int kb = 0; kb < K; kb += 64) {
      __syncthreads();
      // expanded filter of this channel block: swu[n][(dp+1)*3 + (dq+1)][k] = w[a+1-2dp][b+1-2dq][c][kb+k] (0 if outside 5x5)
      if (kb > 0 || tile == (int)blockIdx.x || K > 64) {
        // 9 items of 4 channels per thread, loads batched 3 at a time (three L2 round trips instead of nine; the whole batch of
        // nine does not fit the 80-register budget of three co-resident CTAs without spilling)
#pragma unroll 1
        for (int k0 = 0; k0 < 9; k0 += 3) {
          float4 wu[3];
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            const int e = tid + 256 * (k0 + k);
            const int k4 = (e & 15) * 4, nb = (e >> 4) % 9, nn = e / (16 * 9);
            const int cls = nn / 3, c = nn - cls * 3, a = cls >> 1, b = cls & 1;
            const int r = a + 1 - 2 * (nb / 3 - 1), s = b + 1 - 2 * (nb % 3 - 1);
            wu[k] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (nn < 12 && r >= 0 && r < KT5 && s >= 0 && s < KT5) wu[k] = ld4(w + ((int64_t)((r * KT5 + s) * 3 + c)) * K + kb + k4);
          }
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            const int e = tid + 256 * (k0 + k);
            const int k4 = (e & 15) * 4, nb = (e >> 4) % 9, nn = e / (16 * 9);
            st4(swu + nn * UP_WROW + nb * 64 + k4, wu[k]);
          }
        }
      }
#pragma unroll
      for (int k = 0; k < PXC; ++k)
        if (pdesc[k] >= 0)
          *reinterpret_cast<uint4*>(sx + ((pdesc[k] & 0xff) * UP_PC + ((pdesc[k] >> 8) & 0xff)) * UP_PIX + (pdesc[k] >> 16) * 8) = xv[k];
      __syncthreads();
      if (kb + 64 < K) prefetch(tile, kb + 64);
      else if (tile + (int)gridDim.x < ntiles) prefetch(tile + gridDim.x, 0);
      const uint32_t* swu32 = reinterpret_cast<const uint32_t*>(swu);
      // ldmatrix lane -> (matrix mi = lane / 8: rows +8 for odd mi, k +8 for mi >= 2)
      const int lrow = (lane & 7) + 8 * ((lane >> 3) & 1), lk = 8 * (lane >> 4);
#pragma unroll 1
      for (int nb = 0; nb < 9; ++nb) {
        const int py = warp + nb / 3, px = lrow + nb % 3;        // staged coordinates (halo offset +1, dp/dq offset -1)
        const uint32_t abase = smem_addr(sx + (py * UP_PC + px) * UP_PIX + lk);
#pragma unroll
        for (int kc = 0; kc < 4; ++kc) {
          uint32_t a[4];
          ldsm_x4(a, abase + kc * 32);
          const int kw = (nb * 64 + kc * 16 + 2 * t) >> 1;
          mma16816(acc[0], a, swu32[g * (UP_WROW / 2) + kw], swu32[g * (UP_WROW / 2) + kw + 4]);
          mma16816(acc[1], a, swu32[(g + 8) * (UP_WROW / 2) + kw], swu32[(g + 8) * (UP_WROW / 2) + kw + 4]);
        }
      }
    }
    // ---- epilogue: n = 2t, 2t+1 (n-tile 0) and 8+2t, 9+2t (n-tile 1); rows g and g+8 of the position row
    const int mi = p0 + warp;
    if (mi < Ho) {
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int mj = q0 + g + 8 * half;
        if (mj >= Wo) continue;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int nn = 8 * j + 2 * t;
          if (nn >= 12) continue;
          float v[2] = {acc[j][2 * half], acc[j][2 * half + 1]};
          if (bias) { v[0] += __ldg(bias + nn % 3); v[1] += __ldg(bias + (nn + 1) % 3); }
          act_fwd_vec<2>(v, act, act_param);
          const int arow = nn / 6, off = nn - 6 * arow;             // output row parity, offset inside the 6-float run
          float* dst = large + (((int64_t)n * H + 2 * mi + arow) * W + 2 * mj) * 3 + off;
          *reinterpret_cast<float2*>(dst) = make_float2(v[0], v[1]);
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// c3m_wgrad.  CTA = 8 warps, persistent over tiles of 8 x 16 small pixels (K = 128 pixels per tile);
// warp w owns m16 tile w = m-blocks 2w, 2w+1;  m-block mb = (r = mb/3, s0 = 2*(mb%3)), entry i8 -> (s0 + i8/4, c = i8%4).
constexpr int WG_TH = 8, WG_TW = 16;
constexpr int WG_PR = 2 * WG_TH + 3;            // 19 patch rows
constexpr int WG_PC = 2 * WG_TW + 4;            // 36 patch columns (35 + the column the discarded s = 5 entries read)
constexpr int WG_YPIX = 72;                     // bf16 per staged small pixel

template <typename TSM>
__global__ void __launch_bounds__(256, 2)
c3m_wgrad_kernel(const float* __restrict__ large, const TSM* __restrict__ small, float* __restrict__ dw, float* __restrict__ dbias,
                 int N, int H, int W, int Ho, int Wo, int K, int kblocks) {
  pdl_grid_sync();
  __shared__ __align__(16) bf16 sp[WG_PR * WG_PC * 4];          // [row][col][4]   5,472 B
  __shared__ __align__(16) bf16 sy[WG_TH * WG_TW * WG_YPIX];    // [pixel][72]    18,432 B
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int kb = (blockIdx.x % kblocks) * 64;
  const int tiles_w = (Wo + WG_TW - 1) / WG_TW, tiles_h = (Ho + WG_TH - 1) / WG_TH;
  const int ntiles = N * tiles_h * tiles_w;
  float acc[8][4];
#pragma unroll
  for (int j = 0; j < 8; ++j) { acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f; }
  // ldmatrix.trans lane roles.  A: matrix mi = lane/8 -> (k block = mi >> 1, m block = 2*warp + (mi & 1));
  // B: matrix mi -> (k block = mi & 1, n block = jn + (mi >> 1)).
  const int l8 = lane & 7, lm = lane >> 3;
  const int mblk = min(2 * warp + (lm & 1), 14);
  const int a_r = mblk / 3, a_s0 = 2 * (mblk % 3);
  // Software pipeline: the next tile's patch pixels and small-tensor chunks are loaded into registers while the current
  // tile is multiplied (the kernel was latency-bound on these loads: ncu long-scoreboard stalls 10.9 per issue).
  constexpr int PPX = (WG_PR * WG_PC + 255) / 256;              // 3 patch pixels per thread
  constexpr int PYC = WG_TH * WG_TW * 8 / 256;                  // 4 16-byte chunks of the small tile per thread
  int pab[PPX];                                                 // a | b << 8 of this thread's patch pixels, -1 = none
#pragma unroll
  for (int k = 0; k < PPX; ++k) {
    const int e = tid + 256 * k;
    pab[k] = (e < WG_PR * WG_PC) ? ((e / WG_PC) | ((e % WG_PC) << 8)) : -1;
  }
  float pv[PPX][3];
  uint4 yv[PYC];
  auto prefetch = [&](int tile) {
    const int n = tile / (tiles_h * tiles_w);
    const int rem = tile - n * tiles_h * tiles_w;
    const int p0 = (rem / tiles_w) * WG_TH, q0 = (rem % tiles_w) * WG_TW;
    const int i0 = 2 * p0 - 1, j0 = 2 * q0 - 1;
#pragma unroll
    for (int k = 0; k < PPX; ++k) {
      const int i = i0 + (pab[k] & 0xff), j = j0 + (pab[k] >> 8);
      const bool ok = pab[k] >= 0 && i >= 0 && i < H && j >= 0 && j < W;
      const float* src = large + (((int64_t)n * H + (ok ? i : 0)) * W + (ok ? j : 0)) * 3;
      pv[k][0] = ok ? __ldg(src) : 0.f; pv[k][1] = ok ? __ldg(src + 1) : 0.f; pv[k][2] = ok ? __ldg(src + 2) : 0.f;
    }
#pragma unroll
    for (int k = 0; k < PYC; ++k) {
      const int e = tid + 256 * k;
      const int pix = e >> 3, c8 = (e & 7) * 8;
      const int p = p0 + pix / WG_TW, q = q0 + pix % WG_TW;
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if (p < Ho && q < Wo) {
        const TSM* src = small + (((int64_t)n * Ho + p) * Wo + q) * K + kb + c8;
        if (sizeof(TSM) == 2) {
          v = __ldg(reinterpret_cast<const uint4*>(src));
        } else {
          const float4 f0 = ld4(reinterpret_cast<const float*>(src)), f1 = ld4(reinterpret_cast<const float*>(src) + 4);
          v = make_uint4(pack2(f0.x, f0.y), pack2(f0.z, f0.w), pack2(f1.x, f1.y), pack2(f1.z, f1.w));
        }
      }
      yv[k] = v;
    }
  };
  const int tstride = gridDim.x / kblocks;
  int tile = blockIdx.x / kblocks;
  if (tile < ntiles) prefetch(tile);
  for (; tile < ntiles; tile += tstride) {
    __syncthreads();                                            // previous tile's fragments are consumed
    // The staged pixels have a fourth, unused channel slot (its filter-gradient rows are discarded).  With dbias it holds 1 for
    // in-image pixels: the row of tap (1, 1) -- input pixel (2p, 2q), inside the image for every small pixel -- then accumulates
    // sum_pixels 1 * small[pixel, k] = the bias gradient of the conv, at no extra MMA.
    int ti0 = 0, tj0 = 0;
    if (dbias) {
      const int rem = tile % (tiles_h * tiles_w);
      ti0 = 2 * ((rem / tiles_w) * WG_TH) - 1; tj0 = 2 * ((rem % tiles_w) * WG_TW) - 1;
    }
#pragma unroll
    for (int k = 0; k < PPX; ++k)
      if (pab[k] >= 0) {
        const int i = ti0 + (pab[k] & 0xff), j = tj0 + (pab[k] >> 8);
        const float one = (dbias && i >= 0 && i < H && j >= 0 && j < W) ? 1.f : 0.f;
        *reinterpret_cast<uint2*>(sp + (tid + 256 * k) * 4) = make_uint2(pack2(pv[k][0], pv[k][1]), pack2(pv[k][2], one));
      }
#pragma unroll
    for (int k = 0; k < PYC; ++k) {
      const int e = tid + 256 * k;
      *reinterpret_cast<uint4*>(sy + (e >> 3) * WG_YPIX + (e & 7) * 8) = yv[k];
    }
    __syncthreads();
    if (tile + tstride < ntiles) prefetch(tile + tstride);
    __syncthreads();
#pragma unroll 2
    for (int ks = 0; ks < WG_TH * WG_TW / 16; ++ks) {          // 16 pixels per step = one tile row (WG_TW == 16)
      // A: stored rows = pixels (py = ks, px = 8*(lm>>1) + l8), 8 consecutive m = 2 columns x 4 channels
      uint32_t a[4];
      {
        const int px = 8 * (lm >> 1) + l8;
        ldsm_x4_t(a, smem_addr(sp + (((2 * ks + a_r) * WG_PC) + 2 * px + a_s0) * 4));
      }
#pragma unroll
      for (int jn = 0; jn < 8; jn += 2) {
        uint32_t b[4];
        const int pix = ks * 16 + 8 * (lm & 1) + l8;
        ldsm_x4_t(b, smem_addr(sy + pix * WG_YPIX + (jn + (lm >> 1)) * 8));
        mma16816(acc[jn], a, b[0], b[1]);
        mma16816(acc[jn + 1], a, b[2], b[3]);
      }
    }
  }
  // ---- reduce into dw: rows m = g, g+8 of this warp's m16 tile
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int mb = 2 * warp + half;                             // m = 8*half + g  ->  m block, entry g
    const int r = mb / 3, s = 2 * (mb % 3) + (g >> 2), c = g & 3;
    if (dbias && r == 1 && s == 1 && c == 3) {                  // the ones-channel row of tap (1, 1): bias gradient
#pragma unroll
      for (int jn = 0; jn < 8; ++jn)
        asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(dbias + kb + 2 * t + jn * 8), "f"(acc[jn][2 * half]), "f"(acc[jn][2 * half + 1]) : "memory");
    }
    if (mb >= 15 || s >= KT5 || c >= 3) continue;
    float* dst = dw + ((int64_t)((r * KT5 + s) * 3 + c)) * K + kb + 2 * t;
#pragma unroll
    for (int jn = 0; jn < 8; ++jn) {
      asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(dst + jn * 8), "f"(acc[jn][2 * half]), "f"(acc[jn][2 * half + 1]) : "memory");
    }
  }
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// bnb (optional, bf16 output, one row group): also accumulate the consumer batch norm's backward reductions; *fused = 1 when done
int c3m_conv_down(const gg_conv_desc* d, const float* large, const float* w, const float* bias, void* small, cudaStream_t st,
                  const float* pre = nullptr, const float* mean = nullptr, const float* rstd = nullptr, const float* gamma = nullptr,
                  const float* beta = nullptr, double* sums = nullptr, int bact = 0, float bact_param = 0.f, int* fused = nullptr) {
  const int kblocks = d->K / 64;
  const int ntiles = d->N * ceil_div(d->Ho, DN_TH) * ceil_div(d->Wo, DN_TW);
  const int per_k = std::max(1, std::min(ntiles, (148 * 2) / std::min(kblocks, 296)));     // 2 co-resident CTAs per SM
  const int grid = per_k * kblocks;
  C3Bnb b = {pre, mean, rstd, gamma, beta, sums, bact, bact_param};
  if (pre != nullptr && sums != nullptr && d->small_dtype == GG_BF16 && ((uintptr_t)pre % 16) == 0) {
    Launch(grid, 256, 0, st)(c3m_down_kernel<bf16, true>, large, w, bias, (bf16*)small, d->N, d->H, d->W, d->Ho, d->Wo, d->K, kblocks, d->act, d->act_param, b);
    if (fused) *fused = 1;
  } else if (d->small_dtype == GG_F32) {
    Launch(grid, 256, 0, st)(c3m_down_kernel<float, false>, large, w, bias, (float*)small, d->N, d->H, d->W, d->Ho, d->Wo, d->K, kblocks, d->act, d->act_param, b);
  } else {
    Launch(grid, 256, 0, st)(c3m_down_kernel<bf16, false>, large, w, bias, (bf16*)small, d->N, d->H, d->W, d->Ho, d->Wo, d->K, kblocks, d->act, d->act_param, b);
  }
  return check_launch("c3m_down");
}

int c3m_conv_up(const gg_conv_desc* d, const void* small, const float* w, const float* bias, float* large, cudaStream_t st) {
  const int ntiles = d->N * ceil_div(d->Ho, UP_TH) * ceil_div(d->Wo, UP_TW);
  const int grid = std::min(ntiles, 148 * 3);     // 3 co-resident CTAs per SM (45 KB of shared memory, <= 85 registers each)
  if (d->small_dtype == GG_F32)
    Launch(grid, 256, 0, st)(c3m_up_kernel<float>, (const float*)small, w, bias, large, d->N, d->H, d->W, d->Ho, d->Wo, d->K, d->act, d->act_param);
  else
    Launch(grid, 256, 0, st)(c3m_up_kernel<bf16>, (const bf16*)small, w, bias, large, d->N, d->H, d->W, d->Ho, d->Wo, d->K, d->act, d->act_param);
  return check_launch("c3m_up");
}

// dbias (optional): the conv's bias gradient sum_pixels small[pixel, k] is accumulated (+=) by the same launch
int c3m_conv_wgrad(const gg_conv_desc* d, const float* large, const void* small, float* dw, cudaStream_t st, float* dbias) {
  const int kblocks = d->K / 64;
  const int ntiles = d->N * ceil_div(d->Ho, WG_TH) * ceil_div(d->Wo, WG_TW);
  const int per_k = std::max(1, std::min(ntiles, (148 * 2) / std::min(kblocks, 296)));
  const int grid = per_k * kblocks;
  if (d->small_dtype == GG_F32)
    Launch(grid, 256, 0, st)(c3m_wgrad_kernel<float>, large, (const float*)small, dw, dbias, d->N, d->H, d->W, d->Ho, d->Wo, d->K, kblocks);
  else
    Launch(grid, 256, 0, st)(c3m_wgrad_kernel<bf16>, large, (const bf16*)small, dw, dbias, d->N, d->H, d->W, d->Ho, d->Wo, d->K, kblocks);
  return check_launch("c3m_wgrad");
}

}  // namespace gg
