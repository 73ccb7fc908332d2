// Thin linear + batch norm (train) + activation as ONE kernel per direction: the generator's input projection
//   h0 = relu(g_bn0(reshape(linear(z, gf*8*s16*s16), [-1, s16, s16, gf*8])))      (model.py:304-307)
// z is [B, 100] and the output [B, 8192] is normalised per channel c = column % Cc over B * s16 * s16 values.  As separate
// launches this tiny layer cost ~24 us forward (thin_fwd + memset + colsum + bn_train_apply) and ~32 us backward (colsum +
// bn_bwd_apply + thin_wgrad) of a 1.5 ms train step; here a CTA owns CPC channels -- ALL their columns (every spatial
// position) and ALL rows -- so the batch statistics and the backward reductions never leave the CTA: no atomics, no second
// pass, one launch.
//
// Thread mapping (256 threads): column-in-CTA = t % NCOL, row group = t / NCOL (16 rows each), NCOL = 256 / NRG;
// column-in-CTA = pos * CPC + k  ->  global column = pos * Cc + c0 + k  (k fastest: CPC consecutive channels are contiguous).
#include <algorithm>

#include "common.cuh"

namespace gg {

constexpr int LB_THREADS = 256, LB_ROWS = 16, LB_MAXK = 128;      // 16 rows per thread: NRG = 4 (<= 64 rows) or 8 (<= 128 rows) row groups
// MEASURED (round 2, B200, DCGAN-64 batch 64: rows 64, in 100, out 8192, C 512): forward 19.5-21 us, backward 31.5-34 us per
// launch -- about what the composed path costs (thin_fwd 9.7 + memset 1.2 + colsum 7-12 + apply 4-5; colsum 8 + apply 4 +
// thin_wgrad 19.8), step time unchanged (1.480 vs 1.483 ms).  ncu (profiles/r3a_linbn_ncu.md): 128 CTAs of 8 warps at 100+
// registers = 12 % warps active, issue slots 19-32 % busy, stalls on the global-load and shared-memory scoreboards: a chain of
// dependent round trips with two warps per scheduler.  4 rows per thread (4x the CTAs) made it 3x SLOWER (every CTA re-stages x
// and re-reads its matrix columns 16 times).  The fused kernels are therefore opt-in (GG_FUSE_LINEAR_BN=1); they are parity-tested.

template <typename T> __device__ __forceinline__ float lb_round(float v);
template <> __device__ __forceinline__ float lb_round<float>(float v) { return v; }
template <> __device__ __forceinline__ float lb_round<bf16>(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

// stage x[rows, in_dim] transposed in shared memory: feature i of row r sits at xs[i * RP + ((r + 4 * i) % RP)], RP = NRG * 16
// (64 or 128; rows beyond `rows` are zero).  The global reads are coalesced (a warp reads 32 consecutive features of one row)
// and all of a thread's loads are independent (fully unrolled: in flight together); the rotation by 4 * i keeps every aligned
// group of four rows contiguous and 16-byte aligned for the float4 broadcast reads of the GEMM loop while spreading the
// transposed stores over the banks (4-way instead of 32-way conflicts).  lb_xs() is the matching read address.
template <int RP>
__device__ __forceinline__ const float4* lb_xs(const float* xs, int i, int r4) {     // r4: first row of an aligned group of four
  return reinterpret_cast<const float4*>(xs + i * RP + ((r4 + 4 * i) & (RP - 1)));
}
template <typename TX, int RP>
__device__ __forceinline__ void lb_stage_x(const TX* __restrict__ x, float* xs, int rows, int in_dim) {
  // thread t: feature lane f = t % 32, rows r = t / 32 + 8 * j; features i = f + 32 * m
  constexpr int NM = LB_MAXK / 32, NJ = RP / 8;
  const int f = threadIdx.x & 31, r0 = threadIdx.x >> 5;
#pragma unroll
  for (int m = 0; m < NM; ++m) {
    const int i = f + 32 * m;
    float v[NJ];
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const int r = r0 + 8 * j;
      v[j] = (i < in_dim && r < rows) ? ldf(x + (int64_t)r * in_dim + i) : 0.f;
    }
    if (i < in_dim) {
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const int r = r0 + 8 * j;
        xs[i * RP + ((r + 4 * i) & (RP - 1))] = v[j];
      }
    }
  }
}

template <typename TX, typename TY, int NRG>
__global__ void __launch_bounds__(LB_THREADS)
linbn_fwd_kernel(const TX* __restrict__ x, const float* __restrict__ W, const float* __restrict__ bias, const float* __restrict__ gamma,
                 const float* __restrict__ beta, float* __restrict__ moving_mean, float* __restrict__ moving_var, float* __restrict__ pre,
                 TY* __restrict__ y, float* __restrict__ save_mean, float* __restrict__ save_rstd, int rows, int in_dim, int out_dim, int Cc,
                 int CPC, float eps, float decay, int act, float act_param) {
  pdl_grid_sync();
  constexpr int NCOL = LB_THREADS / NRG;
  extern __shared__ __align__(16) float lb_smem[];
  constexpr int RP = NRG * LB_ROWS;
  float* xs = lb_smem;                                   // [in_dim][RP]
  float* red = xs + in_dim * RP;                         // [2][256] partial (sum, sum of squares)
  float* stat = red + 2 * LB_THREADS;                    // [2][CPC] (mean, rstd)
  lb_stage_x<TX, RP>(x, xs, rows, in_dim);
  __syncthreads();
  const int cl = threadIdx.x % NCOL, rg = threadIdx.x / NCOL;
  const int pos = cl / CPC, k = cl - pos * CPC;
  const int c0 = blockIdx.x * CPC;
  const int col = pos * Cc + c0 + k;
  float acc[LB_ROWS];
#pragma unroll
  for (int r = 0; r < LB_ROWS; ++r) acc[r] = 0.f;
  {
    constexpr int NB = 32;                                // matrix elements of this column per load batch
    for (int i0 = 0; i0 < in_dim; i0 += NB) {
      float m[NB];
#pragma unroll
      for (int u = 0; u < NB; ++u) m[u] = (i0 + u < in_dim) ? __ldg(W + (int64_t)(i0 + u) * out_dim + col) : 0.f;
#pragma unroll
      for (int u = 0; u < NB; ++u) {
        if (i0 + u < in_dim) {
#pragma unroll
          for (int q = 0; q < LB_ROWS / 4; ++q) {
            const float4 v = *lb_xs<RP>(xs, i0 + u, rg * LB_ROWS + 4 * q);
            acc[4 * q] = fmaf(v.x, m[u], acc[4 * q]); acc[4 * q + 1] = fmaf(v.y, m[u], acc[4 * q + 1]);
            acc[4 * q + 2] = fmaf(v.z, m[u], acc[4 * q + 2]); acc[4 * q + 3] = fmaf(v.w, m[u], acc[4 * q + 3]);
          }
        }
      }
    }
  }
  const float b = bias ? __ldg(bias + col) : 0.f;
  float s = 0.f, q2 = 0.f;
#pragma unroll
  for (int r = 0; r < LB_ROWS; ++r) {
    acc[r] += b;
    if (rg * LB_ROWS + r < rows) { s += acc[r]; q2 = fmaf(acc[r], acc[r], q2); }
  }
  red[threadIdx.x] = s;
  red[LB_THREADS + threadIdx.x] = q2;
  __syncthreads();
  const int npos = NCOL / CPC;
  if (threadIdx.x < CPC) {
    // channel c0 + threadIdx.x: fold its npos positions x NRG row groups, in a fixed order, in double
    double a = 0.0, bq = 0.0;
    for (int g = 0; g < NRG; ++g)
      for (int p2 = 0; p2 < npos; ++p2) {
        const int t = g * NCOL + p2 * CPC + threadIdx.x;
        a += (double)red[t]; bq += (double)red[LB_THREADS + t];
      }
    const double inv = 1.0 / ((double)rows * npos);
    const double m = a * inv;
    double var = bq * inv - m * m;
    var = var < 0.0 ? 0.0 : var;
    const float mu = (float)m, rs = rsqrtf((float)var + eps);
    stat[threadIdx.x] = mu; stat[CPC + threadIdx.x] = rs;
    const int ch = c0 + threadIdx.x;
    save_mean[ch] = mu; save_rstd[ch] = rs;
    // assign_moving_average(moving, batch, decay) = moving - (moving - batch) * (1 - decay)   (ops.py:18-24)
    if (moving_mean) moving_mean[ch] -= (moving_mean[ch] - mu) * (1.f - decay);
    if (moving_var) moving_var[ch] -= (moving_var[ch] - (float)var) * (1.f - decay);
  }
  __syncthreads();
  const float mu = stat[k], rs = stat[CPC + k];
  const float ga = gamma ? __ldg(gamma + c0 + k) : 1.f, be = beta ? __ldg(beta + c0 + k) : 0.f;
  {
    float o[LB_ROWS];
#pragma unroll
    for (int j = 0; j < LB_ROWS; ++j) o[j] = fmaf((acc[j] - mu) * rs, ga, be);
    act_fwd_vec<LB_ROWS>(o, act, act_param);
#pragma unroll
    for (int j = 0; j < LB_ROWS; ++j) {
      const int r = rg * LB_ROWS + j;
      if (r < rows) {
        pre[(int64_t)r * out_dim + col] = acc[j];
        stf(y + (int64_t)r * out_dim + col, o[j]);
      }
    }
  }
}

// backward: dy [rows, out_dim] -> (dW += x^T dpre, dgamma +=, dbeta +=); the bias gradient of a train-mode batch norm's
// producer is exactly zero and the input (z) needs no gradient.  ROUND_BF16: dpre is rounded to bf16 before the filter
// gradient, as the two-kernel path does when it materialises dpre in the activation dtype.
template <typename TX, typename TD, int NRG, bool ROUND_BF16>
__global__ void __launch_bounds__(LB_THREADS)
linbn_bwd_kernel(const TX* __restrict__ x, const float* __restrict__ pre, const TD* __restrict__ dy, const float* __restrict__ gamma,
                 const float* __restrict__ beta, const float* __restrict__ save_mean, const float* __restrict__ save_rstd,
                 float* __restrict__ dW, float* __restrict__ dgamma, float* __restrict__ dbeta, int rows, int in_dim, int out_dim, int Cc,
                 int CPC, int act, float act_param) {
  pdl_grid_sync();
  constexpr int NCOL = LB_THREADS / NRG;
  extern __shared__ __align__(16) float lb_smem[];
  constexpr int RP = NRG * LB_ROWS;
  float* xs = lb_smem;                                   // [in_dim][RP]
  float* red = xs + in_dim * RP;                         // [NRG][32][NCOL] = 32 KB: partial filter gradients; first used as [2][256]
  float* stat = red + NRG * 32 * NCOL;                   // [2][CPC]
  lb_stage_x<TX, RP>(x, xs, rows, in_dim);
  const int cl = threadIdx.x % NCOL, rg = threadIdx.x / NCOL;
  const int pos = cl / CPC, k = cl - pos * CPC;
  const int c0 = blockIdx.x * CPC;
  const int col = pos * Cc + c0 + k;
  const int npos = NCOL / CPC;
  const float mu = __ldg(save_mean + c0 + k), rs = __ldg(save_rstd + c0 + k);
  const float ga = gamma ? __ldg(gamma + c0 + k) : 1.f, be = beta ? __ldg(beta + c0 + k) : 0.f;
  float g[LB_ROWS], xh[LB_ROWS];
#pragma unroll
  for (int r = 0; r < LB_ROWS; ++r) {
    const int rr = rg * LB_ROWS + r;
    const bool ok = rr < rows;
    xh[r] = ok ? __ldg(pre + (int64_t)rr * out_dim + col) : mu;
    g[r] = ok ? ldf(dy + (int64_t)rr * out_dim + col) : 0.f;
  }
  float s0 = 0.f, s1 = 0.f;
  {
    float u[LB_ROWS];
#pragma unroll
    for (int r = 0; r < LB_ROWS; ++r) { xh[r] = (xh[r] - mu) * rs; u[r] = fmaf(ga, xh[r], be); }
    act_bwd_pre_vec<LB_ROWS>(g, u, act, act_param);
#pragma unroll
    for (int r = 0; r < LB_ROWS; ++r) { s0 += g[r]; s1 = fmaf(g[r], xh[r], s1); }
  }
  red[threadIdx.x] = s0;
  red[LB_THREADS + threadIdx.x] = s1;
  __syncthreads();                                       // (also: xs is staged)
  if (threadIdx.x < CPC) {
    double a = 0.0, b = 0.0;
    for (int gi = 0; gi < NRG; ++gi)
      for (int p2 = 0; p2 < npos; ++p2) {
        const int t = gi * NCOL + p2 * CPC + threadIdx.x;
        a += (double)red[t]; b += (double)red[LB_THREADS + t];
      }
    const float invM = 1.f / ((float)rows * npos);
    stat[threadIdx.x] = (float)a * invM; stat[CPC + threadIdx.x] = (float)b * invM;
    const int ch = c0 + threadIdx.x;
    if (dbeta) dbeta[ch] += (float)a;
    if (dgamma) dgamma[ch] += (float)b;
  }
  __syncthreads();
  {
    const float sg = stat[k], sgx = stat[CPC + k];
#pragma unroll
    for (int r = 0; r < LB_ROWS; ++r) {
      const float d = ga * rs * (g[r] - sg - xh[r] * sgx);
      g[r] = (rg * LB_ROWS + r < rows) ? (ROUND_BF16 ? lb_round<bf16>(d) : d) : 0.f;      // dpre
    }
  }
  // dW[i][col] += sum_r x[r][i] * dpre[r]: chunks of 32 input features; the NRG row groups meet in shared memory (fixed order)
  for (int i0 = 0; i0 < in_dim; i0 += 32) {
    float part[32];
#pragma unroll
    for (int ii = 0; ii < 32; ++ii) part[ii] = 0.f;
#pragma unroll
    for (int q = 0; q < LB_ROWS / 4; ++q) {
#pragma unroll
      for (int ii = 0; ii < 32; ++ii) {
        if (i0 + ii < in_dim) {
          const float4 v = *lb_xs<RP>(xs, i0 + ii, rg * LB_ROWS + 4 * q);
          part[ii] = fmaf(v.x, g[4 * q], part[ii]); part[ii] = fmaf(v.y, g[4 * q + 1], part[ii]);
          part[ii] = fmaf(v.z, g[4 * q + 2], part[ii]); part[ii] = fmaf(v.w, g[4 * q + 3], part[ii]);
        }
      }
    }
    __syncthreads();                                     // previous chunk's readers are done with `red`
#pragma unroll
    for (int ii = 0; ii < 32; ++ii) red[(rg * 32 + ii) * NCOL + cl] = part[ii];
    __syncthreads();
    constexpr int PER = 32 / NRG;                        // input features this thread finishes
#pragma unroll
    for (int j = 0; j < PER; ++j) {
      const int ii = rg * PER + j;
      if (i0 + ii < in_dim) {
        float sum = 0.f;
#pragma unroll
        for (int gi = 0; gi < NRG; ++gi) sum += red[(gi * 32 + ii) * NCOL + cl];
        dW[(int64_t)(i0 + ii) * out_dim + col] += sum;  // single writer per element
      }
    }
  }
}

static int lb_nrg(int rows) { return rows <= 64 ? 4 : (rows <= 128 ? 8 : 0); }

}  // namespace gg

using namespace gg;

// eligible: a thin input (in_dim <= 128), at most 128 rows, one row group, every channel's columns inside one CTA
extern "C" int gg_linear_bn_ok(int32_t rows, int32_t in_dim, int32_t out_dim, int32_t Cc, int32_t groups) {
  const int nrg = lb_nrg(rows);
  if (nrg == 0 || groups != 1 || in_dim > LB_MAXK || in_dim < 1 || Cc < 1 || out_dim % Cc != 0) return 0;
  const int ncol = LB_THREADS / nrg, npos = out_dim / Cc;
  if (npos > ncol || ncol % npos != 0) return 0;
  const int cpc = ncol / npos;
  return (Cc % cpc == 0) ? 1 : 0;
}

extern "C" int gg_linear_bn_fwd(const void* x, int32_t x_dt, const float* matrix, const float* bias, const float* gamma, const float* beta,
                                float* moving_mean, float* moving_var, float* pre, void* y, int32_t y_dt, float* save_mean, float* save_rstd,
                                int32_t rows, int32_t in_dim, int32_t out_dim, int32_t Cc, float eps, float decay, int32_t act, float act_param,
                                void* stream) {
  GG_REQUIRE(x && matrix && pre && y && save_mean && save_rstd, GG_ERR_INVALID, "linear_bn_fwd: null pointer");
  GG_REQUIRE(gg_linear_bn_ok(rows, in_dim, out_dim, Cc, 1), GG_ERR_UNSUPPORTED, "linear_bn_fwd: shape not eligible (rows %d in %d out %d C %d)", rows, in_dim, out_dim, Cc);
  const int nrg = lb_nrg(rows), ncol = LB_THREADS / nrg, cpc = ncol / (out_dim / Cc);
  const size_t smem = ((size_t)in_dim * nrg * LB_ROWS + 2 * LB_THREADS + 2 * cpc) * sizeof(float);
  cudaStream_t st = (cudaStream_t)stream;
#define GG_LBF(TX, TY, N)                                                                                                                   \
  do {                                                                                                                                      \
    static bool attr = false;                                                                                                               \
    if (!attr) { cudaFuncSetAttribute(linbn_fwd_kernel<TX, TY, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024); attr = true; } \
    Launch(Cc / cpc, LB_THREADS, smem, st)(linbn_fwd_kernel<TX, TY, N>, (const TX*)x, matrix, bias, gamma, beta, moving_mean, moving_var, pre, \
                                           (TY*)y, save_mean, save_rstd, rows, in_dim, out_dim, Cc, cpc, eps, decay, act, act_param);        \
  } while (0)
#define GG_LBF2(TX, TY) do { if (nrg == 4) GG_LBF(TX, TY, 4); else GG_LBF(TX, TY, 8); } while (0)
  if (x_dt == GG_F32 && y_dt == GG_F32) GG_LBF2(float, float);
  else if (x_dt == GG_F32) GG_LBF2(float, bf16);
  else if (y_dt == GG_F32) GG_LBF2(bf16, float);
  else GG_LBF2(bf16, bf16);
  return check_launch("linear_bn_fwd");
}

extern "C" int gg_linear_bn_bwd(const void* x, int32_t x_dt, const float* pre, const void* dy, int32_t dy_dt, const float* gamma, const float* beta,
                                const float* save_mean, const float* save_rstd, float* dmatrix, float* dgamma, float* dbeta, int32_t rows,
                                int32_t in_dim, int32_t out_dim, int32_t Cc, int32_t act, float act_param, int32_t round_bf16, void* stream) {
  GG_REQUIRE(x && pre && dy && save_mean && save_rstd && dmatrix, GG_ERR_INVALID, "linear_bn_bwd: null pointer");
  GG_REQUIRE(gg_linear_bn_ok(rows, in_dim, out_dim, Cc, 1), GG_ERR_UNSUPPORTED, "linear_bn_bwd: shape not eligible");
  const int nrg = lb_nrg(rows), ncol = LB_THREADS / nrg, cpc = ncol / (out_dim / Cc);
  const size_t smem = ((size_t)in_dim * nrg * LB_ROWS + (size_t)nrg * 32 * ncol + 2 * cpc) * sizeof(float);
  cudaStream_t st = (cudaStream_t)stream;
#define GG_LBB(TX, TD, N, R)                                                                                                                  \
  do {                                                                                                                                        \
    static bool attr = false;                                                                                                                 \
    if (!attr) { cudaFuncSetAttribute(linbn_bwd_kernel<TX, TD, N, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024); attr = true; } \
    Launch(Cc / cpc, LB_THREADS, smem, st)(linbn_bwd_kernel<TX, TD, N, R>, (const TX*)x, pre, (const TD*)dy, gamma, beta, save_mean, save_rstd, \
                                           dmatrix, dgamma, dbeta, rows, in_dim, out_dim, Cc, cpc, act, act_param);                            \
  } while (0)
#define GG_LBB2(TX, TD)                                                                       \
  do {                                                                                        \
    if (nrg == 4) { if (round_bf16) GG_LBB(TX, TD, 4, true); else GG_LBB(TX, TD, 4, false); } \
    else { if (round_bf16) GG_LBB(TX, TD, 8, true); else GG_LBB(TX, TD, 8, false); }          \
  } while (0)
  if (x_dt == GG_F32 && dy_dt == GG_F32) GG_LBB2(float, float);
  else if (x_dt == GG_F32) GG_LBB2(float, bf16);
  else if (dy_dt == GG_F32) GG_LBB2(bf16, float);
  else GG_LBB2(bf16, bf16);
  return check_launch("linear_bn_bwd");
}
