// Batch-norm forward/backward fused with the ReLU / LeakyReLU / tanh / sigmoid epilogues.
// Replaces tf.contrib.layers.batch_norm (/root/reference/models/recurrent_z/ops.py:18-24)
// and tf.nn.moments + tf.nn.batch_normalization (rnn_test/recurrent_DCGAN.py:190-191),
// plus the activation that always follows them (model.py:274-276, 307-319).
//
// HBM-bound: x is [rows, C] channel-contiguous.  Statistics pass = one coalesced read of x
// (4 channels per thread, 16 B loads; per-thread fp32 partials over a bounded number of
// rows, block-level smem reduction, one fp64 atomic per channel per CTA so the final
// E[x^2]-mean^2 has no cancellation problem at the 1e-4 parity tolerance); apply pass = one
// read + one write.  Algorithmic bytes: fwd 2 reads + 1 write, bwd 2x(x,dy) reads + 1 write.
#include <stdlib.h>

#include <algorithm>
#include <mutex>

#include <cooperative_groups.h>

#include "common.cuh"

namespace gg {

constexpr int BN_THREADS = 256;
#ifndef GG_BN_UB
#define GG_BN_UB 4          // rows per load batch in the streaming kernels (A/B switch: -DGG_BN_UB=1)
#endif

struct ColGeom {
  int tx, ty;         // block = tx * ty threads; tx over channel vectors, ty over rows
  int cblocks;        // grid.y
  int rblocks;        // grid.x
  int vec;
};

// Threads per CTA of the column REDUCTIONS (colsum, bias_grad).  Their grid is kept at ~2 CTAs per SM (every CTA ends with
// 2 C fp64 atomics on the same addresses), so with 256 threads an SM held 16 warps and a thread walked its rows in ~7
// dependent load batches: latency-bound at ~1.3 TB/s.  More threads per CTA = the same grid, more warps, fewer batches.
// Measured inside the step (same box, GG_COLSUM_THREADS): 256 -> 1.4192, 512 -> 1.4125, 1024 -> 1.4153 ms: 512.
constexpr int CS_MAX_THREADS = 1024;
static int colsum_threads() {
  static const int t = [] {
    const char* v = getenv("GG_COLSUM_THREADS");
    const int n = (v && *v) ? atoi(v) : 512;
    return (n == 256 || n == 512 || n == 1024) ? n : 512;
  }();
  return t;
}

static ColGeom col_geom(int64_t rows_per_group, int C, int groups, bool vec_ok, int threads = BN_THREADS) {
  ColGeom g;
  g.vec = (vec_ok && C % 4 == 0) ? 4 : 1;
  const int lanes = C / g.vec;
  int tx = 1;
  while (tx < lanes && tx < BN_THREADS) tx <<= 1;
  g.tx = tx;
  g.ty = threads / tx;
  g.cblocks = ceil_div(lanes, tx);
  // ~2 CTAs per SM in total: every CTA ends with 2*C fp64 atomics on the same C addresses, so the tail cost grows
  // with the CTA count while the streaming part is bandwidth-trivial at these sizes
  static const int per_sm = [] { const char* v = getenv("GG_COLSUM_PER_SM"); return (v && *v) ? atoi(v) : 2; }();
  int64_t want = std::max<int64_t>(1, (148 * per_sm) / ((int64_t)g.cblocks * groups));
  int64_t maxr = ceil_div64(rows_per_group, (int64_t)g.ty * 4);  // >= 4 rows (one load batch) per thread
  g.rblocks = (int)std::max<int64_t>(1, std::min<int64_t>(want, maxr));
  return g;
}

// ---- per-(group, channel) sums of f0(x,...) and f1(x,...) -----------------------------
// MODE 0: (x, x^2)                                   -> batch statistics
// MODE 1: (g, g*xhat), g = dy*act'(gamma*xhat+beta)  -> BN backward reductions
// MODE 2: (x, 0) column sum                          -> bias gradient
template <typename TX, typename TD, int VEC, int MODE>
__global__ void __launch_bounds__(MODE == 1 ? 512 : CS_MAX_THREADS)      // (the backward reductions need > 64 registers: 512 threads)
colsum_kernel(const TX* __restrict__ x, const TD* __restrict__ dy, int64_t rows_per_group, int C,
              const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ mean,
              const float* __restrict__ rstd, int act, float act_param, double* __restrict__ sums, int tx_dim, int R) {
  pdl_grid_sync();
  const int tx = threadIdx.x % tx_dim, ty = threadIdx.x / tx_dim, ty_dim = (int)blockDim.x / tx_dim;
  const int grp = blockIdx.z;
  const int c = (blockIdx.y * tx_dim + tx) * VEC;
  const bool active = c < C;
  float s0[VEC], s1[VEC], ga[VEC], be[VEC], mu[VEC], rs[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) {
    s0[v] = s1[v] = 0.f;
    ga[v] = 1.f; be[v] = 0.f; mu[v] = 0.f; rs[v] = 1.f;
    if (MODE == 1 && active) {
      if (gamma) ga[v] = gamma[c + v];
      if (beta) be[v] = beta[c + v];
      mu[v] = mean[(int64_t)grp * C + c + v];
      rs[v] = rstd[(int64_t)grp * C + c + v];
    }
  }
  if (active) {
    // rows in batches of 4: all loads of a batch are issued before any arithmetic (memory-level parallelism; the
    // one-row-per-trip loop was latency-bound at ~1.5 TB/s)
    constexpr int UB = GG_BN_UB;
    const int64_t base = (int64_t)grp * rows_per_group;
    // a CTA owns a CONTIGUOUS chunk of rows (its warps stream consecutive 128-512 byte rows: whole DRAM pages), not every
    // gridDim.x-th row block (that layout touched a different page with every load: ~1.7 TB/s from HBM)
    const int64_t rows_per_cta = (rows_per_group + gridDim.x - 1) / gridDim.x;
    const int64_t r_begin = (int64_t)blockIdx.x * rows_per_cta, r_end = min(rows_per_group, r_begin + rows_per_cta);
    const int64_t rstep = ty_dim;
    for (int64_t r = r_begin + ty; r < r_end; r += UB * rstep) {
      float xv[UB][VEC], dv[UB][VEC];
#pragma unroll
      for (int ub = 0; ub < UB; ++ub) {
        const int64_t rr = r + ub * rstep;
        if (rr < r_end) {
          const int64_t off = (base + rr) * C + c;
          if (VEC == 4) {
            float4 t = ld4(x + off); xv[ub][0] = t.x; xv[ub][1] = t.y; xv[ub][2] = t.z; xv[ub][3] = t.w;
            if (MODE == 1) { float4 u = ld4(dy + off); dv[ub][0] = u.x; dv[ub][1] = u.y; dv[ub][2] = u.z; dv[ub][3] = u.w; }
          } else {
            xv[ub][0] = ldf(x + off);
            if (MODE == 1) dv[ub][0] = ldf(dy + off);
          }
        } else {
#pragma unroll
          for (int v = 0; v < VEC; ++v) { xv[ub][v] = (MODE == 1) ? mu[v] : 0.f; dv[ub][v] = 0.f; }   // contributes exactly 0
        }
      }
#pragma unroll
      for (int ub = 0; ub < UB; ++ub) {
        if (MODE == 1) {
          float xh[VEC], u[VEC];
#pragma unroll
          for (int v = 0; v < VEC; ++v) { xh[v] = (xv[ub][v] - mu[v]) * rs[v]; u[v] = fmaf(ga[v], xh[v], be[v]); }
          act_bwd_pre_vec<VEC>(dv[ub], u, act, act_param);
#pragma unroll
          for (int v = 0; v < VEC; ++v) { s0[v] += dv[ub][v]; s1[v] = fmaf(dv[ub][v], xh[v], s1[v]); }
        } else {
#pragma unroll
          for (int v = 0; v < VEC; ++v) {
            s0[v] += xv[ub][v];
            if (MODE == 0) s1[v] = fmaf(xv[ub][v], xv[ub][v], s1[v]);
          }
        }
      }
    }
  }
  // block reduction over ty
  __shared__ float red[2][CS_MAX_THREADS * VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) { red[0][threadIdx.x * VEC + v] = s0[v]; red[1][threadIdx.x * VEC + v] = s1[v]; }
  __syncthreads();
  if (ty == 0 && active) {
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      float a = 0.f, b = 0.f;
      for (int j = 0; j < ty_dim; ++j) { a += red[0][(j * tx_dim + tx) * VEC + v]; b += red[1][(j * tx_dim + tx) * VEC + v]; }
      const int rep = (int)(blockIdx.x % (unsigned)R);        // replicated accumulators (common.cuh: bn_replicas)
      atomicAdd(sums + bn_sum_index(rep, (int)gridDim.z, grp, 0, C, c + v), (double)a);
      if (MODE != 2) atomicAdd(sums + bn_sum_index(rep, (int)gridDim.z, grp, 1, C, c + v), (double)b);
    }
  }
}

__global__ void bn_infer_stats_kernel(const float* __restrict__ mm, const float* __restrict__ mv, float eps, int C,
                                      float* __restrict__ save_mean, float* __restrict__ save_rstd) {
  pdl_grid_sync();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  save_mean[c] = mm[c];
  save_rstd[c] = rsqrtf(mv[c] + eps);
}

// ---- train-mode apply: y = act((x - mean) * rstd * gamma + beta), statistics finalised in the same kernel ----------
// 2-D mapping (tx over channel vectors, ty over rows): a thread owns ONE channel vector for the whole kernel, so the
// per-channel constants live in registers and the row loop has no divisions.  Every thread derives mean / rstd of
// its channels from the fp64 sums (cheap, redundant); the first row-block also stores them for the backward pass and
// applies the EMA updates, sequentially over the row groups (ops.py:18-24: real batch first, then fake).
template <typename TX, typename TY, int VEC, int UB>
__global__ void __launch_bounds__(BN_THREADS)
bn_train_apply_kernel(const TX* __restrict__ x, TY* __restrict__ y, int64_t rows_per_group, int C, int groups,
                      const float* __restrict__ gamma, const float* __restrict__ beta, const double* __restrict__ sums, float eps,
                      float decay, float* __restrict__ moving_mean, float* __restrict__ moving_var, float* __restrict__ save_mean,
                      float* __restrict__ save_rstd, int act, float act_param, int tx_dim) {
  pdl_grid_sync();
  const int tx = threadIdx.x % tx_dim, ty = threadIdx.x / tx_dim, ty_dim = BN_THREADS / tx_dim;
  const int grp = blockIdx.z;
  const int c = (blockIdx.y * tx_dim + tx) * VEC;
  const bool active = c < C;
  const int R = bn_replicas(C, groups);
  const double inv = 1.0 / (double)rows_per_group;
  if (!active) return;
  const int64_t base = (int64_t)grp * rows_per_group;
  const int64_t rows_per_cta = (rows_per_group + gridDim.x - 1) / gridDim.x;      // contiguous row chunk per CTA (see colsum_kernel)
  const int64_t r_begin = (int64_t)blockIdx.x * rows_per_cta, r_end = min(rows_per_group, r_begin + rows_per_cta);
  const int64_t rstep = ty_dim;
  // The kernel is a chain of memory round trips (statistics, then one per load batch; ncu: SMs active 60 % of the launch,
  // DRAM at 20 %): the first batch of x is requested BEFORE the statistics are read, so that the two latencies overlap.
  float o[UB][VEC];
  int64_t r = r_begin + ty;
#pragma unroll
  for (int ub = 0; ub < UB; ++ub) {
    const int64_t rr = r + ub * rstep;
    if (rr < r_end) {
      const int64_t off = (base + rr) * C + c;
      if (VEC == 4) { float4 t = ld4(x + off); o[ub][0] = t.x; o[ub][1] = t.y; o[ub][2] = t.z; o[ub][3] = t.w; }
      else o[ub][0] = ldf(x + off);
    }
  }
  float mu[VEC], rs[VEC], ga[VEC], be[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) {
    const double m = bn_sum_read(sums, R, groups, grp, 0, C, c + v) * inv;
    double var = bn_sum_read(sums, R, groups, grp, 1, C, c + v) * inv - m * m;
    var = var < 0.0 ? 0.0 : var;
    mu[v] = (float)m;
    rs[v] = rsqrtf((float)var + eps);
    ga[v] = gamma ? __ldg(gamma + c + v) : 1.f;
    be[v] = beta ? __ldg(beta + c + v) : 0.f;
  }
  if (blockIdx.x == 0 && ty == 0) {
#pragma unroll
    for (int v = 0; v < VEC; ++v) { save_mean[(int64_t)grp * C + c + v] = mu[v]; save_rstd[(int64_t)grp * C + c + v] = rs[v]; }
    if (grp == 0 && (moving_mean || moving_var)) {
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        float mm = moving_mean ? moving_mean[c + v] : 0.f, mv = moving_var ? moving_var[c + v] : 0.f;
        for (int g = 0; g < groups; ++g) {
          const double m = bn_sum_read(sums, R, groups, g, 0, C, c + v) * inv;
          double var = bn_sum_read(sums, R, groups, g, 1, C, c + v) * inv - m * m;
          var = var < 0.0 ? 0.0 : var;
          // assign_moving_average(moving, batch, decay) = moving - (moving - batch) * (1 - decay)
          mm -= (mm - (float)m) * (1.f - decay);
          mv -= (mv - (float)var) * (1.f - decay);
        }
        if (moving_mean) moving_mean[c + v] = mm;
        if (moving_var) moving_var[c + v] = mv;
      }
    }
  }
  while (r < r_end) {
#pragma unroll
    for (int ub = 0; ub < UB; ++ub) {
      const int64_t rr = r + ub * rstep;
      if (rr < r_end) {
        const int64_t off = (base + rr) * C + c;
#pragma unroll
        for (int v = 0; v < VEC; ++v) o[ub][v] = fmaf((o[ub][v] - mu[v]) * rs[v], ga[v], be[v]);
        act_fwd_vec<VEC>(o[ub], act, act_param);
        if (VEC == 4) st4(y + off, make_float4(o[ub][0], o[ub][1], o[ub][2], o[ub][3]));
        else stf(y + off, o[ub][0]);
      }
    }
    r += UB * rstep;
#pragma unroll
    for (int ub = 0; ub < UB; ++ub) {
      const int64_t rr = r + ub * rstep;
      if (rr < r_end) {
        const int64_t off = (base + rr) * C + c;
        if (VEC == 4) { float4 t = ld4(x + off); o[ub][0] = t.x; o[ub][1] = t.y; o[ub][2] = t.z; o[ub][3] = t.w; }
        else o[ub][0] = ldf(x + off);
      }
    }
  }
}

// ---- backward apply (+ gamma / beta gradients from the same reductions) ----------------------------------------------
template <typename TX, typename TD, typename TO, int VEC, int UB>
__global__ void __launch_bounds__(BN_THREADS, 3)
bn_bwd_apply_kernel(const TX* __restrict__ x, const TD* __restrict__ dy, TO* __restrict__ dx, int64_t rows_per_group, int C,
                    int groups, const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ mean,
                    const float* __restrict__ rstd, const double* __restrict__ sums, float* __restrict__ dgamma,
                    float* __restrict__ dbeta, int act, float act_param, int train, int tx_dim) {
  pdl_grid_sync();
  const int tx = threadIdx.x % tx_dim, ty = threadIdx.x / tx_dim, ty_dim = BN_THREADS / tx_dim;
  const int grp = blockIdx.z;
  const int c = (blockIdx.y * tx_dim + tx) * VEC;
  const bool active = c < C;
  const int R = (sums != nullptr) ? bn_replicas(C, groups) : 1;
  const float invM = 1.f / (float)rows_per_group;
  if (!active) return;
  const int64_t base = (int64_t)grp * rows_per_group;
  const int64_t rows_per_cta = (rows_per_group + gridDim.x - 1) / gridDim.x;      // contiguous row chunk per CTA (see colsum_kernel)
  const int64_t r_begin = (int64_t)blockIdx.x * rows_per_cta, r_end = min(rows_per_group, r_begin + rows_per_cta);
  const int64_t rstep = ty_dim;
  // first batch of (x, dy) requested before the per-channel constants and reductions are read (see bn_train_apply_kernel)
  float xv[UB][VEC], dv[UB][VEC];
  int64_t r = r_begin + ty;
#pragma unroll
  for (int ub = 0; ub < UB; ++ub) {
    const int64_t rr = r + ub * rstep;
    if (rr < r_end) {
      const int64_t off = (base + rr) * C + c;
      if (VEC == 4) {
        float4 t = ld4(x + off); xv[ub][0] = t.x; xv[ub][1] = t.y; xv[ub][2] = t.z; xv[ub][3] = t.w;
        float4 u = ld4(dy + off); dv[ub][0] = u.x; dv[ub][1] = u.y; dv[ub][2] = u.z; dv[ub][3] = u.w;
      } else { xv[ub][0] = ldf(x + off); dv[ub][0] = ldf(dy + off); }
    }
  }
  float mu[VEC], rs[VEC], ga[VEC], be[VEC], sg[VEC], sgx[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) {
    ga[v] = gamma ? __ldg(gamma + c + v) : 1.f;
    be[v] = beta ? __ldg(beta + c + v) : 0.f;
    mu[v] = __ldg(mean + (int64_t)grp * C + c + v);
    rs[v] = __ldg(rstd + (int64_t)grp * C + c + v);
    sg[v] = train ? (float)bn_sum_read(sums, R, groups, grp, 0, C, c + v) * invM : 0.f;
    sgx[v] = train ? (float)bn_sum_read(sums, R, groups, grp, 1, C, c + v) * invM : 0.f;
  }
  if (blockIdx.x == 0 && ty == 0 && grp == 0 && (dgamma || dbeta)) {
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      double a = 0.0, b = 0.0;
      for (int g = 0; g < groups; ++g) { a += bn_sum_read(sums, R, groups, g, 0, C, c + v); b += bn_sum_read(sums, R, groups, g, 1, C, c + v); }
      if (dbeta) dbeta[c + v] += (float)a;
      if (dgamma) dgamma[c + v] += (float)b;
    }
  }
  while (r < r_end) {
#pragma unroll
    for (int ub = 0; ub < UB; ++ub) {
      const int64_t rr = r + ub * rstep;
      if (rr < r_end) {
        const int64_t off = (base + rr) * C + c;
        float xh[VEC], u[VEC], o[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) { xh[v] = (xv[ub][v] - mu[v]) * rs[v]; u[v] = fmaf(ga[v], xh[v], be[v]); }
        act_bwd_pre_vec<VEC>(dv[ub], u, act, act_param);
#pragma unroll
        for (int v = 0; v < VEC; ++v) o[v] = train ? ga[v] * rs[v] * (dv[ub][v] - sg[v] - xh[v] * sgx[v]) : ga[v] * rs[v] * dv[ub][v];
        if (VEC == 4) st4(dx + off, make_float4(o[0], o[1], o[2], o[3]));
        else stf(dx + off, o[0]);
      }
    }
    r += UB * rstep;
#pragma unroll
    for (int ub = 0; ub < UB; ++ub) {
      const int64_t rr = r + ub * rstep;
      if (rr < r_end) {
        const int64_t off = (base + rr) * C + c;
        if (VEC == 4) {
          float4 t = ld4(x + off); xv[ub][0] = t.x; xv[ub][1] = t.y; xv[ub][2] = t.z; xv[ub][3] = t.w;
          float4 u = ld4(dy + off); dv[ub][0] = u.x; dv[ub][1] = u.y; dv[ub][2] = u.z; dv[ub][3] = u.w;
        } else { xv[ub][0] = ldf(x + off); dv[ub][0] = ldf(dy + off); }
      }
    }
  }
}

// ---- backward, single pass (train mode): reductions + grid barrier + apply in ONE cooperative kernel ------------------
// The step's tensors are small enough (<= ~25 MB per layer) that a machine-filling grid holds its whole slice of
// x and dy in registers: each thread loads its rows once, contributes (sum g, sum g*xhat) through the block reduction
// and one fp64 atomic per channel per CTA, waits at the grid barrier, and finishes dx from registers.  Versus the
// two-kernel path (colsum + bwd_apply) this reads x / dy once and saves a launch per batch-norm layer per pass.
constexpr int BNF_ROWS = 8;       // rows held per thread

template <typename TX, typename TD, typename TO>
__global__ void __launch_bounds__(BN_THREADS)
bn_bwd_fused_kernel(const TX* __restrict__ x, const TD* __restrict__ dy, TO* __restrict__ dx, int64_t rows_per_group, int C,
                    int groups, const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ mean,
                    const float* __restrict__ rstd, double* __restrict__ sums, float* __restrict__ dgamma,
                    float* __restrict__ dbeta, int act, float act_param, int tx_dim) {
  namespace cg = cooperative_groups;
  constexpr int VEC = 4;
  const int tx = threadIdx.x % tx_dim, ty = threadIdx.x / tx_dim, ty_dim = BN_THREADS / tx_dim;
  const int grp = blockIdx.z;
  const int c = (blockIdx.y * tx_dim + tx) * VEC;
  const bool active = c < C;
  const int R = bn_replicas(C, groups);
  float mu[VEC], rs[VEC], ga[VEC], be[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) {
    ga[v] = (active && gamma) ? __ldg(gamma + c + v) : 1.f;
    be[v] = (active && beta) ? __ldg(beta + c + v) : 0.f;
    mu[v] = active ? __ldg(mean + (int64_t)grp * C + c + v) : 0.f;
    rs[v] = active ? __ldg(rstd + (int64_t)grp * C + c + v) : 1.f;
  }
  const int64_t base = (int64_t)grp * rows_per_group;
  const int64_t r0 = (int64_t)blockIdx.x * ty_dim + ty, rstep = (int64_t)gridDim.x * ty_dim;
  float xh[BNF_ROWS][VEC], g[BNF_ROWS][VEC];
  // all loads first (memory-level parallelism), then the arithmetic
#pragma unroll
  for (int k = 0; k < BNF_ROWS; ++k) {
    const int64_t r = r0 + k * rstep;
    if (active && r < rows_per_group) {
      const int64_t off = (base + r) * C + c;
      const float4 t = ld4(x + off), u = ld4(dy + off);
      xh[k][0] = t.x; xh[k][1] = t.y; xh[k][2] = t.z; xh[k][3] = t.w;
      g[k][0] = u.x; g[k][1] = u.y; g[k][2] = u.z; g[k][3] = u.w;
    } else {
#pragma unroll
      for (int v = 0; v < VEC; ++v) { xh[k][v] = mu[v]; g[k][v] = 0.f; }
    }
  }
  float s0[VEC] = {0.f, 0.f, 0.f, 0.f}, s1[VEC] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int k = 0; k < BNF_ROWS; ++k) {
    float u[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) { xh[k][v] = (xh[k][v] - mu[v]) * rs[v]; u[v] = fmaf(ga[v], xh[k][v], be[v]); }
    act_bwd_pre_vec<VEC>(g[k], u, act, act_param);
#pragma unroll
    for (int v = 0; v < VEC; ++v) { s0[v] += g[k][v]; s1[v] = fmaf(g[k][v], xh[k][v], s1[v]); }
  }
  __shared__ float red[2][BN_THREADS * VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) { red[0][threadIdx.x * VEC + v] = s0[v]; red[1][threadIdx.x * VEC + v] = s1[v]; }
  __syncthreads();
  if (ty == 0 && active) {
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      float a = 0.f, b = 0.f;
      for (int j = 0; j < ty_dim; ++j) { a += red[0][(j * tx_dim + tx) * VEC + v]; b += red[1][(j * tx_dim + tx) * VEC + v]; }
      const int rep = (int)(blockIdx.x % (unsigned)R);
      atomicAdd(sums + bn_sum_index(rep, groups, grp, 0, C, c + v), (double)a);
      atomicAdd(sums + bn_sum_index(rep, groups, grp, 1, C, c + v), (double)b);
    }
  }
  __threadfence();
  cg::this_grid().sync();
  if (!active) return;
  const float invM = 1.f / (float)rows_per_group;
  float sg[VEC], sgx[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) {
    sg[v] = (float)bn_sum_read(sums, R, groups, grp, 0, C, c + v) * invM;
    sgx[v] = (float)bn_sum_read(sums, R, groups, grp, 1, C, c + v) * invM;
  }
  if (blockIdx.x == 0 && ty == 0 && grp == 0 && (dgamma || dbeta)) {
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      double a = 0.0, b = 0.0;
      for (int gi = 0; gi < groups; ++gi) { a += bn_sum_read(sums, R, groups, gi, 0, C, c + v); b += bn_sum_read(sums, R, groups, gi, 1, C, c + v); }
      if (dbeta) dbeta[c + v] += (float)a;
      if (dgamma) dgamma[c + v] += (float)b;
    }
  }
#pragma unroll
  for (int k = 0; k < BNF_ROWS; ++k) {
    const int64_t r = r0 + k * rstep;
    if (r < rows_per_group) {
      float o[VEC];
#pragma unroll
      for (int v = 0; v < VEC; ++v) o[v] = ga[v] * rs[v] * (g[k][v] - sg[v] - xh[k][v] * sgx[v]);
      st4(dx + (base + r) * C + c, make_float4(o[0], o[1], o[2], o[3]));
    }
  }
}

__global__ void add_colsum_kernel(const double* __restrict__ sums, int C, float* __restrict__ db) {
  pdl_grid_sync();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) db[c] += (float)sums[c];
}

// ---- host -----------------------------------------------------------------------------
template <typename TX, typename TD, int MODE>
static void launch_colsum(const void* x, const void* dy, int64_t rpg, int C, int groups, const float* gamma, const float* beta,
                          const float* mean, const float* rstd, int act, float ap, double* sums, bool vec_ok, cudaStream_t st, int R) {
  const int threads = MODE == 1 ? std::min(512, colsum_threads()) : colsum_threads();
  ColGeom g = col_geom(rpg, C, groups, vec_ok, threads);
  dim3 grid(g.rblocks, g.cblocks, groups);
  if (g.vec == 4)
    Launch(grid, threads, 0, st)(colsum_kernel<TX, TD, 4, MODE>, (const TX*)x, (const TD*)dy, rpg, C, gamma, beta, mean, rstd, act, ap, sums, g.tx, R);
  else
    Launch(grid, threads, 0, st)(colsum_kernel<TX, TD, 1, MODE>, (const TX*)x, (const TD*)dy, rpg, C, gamma, beta, mean, rstd, act, ap, sums, g.tx, R);
}

// streaming (apply) kernels: same 2-D mapping, but sized to fill the machine (no atomics at the end)
static ColGeom apply_geom(int64_t rows_per_group, int C, int groups, bool vec_ok, int per_sm_dflt = 4) {
  ColGeom g = col_geom(rows_per_group, C, groups, vec_ok);
  // CTAs per SM of the streaming kernels: 256 threads x <= 64 registers -> 4 resident CTAs per SM; a grid of exactly one
  // resident wave (148 x 4) avoids the half-empty second wave the earlier 148 x 6 produced (GG_APPLY_PER_SM: A/B knob)
  static const int per_sm_env = [] { const char* v = getenv("GG_APPLY_PER_SM"); return (v && *v) ? atoi(v) : 0; }();
  const int per_sm = per_sm_env > 0 ? per_sm_env : per_sm_dflt;
  int64_t want = std::max<int64_t>(1, (148 * per_sm) / ((int64_t)g.cblocks * groups));
  int64_t maxr = ceil_div64(rows_per_group, (int64_t)g.ty * 2);  // >= 2 rows per thread
  g.rblocks = (int)std::max<int64_t>(1, std::min<int64_t>(want, maxr));
  return g;
}

static inline bool aligned16(const void* p) { return ((uintptr_t)p % 16) == 0; }
static inline int apply_blocks(int64_t total) { return (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div64(total, BN_THREADS), 148 * 8)); }

}  // namespace gg

using namespace gg;

extern "C" size_t gg_bn_workspace_bytes(int32_t C, int32_t groups) { return (size_t)bn_replicas(C, groups) * 2 * C * groups * sizeof(double); }

namespace gg {
// sums[groups][2][C] += (sum x, sum x^2) per row group -- the separate statistics pass (used when the producer
// kernel could not fuse it into its epilogue)
int bn_accumulate_stats(const void* x, int x_dt, int64_t rows, int C, int groups, double* sums, cudaStream_t st) {
  GG_REQUIRE(rows % groups == 0, GG_ERR_INVALID, "bn stats: rows not divisible by groups");
  const bool vec_ok = aligned16(x);
  GG_DISPATCH_DTYPE(x_dt, TX, (launch_colsum<TX, TX, 0>(x, nullptr, rows / groups, C, groups, nullptr, nullptr, nullptr, nullptr, 0, 0.f, sums, vec_ok, st, bn_replicas(C, groups))));
  return check_launch("bn_stats");
}
}  // namespace gg

static int bn_finalize_and_apply(const void* x, int32_t x_dt, void* y, int32_t y_dt, int64_t rpg, int32_t C, int32_t groups,
                                 const float* gamma, const float* beta, float* moving_mean, float* moving_var, float* save_mean,
                                 float* save_rstd, float eps, float decay, int32_t act, float act_param, const double* sums,
                                 cudaStream_t st);

// Same as gg_bn_fwd_train, but the per-group (sum, sum of squares) are already in `stats` (written by
// gg_conv_down_stats / gg_conv_up_stats): only finalize (+EMA) and the apply pass run.
extern "C" int gg_bn_fwd_train_stats(const void* x, int32_t x_dt, void* y, int32_t y_dt, int64_t rows, int32_t C, int32_t groups,
                                     const float* gamma, const float* beta, float* moving_mean, float* moving_var, float* save_mean,
                                     float* save_rstd, float eps, float decay, int32_t act, float act_param, const double* stats,
                                     void* stream) {
  GG_REQUIRE(x && y && save_mean && save_rstd && stats, GG_ERR_INVALID, "bn_fwd_train_stats: null pointer");
  GG_REQUIRE(groups >= 1 && rows > 0 && rows % groups == 0, GG_ERR_INVALID, "bn_fwd_train_stats: rows not divisible by groups");
  return bn_finalize_and_apply(x, x_dt, y, y_dt, rows / groups, C, groups, gamma, beta, moving_mean, moving_var, save_mean, save_rstd,
                               eps, decay, act, act_param, stats, (cudaStream_t)stream);
}

extern "C" int gg_bn_fwd_train(const void* x, int32_t x_dt, void* y, int32_t y_dt, int64_t rows, int32_t C, int32_t groups,
                               const float* gamma, const float* beta, float* moving_mean, float* moving_var, float* save_mean,
                               float* save_rstd, float eps, float decay, int32_t act, float act_param, void* ws, size_t ws_bytes,
                               void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  GG_REQUIRE(x && y && save_mean && save_rstd && ws, GG_ERR_INVALID, "bn_fwd_train: null pointer");
  GG_REQUIRE(groups >= 1 && rows > 0 && rows % groups == 0, GG_ERR_INVALID, "bn_fwd_train: rows %lld not divisible by groups %d", (long long)rows, groups);
  GG_REQUIRE(ws_bytes >= gg_bn_workspace_bytes(C, groups), GG_ERR_WORKSPACE, "bn_fwd_train: workspace too small");
  const int64_t rpg = rows / groups;
  double* sums = (double*)ws;
  cudaMemsetAsync(sums, 0, gg_bn_workspace_bytes(C, groups), st);
  const bool vec_ok = aligned16(x) && aligned16(y);
  GG_DISPATCH_DTYPE(x_dt, TX, (launch_colsum<TX, TX, 0>(x, nullptr, rpg, C, groups, nullptr, nullptr, nullptr, nullptr, 0, 0.f, sums, vec_ok, st, bn_replicas(C, groups))));
  int rc = check_launch("bn_stats");
  if (rc) return rc;
  return bn_finalize_and_apply(x, x_dt, y, y_dt, rpg, C, groups, gamma, beta, moving_mean, moving_var, save_mean, save_rstd, eps, decay,
                               act, act_param, sums, st);
}

static int bn_finalize_and_apply(const void* x, int32_t x_dt, void* y, int32_t y_dt, int64_t rpg, int32_t C, int32_t groups,
                                 const float* gamma, const float* beta, float* moving_mean, float* moving_var, float* save_mean,
                                 float* save_rstd, float eps, float decay, int32_t act, float act_param, const double* sums,
                                 cudaStream_t st) {
  const ColGeom g = apply_geom(rpg, C, groups, aligned16(x) && aligned16(y));
  dim3 grid(g.rblocks, g.cblocks, groups);
  // GG_APPLY_UB=8: 8 loads in flight per thread (97 registers -> 2 resident CTAs per SM: pair it with GG_APPLY_PER_SM=2); default 4
  static const int ub_env = [] { const char* v = getenv("GG_APPLY_UB"); return (v && *v) ? atoi(v) : 4; }();
  const bool ub8 = ub_env >= 8 && ceil_div64(rpg, (int64_t)g.rblocks * g.ty) > 4;
#define GG_APPLY(TX, TY)                                                                                              \
  do {                                                                                                                \
    if (g.vec == 4 && ub8) Launch(grid, BN_THREADS, 0, st)(bn_train_apply_kernel<TX, TY, 4, 8>, (const TX*)x, (TY*)y, rpg, C, groups, gamma, beta, sums, eps, decay, moving_mean, moving_var, save_mean, save_rstd, act, act_param, g.tx); \
    else if (g.vec == 4) Launch(grid, BN_THREADS, 0, st)(bn_train_apply_kernel<TX, TY, 4, 4>, (const TX*)x, (TY*)y, rpg, C, groups, gamma, beta, sums, eps, decay, moving_mean, moving_var, save_mean, save_rstd, act, act_param, g.tx); \
    else Launch(grid, BN_THREADS, 0, st)(bn_train_apply_kernel<TX, TY, 1, 4>, (const TX*)x, (TY*)y, rpg, C, groups, gamma, beta, sums, eps, decay, moving_mean, moving_var, save_mean, save_rstd, act, act_param, g.tx);          \
  } while (0)
  if (x_dt == GG_F32 && y_dt == GG_F32) GG_APPLY(float, float);
  else if (x_dt == GG_F32) GG_APPLY(float, bf16);
  else if (y_dt == GG_F32) GG_APPLY(bf16, float);
  else GG_APPLY(bf16, bf16);
  return check_launch("bn_train_apply");
}

extern "C" int gg_bn_infer_stats(const float* moving_mean, const float* moving_var, float eps, int32_t C, float* save_mean,
                                 float* save_rstd, void* stream) {
  GG_REQUIRE(moving_mean && moving_var && save_mean && save_rstd, GG_ERR_INVALID, "bn_infer_stats: null pointer");
  Launch(ceil_div(C, 128), 128, 0, (cudaStream_t)stream)(bn_infer_stats_kernel, moving_mean, moving_var, eps, C, save_mean, save_rstd);
  return check_launch("bn_infer_stats");
}

extern "C" int gg_bn_fwd_infer(const void* x, int32_t x_dt, void* y, int32_t y_dt, int64_t rows, int32_t C, const float* gamma,
                               const float* beta, const float* moving_mean, const float* moving_var, float eps, int32_t act,
                               float act_param, void* stream) {
  // inference statistics are folded on the fly: mean = moving_mean, rstd = rsqrt(moving_var+eps)
  // (bn_apply reads mean/rstd arrays, so stage them through a tiny kernel into caller memory is
  // avoided by a dedicated lambda kernel below)
  cudaStream_t st = (cudaStream_t)stream;
  GG_REQUIRE(x && y && moving_mean && moving_var, GG_ERR_INVALID, "bn_fwd_infer: null pointer");
  // Reuse bn_apply with rstd computed per element: implemented as a separate small-footprint kernel.
  extern int bn_infer_apply(const void*, int, void*, int, int64_t, int, const float*, const float*, const float*, const float*, float, int, float, cudaStream_t);
  return bn_infer_apply(x, x_dt, y, y_dt, rows, C, gamma, beta, moving_mean, moving_var, eps, act, act_param, st);
}

namespace gg {
template <typename TX, typename TY, int VEC>
__global__ void __launch_bounds__(BN_THREADS)
bn_infer_apply_kernel(const TX* __restrict__ x, TY* __restrict__ y, int64_t rows, int C, int lanes, const float* __restrict__ gamma,
                      const float* __restrict__ beta, const float* __restrict__ mm, const float* __restrict__ mv, float eps, int act,
                      float act_param) {
  pdl_grid_sync();
  const int64_t total = rows * lanes;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / lanes;
    const int c = (int)(i - r * lanes) * VEC;
    const int64_t off = r * C + c;
    float xv[VEC], o[VEC];
    if (VEC == 4) { float4 t = ld4(x + off); xv[0] = t.x; xv[1] = t.y; xv[2] = t.z; xv[3] = t.w; }
    else xv[0] = ldf(x + off);
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      const float ga = gamma ? __ldg(gamma + c + v) : 1.f, be = beta ? __ldg(beta + c + v) : 0.f;
      const float rs = rsqrtf(__ldg(mv + c + v) + eps);
      o[v] = fmaf((xv[v] - __ldg(mm + c + v)) * rs, ga, be);
    }
    act_fwd_vec<VEC>(o, act, act_param);
    if (VEC == 4) st4(y + off, make_float4(o[0], o[1], o[2], o[3]));
    else stf(y + off, o[0]);
  }
}
}  // namespace gg

int bn_infer_apply(const void* x, int x_dt, void* y, int y_dt, int64_t rows, int C, const float* gamma, const float* beta,
                   const float* mm, const float* mv, float eps, int act, float act_param, cudaStream_t st) {
  const bool vec_ok = aligned16(x) && aligned16(y);
  const int vec = (vec_ok && C % 4 == 0) ? 4 : 1;
  const int lanes = C / vec;
  dim3 grid(apply_blocks(rows * lanes));
#define GG_IAPPLY(TX, TY)                                                                                             \
  do {                                                                                                                \
    if (vec == 4) Launch(grid, BN_THREADS, 0, st)(bn_infer_apply_kernel<TX, TY, 4>, (const TX*)x, (TY*)y, rows, C, lanes, gamma, beta, mm, mv, eps, act, act_param); \
    else Launch(grid, BN_THREADS, 0, st)(bn_infer_apply_kernel<TX, TY, 1>, (const TX*)x, (TY*)y, rows, C, lanes, gamma, beta, mm, mv, eps, act, act_param);          \
  } while (0)
  if (x_dt == GG_F32 && y_dt == GG_F32) GG_IAPPLY(float, float);
  else if (x_dt == GG_F32) GG_IAPPLY(float, bf16);
  else if (y_dt == GG_F32) GG_IAPPLY(bf16, float);
  else GG_IAPPLY(bf16, bf16);
  return check_launch("bn_infer_apply");
}

namespace gg {
// Cooperative single-pass backward; *fused = 0 when the problem does not fit (rows per thread > BNF_ROWS at the
// co-resident grid size) -- the caller then runs the two-kernel path.
template <typename TX, typename TD, typename TO>
static int bn_bwd_fused_t(const void* x, const void* dy, void* dx, int64_t rpg, int C, int groups, const float* gamma, const float* beta,
                          const float* mean, const float* rstd, double* sums, float* dgamma, float* dbeta, int act, float ap,
                          cudaStream_t st, int* fused) {
  static int max_ctas = -1;                    // co-resident CTAs of this instantiation on the device
  static std::once_flag once;
  std::call_once(once, [] {
    int per_sm = 0, dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, bn_bwd_fused_kernel<TX, TD, TO>, BN_THREADS, 0) == cudaSuccess)
      max_ctas = per_sm * sms;
    (void)cudaGetLastError();
  });
  *fused = 0;
  if (max_ctas <= 0) return GG_OK;
  ColGeom g = col_geom(rpg, C, groups, true);
  const int64_t per_slice = (int64_t)g.cblocks * groups;
  const int64_t need = ceil_div64(rpg, (int64_t)g.ty * BNF_ROWS);        // row blocks so that a thread holds <= BNF_ROWS rows
  const int64_t cap = std::min<int64_t>(max_ctas, 148 * 4) / per_slice;   // row blocks that can be co-resident
  if (need > cap || need < 1) return GG_OK;
  // use the whole co-resident budget when the tensor is big enough (>= 2 rows per thread), else the minimum
  int64_t rblocks = std::max<int64_t>(need, std::min<int64_t>(cap, ceil_div64(rpg, (int64_t)g.ty * 2)));
  dim3 grid((unsigned)rblocks, (unsigned)g.cblocks, (unsigned)groups);
  cudaLaunchConfig_t cfg = cudaLaunchConfig_t();
  cfg.gridDim = grid; cfg.blockDim = dim3(BN_THREADS); cfg.dynamicSmemBytes = 0; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;
  attr[0].val.cooperative = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, bn_bwd_fused_kernel<TX, TD, TO>, (const TX*)x, (const TD*)dy, (TO*)dx, rpg, C, groups, gamma, beta,
                                     mean, rstd, sums, dgamma, dbeta, act, ap, g.tx);
  if (e != cudaSuccess) { (void)cudaGetLastError(); return GG_OK; }       // not launchable here: fall back
  *fused = 1;
  return check_launch("bn_bwd_fused");
}

static int bn_bwd_fused(const void* x, int x_dt, const void* dy, int dy_dt, void* dx, int dx_dt, int64_t rpg, int C, int groups,
                        const float* gamma, const float* beta, const float* mean, const float* rstd, double* sums, float* dgamma,
                        float* dbeta, int act, float ap, cudaStream_t st, int* fused) {
  *fused = 0;
  // Opt-in (GG_BN_FUSED=1): measured on B200 inside the captured step, the cooperative launch + grid barrier costs as
  // much as the second pass it saves (23 us per launch vs 12 + 7 us for colsum + apply at these sizes).
  if (!(getenv("GG_BN_FUSED") && getenv("GG_BN_FUSED")[0] == '1')) return GG_OK;
  // the combinations the models produce: fp32 pre-norm tensor, bf16 / fp32 incoming and outgoing gradients
  if (x_dt == GG_F32 && dy_dt == GG_BF16 && dx_dt == GG_BF16)
    return bn_bwd_fused_t<float, bf16, bf16>(x, dy, dx, rpg, C, groups, gamma, beta, mean, rstd, sums, dgamma, dbeta, act, ap, st, fused);
  if (x_dt == GG_F32 && dy_dt == GG_F32 && dx_dt == GG_F32)
    return bn_bwd_fused_t<float, float, float>(x, dy, dx, rpg, C, groups, gamma, beta, mean, rstd, sums, dgamma, dbeta, act, ap, st, fused);
  if (x_dt == GG_F32 && dy_dt == GG_F32 && dx_dt == GG_BF16)
    return bn_bwd_fused_t<float, float, bf16>(x, dy, dx, rpg, C, groups, gamma, beta, mean, rstd, sums, dgamma, dbeta, act, ap, st, fused);
  return GG_OK;
}
}  // namespace gg

extern "C" int gg_bn_bwd(const void* x, int32_t x_dt, const void* dy, int32_t dy_dt, void* dx, int32_t dx_dt, int64_t rows, int32_t C,
                         int32_t groups, const float* gamma, const float* beta, const float* save_mean, const float* save_rstd,
                         float* dgamma, float* dbeta, int32_t act, float act_param, int32_t train, void* ws, size_t ws_bytes,
                         void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  GG_REQUIRE(x && dy && dx && save_mean && save_rstd, GG_ERR_INVALID, "bn_bwd: null pointer");
  GG_REQUIRE(groups >= 1 && rows > 0 && rows % groups == 0, GG_ERR_INVALID, "bn_bwd: rows not divisible by groups");
  const int64_t rpg = rows / groups;
  double* sums = (double*)ws;
  const bool need_sums = train || dgamma || dbeta;
  const bool vec_ok = aligned16(x) && aligned16(dy) && aligned16(dx);
  int rc;
  if (need_sums) {
    GG_REQUIRE(ws && ws_bytes >= gg_bn_workspace_bytes(C, groups), GG_ERR_WORKSPACE, "bn_bwd: workspace too small");
    if (train != 2 && train != 3) cudaMemsetAsync(sums, 0, gg_bn_workspace_bytes(C, groups), st);   // train == 2: the caller zeroed ws (one memset per update for all layers)
    if (train == 3) goto apply;              // ws already holds the reductions (gg_conv_dgrad_bnbwd produced them with dy)
    if (train && vec_ok && C % 4 == 0) {
      int fused = 0;
      rc = bn_bwd_fused(x, x_dt, dy, dy_dt, dx, dx_dt, rpg, C, groups, gamma, beta, save_mean, save_rstd, sums, dgamma, dbeta, act, act_param, st, &fused);
      if (rc) return rc;
      if (fused) return GG_OK;
    }
#define GG_CS(TX, TD) launch_colsum<TX, TD, 1>(x, dy, rpg, C, groups, gamma, beta, save_mean, save_rstd, act, act_param, sums, vec_ok, st, bn_replicas(C, groups))
    if (x_dt == GG_F32 && dy_dt == GG_F32) GG_CS(float, float);
    else if (x_dt == GG_F32) GG_CS(float, bf16);
    else if (dy_dt == GG_F32) GG_CS(bf16, float);
    else GG_CS(bf16, bf16);
#undef GG_CS
    rc = check_launch("bn_bwd_reduce");
    if (rc) return rc;
  }
apply:
  const ColGeom g = apply_geom(rpg, C, groups, vec_ok, 3);      // bn_bwd_apply: 80 registers -> 3 resident CTAs per SM
  dim3 grid(g.rblocks, g.cblocks, groups);
#define GG_BA(TX, TD, TO)                                                                                                          \
  do {                                                                                                                             \
    if (g.vec == 4) Launch(grid, BN_THREADS, 0, st)(bn_bwd_apply_kernel<TX, TD, TO, 4, 4>, (const TX*)x, (const TD*)dy, (TO*)dx, rpg, C, groups, gamma, beta, save_mean, save_rstd, sums, dgamma, dbeta, act, act_param, train, g.tx); \
    else Launch(grid, BN_THREADS, 0, st)(bn_bwd_apply_kernel<TX, TD, TO, 1, 4>, (const TX*)x, (const TD*)dy, (TO*)dx, rpg, C, groups, gamma, beta, save_mean, save_rstd, sums, dgamma, dbeta, act, act_param, train, g.tx);          \
  } while (0)
  const int key = (x_dt == GG_BF16 ? 4 : 0) | (dy_dt == GG_BF16 ? 2 : 0) | (dx_dt == GG_BF16 ? 1 : 0);
  switch (key) {
    case 0: GG_BA(float, float, float); break;
    case 1: GG_BA(float, float, bf16); break;
    case 2: GG_BA(float, bf16, float); break;
    case 3: GG_BA(float, bf16, bf16); break;
    case 4: GG_BA(bf16, float, float); break;
    case 5: GG_BA(bf16, float, bf16); break;
    case 6: GG_BA(bf16, bf16, float); break;
    default: GG_BA(bf16, bf16, bf16); break;
  }
  return check_launch("bn_bwd_apply");
}

extern "C" int gg_bias_grad(const void* dy, int32_t dy_dt, float* db, int64_t rows, int32_t C, void* stream) {
  // column sums through the same coalesced reduction; needs a small fp64 scratch: use a
  // per-call static device buffer is not allowed (no allocation) -> accumulate via a
  // dedicated float-atomic kernel instead.
  extern int bias_grad_impl(const void*, int, float*, int64_t, int, cudaStream_t);
  GG_REQUIRE(dy && db && rows > 0 && C > 0, GG_ERR_INVALID, "bias_grad: bad argument");
  return bias_grad_impl(dy, dy_dt, db, rows, C, (cudaStream_t)stream);
}

namespace gg {
template <typename TD, int VEC>
__global__ void __launch_bounds__(CS_MAX_THREADS)
bias_grad_kernel(const TD* __restrict__ dy, float* __restrict__ db, int64_t rows, int C, int tx_dim) {
  pdl_grid_sync();
  const int tx = threadIdx.x % tx_dim, ty = threadIdx.x / tx_dim, ty_dim = (int)blockDim.x / tx_dim;
  const int c = (blockIdx.y * tx_dim + tx) * VEC;
  float s[VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) s[v] = 0.f;
  if (c < C) {
    constexpr int UB = GG_BN_UB;                            // 4 rows per trip, loads first (see colsum_kernel)
    const int64_t rows_per_cta = (rows + gridDim.x - 1) / gridDim.x;              // contiguous row chunk per CTA (see colsum_kernel)
    const int64_t r_begin = (int64_t)blockIdx.x * rows_per_cta, r_end = min(rows, r_begin + rows_per_cta);
    const int64_t rstep = ty_dim;
    for (int64_t r = r_begin + ty; r < r_end; r += UB * rstep) {
      float t[UB][VEC];
#pragma unroll
      for (int ub = 0; ub < UB; ++ub) {
        const int64_t rr = r + ub * rstep;
        if (rr < r_end) {
          if (VEC == 4) { float4 q = ld4(dy + rr * C + c); t[ub][0] = q.x; t[ub][1] = q.y; t[ub][2] = q.z; t[ub][3] = q.w; }
          else t[ub][0] = ldf(dy + rr * C + c);
        } else {
#pragma unroll
          for (int v = 0; v < VEC; ++v) t[ub][v] = 0.f;
        }
      }
#pragma unroll
      for (int ub = 0; ub < UB; ++ub)
#pragma unroll
        for (int v = 0; v < VEC; ++v) s[v] += t[ub][v];
    }
  }
  __shared__ float red[CS_MAX_THREADS * VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) red[threadIdx.x * VEC + v] = s[v];
  __syncthreads();
  if (ty == 0 && c < C) {
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      float a = 0.f;
      for (int j = 0; j < ty_dim; ++j) a += red[(j * tx_dim + tx) * VEC + v];
      atomicAdd(db + c + v, a);
    }
  }
}
}  // namespace gg

int bias_grad_impl(const void* dy, int dy_dt, float* db, int64_t rows, int C, cudaStream_t st) {
  const int threads = colsum_threads();
  ColGeom g = col_geom(rows, C, 1, aligned16(dy), threads);
  dim3 grid(g.rblocks, g.cblocks, 1);
  if (dy_dt == GG_F32) {
    if (g.vec == 4) Launch(grid, threads, 0, st)(bias_grad_kernel<float, 4>, (const float*)dy, db, rows, C, g.tx);
    else Launch(grid, threads, 0, st)(bias_grad_kernel<float, 1>, (const float*)dy, db, rows, C, g.tx);
  } else {
    if (g.vec == 4) Launch(grid, threads, 0, st)(bias_grad_kernel<bf16, 4>, (const bf16*)dy, db, rows, C, g.tx);
    else Launch(grid, threads, 0, st)(bias_grad_kernel<bf16, 1>, (const bf16*)dy, db, rows, C, g.tx);
  }
  return check_launch("bias_grad");
}

// ops.get_std (ops.py:125-128): column statistics over the batch axis, then mean over features.
namespace gg {
__global__ void get_std_final_kernel(const double* __restrict__ sums, int64_t B, int64_t F, float* __restrict__ out) {
  pdl_grid_sync();
  __shared__ double red[256];
  double acc = 0.0;
  for (int64_t f = threadIdx.x; f < F; f += blockDim.x) {
    const double m = sums[f] / (double)B;
    double var = sums[F + f] / (double)B - m * m;
    acc += var < 0.0 ? 0.0 : var;
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = (float)sqrt(red[0] / (double)F);
}
}  // namespace gg

extern "C" int gg_get_std(const void* x, int32_t x_dt, int64_t B, int64_t F, float* out, void* ws, size_t ws_bytes, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  GG_REQUIRE(x && out && ws && B > 0 && F > 0 && F < (1ll << 31), GG_ERR_INVALID, "get_std: bad argument");
  GG_REQUIRE(ws_bytes >= (size_t)2 * F * sizeof(double), GG_ERR_WORKSPACE, "get_std: workspace too small");
  double* sums = (double*)ws;
  cudaMemsetAsync(sums, 0, (size_t)2 * F * sizeof(double), st);
  GG_DISPATCH_DTYPE(x_dt, TX, (launch_colsum<TX, TX, 0>(x, nullptr, B, (int)F, 1, nullptr, nullptr, nullptr, nullptr, 0, 0.f, sums, aligned16(x), st, 1)));
  int rc = check_launch("get_std_stats");
  if (rc) return rc;
  Launch(1, 256, 0, st)(get_std_final_kernel, sums, B, F, out);
  return check_launch("get_std_final");
}
