// Discriminator loss head as two launches (was eight plus a batch-norm reduction pass):
//   model.py:277   h4 = linear(reshape(h3, [B, -1]), 1, 'd_h3_lin')            (z_model_lib.py:416 dvideo_h4 likewise)
//   model.py:121-131  d_loss_real / d_loss_fake / g_loss = reduce_mean(sigmoid_cross_entropy_with_logits(h4, 1 or 0))
// forward : logits[r] = h[r, :] . w + b  (one CTA per row), then the CTA that takes the last ticket evaluates every loss
//           segment and writes d loss / d logits -- same formula and summation order as sigmoid_ce_kernel;
// backward: ONE pass over h (a CTA owns 64 columns and all rows): dW[k] += sum_r dl[r] h[r,k], db += sum_r dl[r],
//           dh[r,k] = dl[r] w[k], and -- when h is the output of a train-mode batch norm -- that batch norm's backward
//           reductions (sum g, sum g * xhat per channel = column % C and row group) from the dh values still in registers,
//           so gg_bn_bwd(train = 3) runs its apply pass only.
// Latency-bound (2-4 MB, L2-resident): what is saved is launches and dependent round trips, not bytes.
#include "common.cuh"

namespace gg {

constexpr int LH_THREADS = 256;
constexpr int LH_MAX_SEGS = 4;

struct LhSegs {
  int n;
  int begin[LH_MAX_SEGS], end[LH_MAX_SEGS];
  float target[LH_MAX_SEGS], weight[LH_MAX_SEGS];
};

template <typename TX>
__global__ void __launch_bounds__(LH_THREADS)
loss_head_fwd_kernel(const TX* __restrict__ h, const float* __restrict__ w, const float* __restrict__ bias, int rows, int in_dim, LhSegs segs,
                     float* logits, float* __restrict__ parts, float* __restrict__ dlogits, unsigned int* ticket) {
  pdl_grid_sync();
  const int r = blockIdx.x;
  const TX* xr = h + (int64_t)r * in_dim;
  float acc = 0.f;
#pragma unroll 8
  for (int k = threadIdx.x * 4; k < in_dim; k += LH_THREADS * 4) {      // same trip order as skinny_fwd_kernel: identical logits
    const float4 xv = ld4(xr + k), wv = ld4(w + k);
    acc = fmaf(xv.x, wv.x, fmaf(xv.y, wv.y, fmaf(xv.z, wv.z, fmaf(xv.w, wv.w, acc))));
  }
  __shared__ float red[LH_THREADS];
  __shared__ bool last;
  const float s = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = bias ? bias[0] : 0.f;
    for (int q = 0; q < LH_THREADS / 32; ++q) t += red[q];
    logits[r] = t;
    __threadfence();
    const unsigned int prev = atomicAdd(ticket, 1u);
    last = (prev == (unsigned int)rows - 1u);
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  // the last CTA: every segment's mean cross-entropy (tf.nn.sigmoid_cross_entropy_with_logits: max(x,0) - x*z + log(1+exp(-|x|)))
  float total = 0.f;
  for (int i = 0; i < segs.n; ++i) {
    const int a = segs.begin[i], n = segs.end[i] - segs.begin[i];
    const float tg = segs.target[i], wt = segs.weight[i], inv = 1.f / (float)n;
    float acc2 = 0.f;
    for (int j = threadIdx.x; j < n; j += LH_THREADS) {
      const float x = __ldcg(logits + a + j);
      acc2 += fmaxf(x, 0.f) - x * tg + log1pf(expf(-fabsf(x)));
      if (dlogits) dlogits[a + j] = wt * inv * (1.f / (1.f + expf(-x)) - tg);
    }
    __syncthreads();
    red[threadIdx.x] = acc2;
    __syncthreads();
    for (int st = LH_THREADS / 2; st > 0; st >>= 1) {
      if (threadIdx.x < st) red[threadIdx.x] += red[threadIdx.x + st];
      __syncthreads();
    }
    const float part = wt * red[0] * inv;
    if (threadIdx.x == 0) parts[1 + i] = part;
    total = (i == 0) ? (0.f + part) : (total + part);
  }
  if (threadIdx.x == 0) {
    parts[0] = total;
    *ticket = 0u;                 // handed back zeroed for the next launch
  }
}

// 8 consecutive elements of a row (16 B of bf16, 32 B of fp32)
__device__ __forceinline__ void ld8(const float* p, float (&v)[8]) {
  const float4 a = ld4(p), b = ld4(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void ld8(const bf16* p, float (&v)[8]) {
  const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
  const __nv_bfloat162* q = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) { const float2 f = __bfloat1622float2(q[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
}
// store 8 values; v is replaced by what was stored (the rounded values when the tensor is bf16)
__device__ __forceinline__ void st8_round(float* p, float (&v)[8]) {
  st4(p, make_float4(v[0], v[1], v[2], v[3]));
  st4(p + 4, make_float4(v[4], v[5], v[6], v[7]));
}
__device__ __forceinline__ void st8_round(bf16* p, float (&v)[8]) {
  uint4 u;
  __nv_bfloat162* q = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    q[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    const float2 f = __bfloat1622float2(q[i]);
    v[2 * i] = f.x; v[2 * i + 1] = f.y;
  }
  *reinterpret_cast<uint4*>(p) = u;
}

struct LhBn {                // the batch norm that produced h (all NULL / 0 when there is none)
  const float* pre;          // fp32 pre-norm tensor, same [rows, in_dim] layout
  const float* mean;         // [groups][C]
  const float* rstd;
  const float* gamma;        // [C] or NULL
  const float* beta;
  double* sums;              // [groups][2][C], zeroed by the caller
  int C, groups, act;
  float act_param;
};

constexpr int LH_COLS = 64;                  // columns per CTA
constexpr int LH_CV = LH_COLS / 8;           // 8 column vectors ...
constexpr int LH_RL = LH_THREADS / LH_CV;    // ... x 32 row lanes

template <typename TX, typename TO>
__global__ void __launch_bounds__(LH_THREADS)
loss_head_bwd_kernel(const TX* __restrict__ h, const float* __restrict__ dl, const float* __restrict__ w, int rows, int in_dim,
                     float* __restrict__ dW, float* __restrict__ dbias, TO* __restrict__ dh, LhBn bn) {
  pdl_grid_sync();
  const int cv = threadIdx.x % LH_CV, rl = threadIdx.x / LH_CV;
  const int k0 = blockIdx.x * LH_COLS + cv * 8;
  __shared__ float red[3][LH_RL][LH_COLS + 1];
  float wv[8];
  ld8(w + k0, wv);
  float dw[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) dw[i] = 0.f;
  const bool has_bn = bn.sums != nullptr;
  const int groups = has_bn ? bn.groups : 1;
  const int rpg = rows / groups;
  const int c0 = has_bn ? (k0 % bn.C) : 0;
  float ga[8], be[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    ga[i] = (has_bn && bn.gamma) ? __ldg(bn.gamma + c0 + i) : 1.f;
    be[i] = (has_bn && bn.beta) ? __ldg(bn.beta + c0 + i) : 0.f;
  }
  float dbacc = 0.f;
  for (int g = 0; g < groups; ++g) {
    float mu[8], rs[8], s0[8], s1[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      s0[i] = s1[i] = 0.f;
      mu[i] = has_bn ? __ldg(bn.mean + (int64_t)g * bn.C + c0 + i) : 0.f;
      rs[i] = has_bn ? __ldg(bn.rstd + (int64_t)g * bn.C + c0 + i) : 1.f;
    }
    for (int r = g * rpg + rl; r < (g + 1) * rpg; r += LH_RL) {
      const int64_t off = (int64_t)r * in_dim + k0;
      float hv[8], pv[8], dv[8];
      ld8(h + off, hv);
      if (has_bn) ld8(bn.pre + off, pv);
      const float d = __ldg(dl + r);
      if (blockIdx.x == 0 && cv == 0) dbacc += d;
#pragma unroll
      for (int i = 0; i < 8; ++i) { dw[i] = fmaf(d, hv[i], dw[i]); dv[i] = d * wv[i]; }
      if (dh) st8_round(dh + off, dv);
      if (has_bn) {
        float xh[8], u[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { xh[i] = (pv[i] - mu[i]) * rs[i]; u[i] = fmaf(ga[i], xh[i], be[i]); }
        act_bwd_pre_vec<8>(dv, u, bn.act, bn.act_param);
#pragma unroll
        for (int i = 0; i < 8; ++i) { s0[i] += dv[i]; s1[i] = fmaf(dv[i], xh[i], s1[i]); }
      }
    }
    if (has_bn) {
      __syncthreads();
#pragma unroll
      for (int i = 0; i < 8; ++i) { red[1][rl][cv * 8 + i] = s0[i]; red[2][rl][cv * 8 + i] = s1[i]; }
      __syncthreads();
      if (threadIdx.x < 2 * LH_COLS) {
        const int which = threadIdx.x / LH_COLS, c = threadIdx.x % LH_COLS;
        float a = 0.f;
        for (int j = 0; j < LH_RL; ++j) a += red[1 + which][j][c];
        atomicAdd(bn.sums + bn_sum_index(0, groups, g, which, bn.C, (blockIdx.x * LH_COLS + c) % bn.C), (double)a);
      }
    }
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 8; ++i) red[0][rl][cv * 8 + i] = dw[i];
  if (blockIdx.x == 0 && cv == 0) red[1][rl][0] = dbacc;
  __syncthreads();
  if (threadIdx.x < LH_COLS && dW) {
    float a = 0.f;
    for (int j = 0; j < LH_RL; ++j) a += red[0][j][threadIdx.x];
    dW[blockIdx.x * LH_COLS + threadIdx.x] += a;                // this CTA is the only writer of its columns
  }
  if (blockIdx.x == 0 && threadIdx.x == LH_COLS && dbias) {
    float a = 0.f;
    for (int j = 0; j < LH_RL; ++j) a += red[1][j][0];
    dbias[0] += a;
  }
}

static inline bool lh_al16(const void* p) { return ((uintptr_t)p % 16) == 0; }

}  // namespace gg

using namespace gg;

extern "C" int gg_loss_head_ok(int32_t rows, int32_t in_dim, int32_t nsegs) {
  return rows > 0 && in_dim > 0 && in_dim % LH_COLS == 0 && nsegs >= 1 && nsegs <= LH_MAX_SEGS;
}

extern "C" int gg_loss_head_fwd(const void* h, int32_t h_dt, const float* w, const float* bias, int32_t rows, int32_t in_dim,
                                const int32_t* seg_begin, const int32_t* seg_end, const float* seg_target, const float* seg_weight, int32_t nsegs,
                                float* logits, float* parts, float* dlogits, void* ticket, void* stream) {
  GG_REQUIRE(h && w && logits && parts && ticket && seg_begin && seg_end && seg_target && seg_weight, GG_ERR_INVALID, "loss_head_fwd: null pointer");
  GG_REQUIRE(gg_loss_head_ok(rows, in_dim, nsegs) && lh_al16(h) && lh_al16(w), GG_ERR_UNSUPPORTED,
             "loss_head_fwd: needs in_dim %% 64 == 0, 1..4 segments, 16-byte aligned tensors (rows=%d in_dim=%d nsegs=%d)", rows, in_dim, nsegs);
  LhSegs s;
  s.n = nsegs;
  for (int i = 0; i < nsegs; ++i) {
    GG_REQUIRE(seg_begin[i] >= 0 && seg_end[i] > seg_begin[i] && seg_end[i] <= rows, GG_ERR_INVALID, "loss_head_fwd: segment %d out of range", i);
    s.begin[i] = seg_begin[i]; s.end[i] = seg_end[i]; s.target[i] = seg_target[i]; s.weight[i] = seg_weight[i];
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (h_dt == GG_BF16)
    Launch(rows, LH_THREADS, 0, st)(loss_head_fwd_kernel<bf16>, (const bf16*)h, w, bias, rows, in_dim, s, logits, parts, dlogits, (unsigned int*)ticket);
  else
    Launch(rows, LH_THREADS, 0, st)(loss_head_fwd_kernel<float>, (const float*)h, w, bias, rows, in_dim, s, logits, parts, dlogits, (unsigned int*)ticket);
  return check_launch("loss_head_fwd");
}

extern "C" int gg_loss_head_bwd(const void* h, int32_t h_dt, const float* dlogits, const float* w, int32_t rows, int32_t in_dim, float* dW,
                                float* dbias, void* dh, const float* pre, const float* save_mean, const float* save_rstd, const float* gamma,
                                const float* beta, int32_t act, float act_param, int32_t groups, int32_t C, double* sums, int32_t* fused,
                                void* stream) {
  GG_REQUIRE(h && dlogits && w && rows > 0 && in_dim > 0, GG_ERR_INVALID, "loss_head_bwd: bad argument");
  GG_REQUIRE(in_dim % LH_COLS == 0 && lh_al16(h) && lh_al16(w) && (!dh || lh_al16(dh)) && (!dW || lh_al16(dW)), GG_ERR_UNSUPPORTED,
             "loss_head_bwd: needs in_dim %% 64 == 0 and 16-byte aligned tensors (in_dim=%d)", in_dim);
  LhBn bn = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0, 1, GG_ACT_NONE, 0.f};
  const bool want_bn = pre && save_mean && save_rstd && sums && dh;
  // channel = column % C must put a thread's 8 columns in ONE period, and the row groups must be whole
  const bool bn_ok = want_bn && C > 0 && C % 8 == 0 && in_dim % C == 0 && groups >= 1 && rows % groups == 0 && lh_al16(pre);
  if (bn_ok) { bn.pre = pre; bn.mean = save_mean; bn.rstd = save_rstd; bn.gamma = gamma; bn.beta = beta; bn.sums = sums; bn.C = C; bn.groups = groups; bn.act = act; bn.act_param = act_param; }
  if (fused) *fused = bn_ok ? 1 : 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int blocks = in_dim / LH_COLS;
  if (h_dt == GG_BF16)
    Launch(blocks, LH_THREADS, 0, st)(loss_head_bwd_kernel<bf16, bf16>, (const bf16*)h, dlogits, w, rows, in_dim, dW, dbias, (bf16*)dh, bn);
  else
    Launch(blocks, LH_THREADS, 0, st)(loss_head_bwd_kernel<float, float>, (const float*)h, dlogits, w, rows, in_dim, dW, dbias, (float*)dh, bn);
  return check_launch("loss_head_bwd");
}
