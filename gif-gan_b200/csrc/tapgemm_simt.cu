// SIMT fp32-accumulate implicit-GEMM kernels for the strided-conv relation of
// include/gifgan.h (gg_conv_down / gg_conv_up / gg_conv_wgrad) and for linear layers.
//
// These are the "fp32 mode" of the framework (1e-4 parity against the oracle; the
// reference computes in fp32: /root/reference/models/recurrent_z/ops.py:57,70,86,115)
// and, in bf16 mode, the kernels for the 3-channel image-side layers (d_h0_conv, g_h4)
// whose arithmetic intensity (~31 flop/B) makes them HBM-bound, not tensor-bound.
//
// Formulation ("tap GEMM"):  out[m, n] = sum_{tap t} sum_k A[pix(m) * S + off_t, k] * W_t[k, n]
//   down: m over the small grid, A = large,  S = stride, off_t = t - pad,  W_t = w[t] ([C][K], n contiguous)
//   up  : one launch covers all stride^ndim output-parity classes (blockIdx.z); for class a,
//         m over {i : i % s == a}, A = small, S = 1, off_t = (a + pad - t)/s for the taps with
//         (a + pad - t) % s == 0, W_t[k, n] = w[t][n][k] (k contiguous).  No stride-inserted zeros.
// wgrad:  dw[t][c][k] += sum_o large[s*o + t - p, c] * small[o, k]   (split over o, fp32 atomics)
#include <algorithm>

#include "common.cuh"

namespace gg {

constexpr int MAX_TAPS = 32;
constexpr int MAX_CLASSES = 8;
constexpr int BM = 64, BN = 64, BK = 16, LDS = 68;

struct Tap {
  int16_t od, oh, ow, pad_;
  int32_t wofs;
};
struct ClassInfo {
  int Md, Mh, Mw;        // M-grid of this class (per batch item)
  int od0, oh0, ow0;     // output coordinate = m * os + o0
  int tap_begin, tap_end;
};
struct PixGemmParams {
  int Mn;
  int Ad, Ah, Aw, Ac;    // A tensor dims (per batch item) and channels (= reduction length per tap)
  int Sd, Sh, Sw;        // A coordinate = m * S + off
  int Nc;                // output channels
  int ldbk, ldbn;        // W_t(k, n) = w[wofs + k*ldbk + n*ldbn]
  int OD, OH, OW;        // output tensor dims (per batch item)
  int osd, osh, osw;     // output coordinate stride
  int act;
  float act_param;
  int nclasses;
  ClassInfo cls[MAX_CLASSES];
  Tap taps[MAX_TAPS];
};

template <typename TA, typename TO, bool B_KCONTIG, bool VEC_A, bool VEC_B, bool VEC_O>
__global__ void __launch_bounds__(256)
pixgemm_kernel(const TA* __restrict__ A, const float* __restrict__ Wt, const float* __restrict__ bias,
               TO* __restrict__ out, const __grid_constant__ PixGemmParams p) {
  pdl_grid_sync();
  const ClassInfo ci = p.cls[blockIdx.z];
  const int64_t Mper = (int64_t)ci.Md * ci.Mh * ci.Mw;
  const int64_t M = Mper * p.Mn;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  if (m0 >= M) return;
  const int n0 = blockIdx.y * BN;

  __shared__ __align__(16) float As[BK][LDS];
  __shared__ __align__(16) float Bs[BK][LDS];

  const int tid = threadIdx.x;
  // ---- A-load role: one pixel row, 4 consecutive k
  const int a_row = tid >> 2, a_kq = (tid & 3) * 4;
  const int64_t am = m0 + a_row;
  const bool a_in = am < M;
  int a_nb = 0, a_md = 0, a_mh = 0, a_mw = 0;
  if (a_in) {
    int64_t r = am;
    a_nb = (int)(r / Mper); r -= (int64_t)a_nb * Mper;
    a_md = (int)(r / (ci.Mh * ci.Mw)); r -= (int64_t)a_md * ci.Mh * ci.Mw;
    a_mh = (int)(r / ci.Mw); a_mw = (int)(r - (int64_t)a_mh * ci.Mw);
  }
  // ---- B-load role
  const int b_k = B_KCONTIG ? (tid & 3) * 4 : (tid >> 4);
  const int b_n = B_KCONTIG ? (tid >> 2) : (tid & 15) * 4;
  // ---- compute role
  const int ty = tid >> 4, tx = tid & 15;

  // two-level accumulation: `part` collects FOLD k-chunks (128 products), then folds into `acc`, so the
  // rounding error of a K = 25*512 reduction grows like a blocked sum, not a 12800-term running sum
  float acc[4][4], part[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) { acc[i][j] = 0.f; part[i][j] = 0.f; }

  const int KC = (p.Ac + BK - 1) / BK;
  const int ntap = ci.tap_end - ci.tap_begin;
  const int iters = ntap * KC;

  float a_reg[4], b_reg[4];
  auto load_tiles = [&](int it) {
    const int t = ci.tap_begin + it / KC;
    const int kc = (it % KC) * BK;
    const Tap tap = p.taps[t];
    // A
    {
      const int ad = a_md * p.Sd + tap.od, ah = a_mh * p.Sh + tap.oh, aw = a_mw * p.Sw + tap.ow;
      const bool ok = a_in && ad >= 0 && ad < p.Ad && ah >= 0 && ah < p.Ah && aw >= 0 && aw < p.Aw;
      const int k = kc + a_kq;
      a_reg[0] = a_reg[1] = a_reg[2] = a_reg[3] = 0.f;
      if (ok) {
        const TA* src = A + ((((int64_t)a_nb * p.Ad + ad) * p.Ah + ah) * p.Aw + aw) * p.Ac + k;
        if (VEC_A) {
          if (k < p.Ac) { float4 v = ld4(src); a_reg[0] = v.x; a_reg[1] = v.y; a_reg[2] = v.z; a_reg[3] = v.w; }
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i) if (k + i < p.Ac) a_reg[i] = ldf(src + i);
        }
      }
    }
    // B
    {
      b_reg[0] = b_reg[1] = b_reg[2] = b_reg[3] = 0.f;
      if (B_KCONTIG) {
        const int k = kc + b_k, n = n0 + b_n;
        if (n < p.Nc) {
          const float* src = Wt + tap.wofs + (int64_t)n * p.ldbn + k;
          if (VEC_B) {
            if (k < p.Ac) { float4 v = ld4(src); b_reg[0] = v.x; b_reg[1] = v.y; b_reg[2] = v.z; b_reg[3] = v.w; }
          } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) if (k + i < p.Ac) b_reg[i] = __ldg(src + i);
          }
        }
      } else {
        const int k = kc + b_k, n = n0 + b_n;
        if (k < p.Ac) {
          const float* src = Wt + tap.wofs + (int64_t)k * p.ldbk + n;
          if (VEC_B) {
            if (n < p.Nc) { float4 v = ld4(src); b_reg[0] = v.x; b_reg[1] = v.y; b_reg[2] = v.z; b_reg[3] = v.w; }
          } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) if (n + i < p.Nc) b_reg[i] = __ldg(src + i);
          }
        }
      }
    }
  };
  auto store_tiles = [&]() {
#pragma unroll
    for (int i = 0; i < 4; ++i) As[a_kq + i][a_row] = a_reg[i];
    if (B_KCONTIG) {
#pragma unroll
      for (int i = 0; i < 4; ++i) Bs[b_k + i][b_n] = b_reg[i];
    } else {
      *reinterpret_cast<float4*>(&Bs[b_k][b_n]) = make_float4(b_reg[0], b_reg[1], b_reg[2], b_reg[3]);
    }
  };

  if (iters > 0) load_tiles(0);
  for (int it = 0; it < iters; ++it) {
    store_tiles();
    __syncthreads();
    if (it + 1 < iters) load_tiles(it + 1);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) part[i][j] = fmaf(av[i], bv[j], part[i][j]);
    }
    if ((it & 7) == 7 || it == iters - 1) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { acc[i][j] += part[i][j]; part[i][j] = 0.f; }
    }
    __syncthreads();
  }

  // ---- epilogue
  const int n = n0 + tx * 4;
  float bv[4] = {0.f, 0.f, 0.f, 0.f};
  if (bias != nullptr) {
#pragma unroll
    for (int j = 0; j < 4; ++j) if (n + j < p.Nc) bv[j] = __ldg(bias + n + j);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t m = m0 + ty * 4 + i;
    if (m >= M) continue;
    int64_t r = m;
    const int nb = (int)(r / Mper); r -= (int64_t)nb * Mper;
    const int md = (int)(r / (ci.Mh * ci.Mw)); r -= (int64_t)md * ci.Mh * ci.Mw;
    const int mh = (int)(r / ci.Mw), mw = (int)(r - (int64_t)mh * ci.Mw);
    const int od = md * p.osd + ci.od0, oh = mh * p.osh + ci.oh0, ow = mw * p.osw + ci.ow0;
    TO* dst = out + ((((int64_t)nb * p.OD + od) * p.OH + oh) * p.OW + ow) * p.Nc + n;
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = acc[i][j] + bv[j];
    act_fwd_vec<4>(v, p.act, p.act_param);
    if (VEC_O) {
      if (n < p.Nc) st4(dst, make_float4(v[0], v[1], v[2], v[3]));
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) if (n + j < p.Nc) stf(dst + j, v[j]);
    }
  }
}

// ------------------------------------------------------------------------------------
struct WgradParams {
  int N, D, H, W, C, Do, Ho, Wo, K;
  int kd, kh, kw, sd, sh, sw, pd, ph, pw;
  int splits;
  int64_t Mtot, chunk;
};

template <typename TL, typename TS, bool VEC_L, bool VEC_S>
__global__ void __launch_bounds__(256)
wgrad_kernel(const TL* __restrict__ large, const TS* __restrict__ small, float* __restrict__ dw,
             const __grid_constant__ WgradParams p) {
  pdl_grid_sync();
  const int c0 = blockIdx.x * BM, k0 = blockIdx.y * BN;
  const int tap = blockIdx.z / p.splits, split = blockIdx.z % p.splits;
  const int td = tap / (p.kh * p.kw), th = (tap / p.kw) % p.kh, tw = tap % p.kw;
  const int64_t mbeg = (int64_t)split * p.chunk;
  const int64_t mend = min(p.Mtot, mbeg + p.chunk);
  if (mbeg >= mend) return;

  __shared__ __align__(16) float As[BK][LDS];
  __shared__ __align__(16) float Bs[BK][LDS];
  const int tid = threadIdx.x;
  const int lp = tid >> 4, lq = (tid & 15) * 4;  // load role: pixel lp, channels lq..lq+3
  const int ty = tid >> 4, tx = tid & 15;

  float acc[4][4], part[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) { acc[i][j] = 0.f; part[i][j] = 0.f; }

  float a_reg[4], b_reg[4];
  const int64_t HoWo = (int64_t)p.Ho * p.Wo, DHW = (int64_t)p.Do * HoWo;
  auto load_tiles = [&](int64_t mb) {
    const int64_t m = mb + lp;
    a_reg[0] = a_reg[1] = a_reg[2] = a_reg[3] = 0.f;
    b_reg[0] = b_reg[1] = b_reg[2] = b_reg[3] = 0.f;
    if (m < mend) {
      int64_t r = m;
      const int nb = (int)(r / DHW); r -= (int64_t)nb * DHW;
      const int od = (int)(r / HoWo); r -= (int64_t)od * HoWo;
      const int oh = (int)(r / p.Wo), ow = (int)(r - (int64_t)oh * p.Wo);
      const int id = od * p.sd + td - p.pd, ih = oh * p.sh + th - p.ph, iw = ow * p.sw + tw - p.pw;
      if (id >= 0 && id < p.D && ih >= 0 && ih < p.H && iw >= 0 && iw < p.W) {
        const int c = c0 + lq;
        const TL* src = large + ((((int64_t)nb * p.D + id) * p.H + ih) * p.W + iw) * p.C + c;
        if (VEC_L) {
          if (c < p.C) { float4 v = ld4(src); a_reg[0] = v.x; a_reg[1] = v.y; a_reg[2] = v.z; a_reg[3] = v.w; }
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i) if (c + i < p.C) a_reg[i] = ldf(src + i);
        }
        // small is only needed where large is in range (the product is zero otherwise)
        const int k = k0 + lq;
        const TS* ss = small + m * p.K + k;
        if (VEC_S) {
          if (k < p.K) { float4 v = ld4(ss); b_reg[0] = v.x; b_reg[1] = v.y; b_reg[2] = v.z; b_reg[3] = v.w; }
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i) if (k + i < p.K) b_reg[i] = ldf(ss + i);
        }
      }
    }
  };

  load_tiles(mbeg);
  int it = 0;
  for (int64_t mb = mbeg; mb < mend; mb += BK, ++it) {
    *reinterpret_cast<float4*>(&As[lp][lq]) = make_float4(a_reg[0], a_reg[1], a_reg[2], a_reg[3]);
    *reinterpret_cast<float4*>(&Bs[lp][lq]) = make_float4(b_reg[0], b_reg[1], b_reg[2], b_reg[3]);
    __syncthreads();
    if (mb + BK < mend) load_tiles(mb + BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) part[i][j] = fmaf(av[i], bv[j], part[i][j]);
    }
    if ((it & 7) == 7 || mb + BK >= mend) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { acc[i][j] += part[i][j]; part[i][j] = 0.f; }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = c0 + ty * 4 + i;
    if (c >= p.C) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = k0 + tx * 4 + j;
      if (k < p.K) atomicAdd(dw + ((int64_t)tap * p.C + c) * p.K + k, acc[i][j]);
    }
  }
}

// ------------------------------------------------------------------------------------
// host side
static int validate(const gg_conv_desc* d) {
  GG_REQUIRE(d != nullptr, GG_ERR_INVALID, "conv: null descriptor");
  GG_REQUIRE(d->N > 0 && d->D > 0 && d->H > 0 && d->W > 0 && d->C > 0 && d->Do > 0 && d->Ho > 0 && d->Wo > 0 && d->K > 0,
             GG_ERR_INVALID, "conv: non-positive dimension");
  GG_REQUIRE(d->kd > 0 && d->kh > 0 && d->kw > 0 && d->kd * d->kh * d->kw <= MAX_TAPS, GG_ERR_UNSUPPORTED,
             "conv: filter has %d taps, kernels support <= %d", d->kd * d->kh * d->kw, MAX_TAPS);
  GG_REQUIRE(d->sd >= 1 && d->sh >= 1 && d->sw >= 1 && d->sd * d->sh * d->sw <= MAX_CLASSES, GG_ERR_UNSUPPORTED,
             "conv: stride product > %d unsupported", MAX_CLASSES);
  GG_REQUIRE((d->large_dtype == GG_F32 || d->large_dtype == GG_BF16) && (d->small_dtype == GG_F32 || d->small_dtype == GG_BF16),
             GG_ERR_INVALID, "conv: bad dtype");
  // consistency of the relation: every small index must map inside [-(k-1), large+k-1]
  GG_REQUIRE((d->Do - 1) * d->sd - d->pd < d->D && (d->Ho - 1) * d->sh - d->ph < d->H && (d->Wo - 1) * d->sw - d->pw < d->W,
             GG_ERR_INVALID, "conv: small grid does not fit the large grid");
  return GG_OK;
}

template <typename TA, typename TO, bool BKC>
static int launch_pixgemm(const void* A, const float* Wt, const float* bias, void* out, const PixGemmParams& p,
                          int64_t maxM, int red_len, cudaStream_t st) {
  dim3 grid((unsigned)ceil_div64(maxM, BM), (unsigned)ceil_div(p.Nc, BN), (unsigned)p.nclasses);
  const bool vecA = (p.Ac % 4 == 0) && ((uintptr_t)A % 16 == 0);
  const bool vecB = ((uintptr_t)Wt % 16 == 0) && (BKC ? (p.ldbn % 4 == 0 && red_len % 4 == 0) : (p.ldbk % 4 == 0 && p.Nc % 4 == 0));
  bool wofs_ok = true;
  for (int i = 0; i < MAX_TAPS; ++i) wofs_ok = wofs_ok && (p.taps[i].wofs % 4 == 0);
  const bool vecO = (p.Nc % 4 == 0) && ((uintptr_t)out % 16 == 0);
  const bool vB = vecB && wofs_ok;
#define GG_LAUNCH(VA, VB, VO)                                                                                   \
  Launch(grid, 256, 0, st)(pixgemm_kernel<TA, TO, BKC, VA, VB, VO>, (const TA*)A, Wt, bias, (TO*)out, p)
  if (vecA && vB && vecO) GG_LAUNCH(true, true, true);
  else if (vecA && vB) GG_LAUNCH(true, true, false);
  else if (vB && vecO) GG_LAUNCH(false, true, true);
  else if (vecA && vecO) GG_LAUNCH(true, false, true);
  else GG_LAUNCH(false, false, false);
#undef GG_LAUNCH
  return check_launch("pixgemm");
}

template <bool BKC>
static int dispatch_pixgemm(int a_dt, int o_dt, const void* A, const float* Wt, const float* bias, void* out,
                            const PixGemmParams& p, int64_t maxM, int red_len, cudaStream_t st) {
  if (a_dt == GG_F32 && o_dt == GG_F32) return launch_pixgemm<float, float, BKC>(A, Wt, bias, out, p, maxM, red_len, st);
  if (a_dt == GG_F32 && o_dt == GG_BF16) return launch_pixgemm<float, bf16, BKC>(A, Wt, bias, out, p, maxM, red_len, st);
  if (a_dt == GG_BF16 && o_dt == GG_F32) return launch_pixgemm<bf16, float, BKC>(A, Wt, bias, out, p, maxM, red_len, st);
  return launch_pixgemm<bf16, bf16, BKC>(A, Wt, bias, out, p, maxM, red_len, st);
}

int simt_conv_down(const gg_conv_desc* d, const void* large, const float* w, const float* bias, void* small, cudaStream_t st) {
  int rc = validate(d);
  if (rc) return rc;
  PixGemmParams p{};
  p.Mn = d->N;
  p.Ad = d->D; p.Ah = d->H; p.Aw = d->W; p.Ac = d->C;
  p.Sd = d->sd; p.Sh = d->sh; p.Sw = d->sw;
  p.Nc = d->K; p.ldbk = d->K; p.ldbn = 1;
  p.OD = d->Do; p.OH = d->Ho; p.OW = d->Wo; p.osd = p.osh = p.osw = 1;
  p.act = d->act; p.act_param = d->act_param;
  p.nclasses = 1;
  p.cls[0] = ClassInfo{d->Do, d->Ho, d->Wo, 0, 0, 0, 0, d->kd * d->kh * d->kw};
  int t = 0;
  for (int a = 0; a < d->kd; ++a)
    for (int b = 0; b < d->kh; ++b)
      for (int c = 0; c < d->kw; ++c, ++t)
        p.taps[t] = Tap{(int16_t)(a - d->pd), (int16_t)(b - d->ph), (int16_t)(c - d->pw), 0, t * d->C * d->K};
  const int64_t M = (int64_t)d->N * d->Do * d->Ho * d->Wo;
  return dispatch_pixgemm<false>(d->large_dtype, d->small_dtype, large, w, bias, small, p, M, d->C, st);
}

int simt_conv_up(const gg_conv_desc* d, const void* small, const float* w, const float* bias, void* large, cudaStream_t st) {
  int rc = validate(d);
  if (rc) return rc;
  PixGemmParams p{};
  p.Mn = d->N;
  p.Ad = d->Do; p.Ah = d->Ho; p.Aw = d->Wo; p.Ac = d->K;
  p.Sd = p.Sh = p.Sw = 1;
  p.Nc = d->C; p.ldbk = 1; p.ldbn = d->K;
  p.OD = d->D; p.OH = d->H; p.OW = d->W; p.osd = d->sd; p.osh = d->sh; p.osw = d->sw;
  p.act = d->act; p.act_param = d->act_param;
  int ncls = 0, nt = 0;
  int64_t maxM = 0;
  for (int ad = 0; ad < d->sd; ++ad)
    for (int ah = 0; ah < d->sh; ++ah)
      for (int aw = 0; aw < d->sw; ++aw) {
        ClassInfo ci{};
        ci.Md = (d->D - ad + d->sd - 1) / d->sd; ci.Mh = (d->H - ah + d->sh - 1) / d->sh; ci.Mw = (d->W - aw + d->sw - 1) / d->sw;
        ci.od0 = ad; ci.oh0 = ah; ci.ow0 = aw;
        ci.tap_begin = nt;
        for (int a = 0; a < d->kd; ++a) {
          if ((ad + d->pd - a) % d->sd != 0) continue;
          for (int b = 0; b < d->kh; ++b) {
            if ((ah + d->ph - b) % d->sh != 0) continue;
            for (int c = 0; c < d->kw; ++c) {
              if ((aw + d->pw - c) % d->sw != 0) continue;
              // floor division is exact here (remainder zero), also for negative numerators
              p.taps[nt++] = Tap{(int16_t)((ad + d->pd - a) / d->sd), (int16_t)((ah + d->ph - b) / d->sh),
                                 (int16_t)((aw + d->pw - c) / d->sw), 0, ((a * d->kh + b) * d->kw + c) * d->C * d->K};
            }
          }
        }
        ci.tap_end = nt;
        if (ci.Md <= 0 || ci.Mh <= 0 || ci.Mw <= 0) { ci.Md = ci.Mh = ci.Mw = 0; }
        maxM = std::max<int64_t>(maxM, (int64_t)d->N * ci.Md * ci.Mh * ci.Mw);
        p.cls[ncls++] = ci;
      }
  p.nclasses = ncls;
  return dispatch_pixgemm<true>(d->small_dtype, d->large_dtype, small, w, bias, large, p, maxM, d->K, st);
}

template <typename TL, typename TS>
static int launch_wgrad(const void* large, const void* small, float* dw, const WgradParams& p, cudaStream_t st) {
  dim3 grid((unsigned)ceil_div(p.C, BM), (unsigned)ceil_div(p.K, BN), (unsigned)(p.kd * p.kh * p.kw * p.splits));
  const bool vl = (p.C % 4 == 0) && ((uintptr_t)large % 16 == 0), vs = (p.K % 4 == 0) && ((uintptr_t)small % 16 == 0);
  if (vl && vs) Launch(grid, 256, 0, st)(wgrad_kernel<TL, TS, true, true>, (const TL*)large, (const TS*)small, dw, p);
  else if (vs) Launch(grid, 256, 0, st)(wgrad_kernel<TL, TS, false, true>, (const TL*)large, (const TS*)small, dw, p);
  else if (vl) Launch(grid, 256, 0, st)(wgrad_kernel<TL, TS, true, false>, (const TL*)large, (const TS*)small, dw, p);
  else Launch(grid, 256, 0, st)(wgrad_kernel<TL, TS, false, false>, (const TL*)large, (const TS*)small, dw, p);
  return check_launch("wgrad");
}

int simt_conv_wgrad(const gg_conv_desc* d, const void* large, const void* small, float* dw, cudaStream_t st) {
  int rc = validate(d);
  if (rc) return rc;
  WgradParams p{};
  p.N = d->N; p.D = d->D; p.H = d->H; p.W = d->W; p.C = d->C; p.Do = d->Do; p.Ho = d->Ho; p.Wo = d->Wo; p.K = d->K;
  p.kd = d->kd; p.kh = d->kh; p.kw = d->kw; p.sd = d->sd; p.sh = d->sh; p.sw = d->sw; p.pd = d->pd; p.ph = d->ph; p.pw = d->pw;
  p.Mtot = (int64_t)d->N * d->Do * d->Ho * d->Wo;
  const int64_t tiles = (int64_t)ceil_div(d->C, BM) * ceil_div(d->K, BN) * d->kd * d->kh * d->kw;
  int64_t splits = std::max<int64_t>(1, (148 * 4 + tiles - 1) / tiles);
  splits = std::max<int64_t>(splits, ceil_div64(p.Mtot, 1024));   // <= 1024 pixels per CTA: bounded error growth, more parallelism
  splits = std::min<int64_t>(splits, std::max<int64_t>(1, p.Mtot / 128));
  splits = std::min<int64_t>(splits, 65535 / (d->kd * d->kh * d->kw));
  p.chunk = ceil_div64(ceil_div64(p.Mtot, splits), BK) * BK;
  p.splits = (int)ceil_div64(p.Mtot, p.chunk);
  if (d->large_dtype == GG_F32 && d->small_dtype == GG_F32) return launch_wgrad<float, float>(large, small, dw, p, st);
  if (d->large_dtype == GG_F32) return launch_wgrad<float, bf16>(large, small, dw, p, st);
  if (d->small_dtype == GG_F32) return launch_wgrad<bf16, float>(large, small, dw, p, st);
  return launch_wgrad<bf16, bf16>(large, small, dw, p, st);
}

// ---- linear layers through the same kernels (one tap, 1x1x1 grid) -------------------
int simt_linear_fwd(const void* x, int x_dt, const float* matrix, const float* bias, void* y, int y_dt, int rows, int in_dim,
                    int out_dim, int act, float act_param, cudaStream_t st) {
  PixGemmParams p{};
  p.Mn = rows; p.Ad = p.Ah = p.Aw = 1; p.Ac = in_dim; p.Sd = p.Sh = p.Sw = 1;
  p.Nc = out_dim; p.ldbk = out_dim; p.ldbn = 1;
  p.OD = p.OH = p.OW = 1; p.osd = p.osh = p.osw = 1; p.act = act; p.act_param = act_param;
  p.nclasses = 1; p.cls[0] = ClassInfo{1, 1, 1, 0, 0, 0, 0, 1};
  p.taps[0] = Tap{0, 0, 0, 0, 0};
  return dispatch_pixgemm<false>(x_dt, y_dt, x, matrix, bias, y, p, rows, in_dim, st);
}

int simt_linear_dgrad(const void* dy, int dy_dt, const float* matrix, void* dx, int dx_dt, int rows, int in_dim, int out_dim,
                      cudaStream_t st) {
  PixGemmParams p{};
  p.Mn = rows; p.Ad = p.Ah = p.Aw = 1; p.Ac = out_dim; p.Sd = p.Sh = p.Sw = 1;
  p.Nc = in_dim; p.ldbk = 1; p.ldbn = out_dim;
  p.OD = p.OH = p.OW = 1; p.osd = p.osh = p.osw = 1; p.act = GG_ACT_NONE;
  p.nclasses = 1; p.cls[0] = ClassInfo{1, 1, 1, 0, 0, 0, 0, 1};
  p.taps[0] = Tap{0, 0, 0, 0, 0};
  return dispatch_pixgemm<true>(dy_dt, dx_dt, dy, matrix, nullptr, dx, p, rows, out_dim, st);
}

int simt_linear_wgrad(const void* x, int x_dt, const void* dy, int dy_dt, float* dmatrix, int rows, int in_dim, int out_dim,
                      cudaStream_t st) {
  gg_conv_desc d{};
  d.N = rows; d.D = d.H = d.W = 1; d.C = in_dim; d.Do = d.Ho = d.Wo = 1; d.K = out_dim;
  d.kd = d.kh = d.kw = 1; d.sd = d.sh = d.sw = 1; d.large_dtype = x_dt; d.small_dtype = dy_dt;
  return simt_conv_wgrad(&d, x, dy, dmatrix, st);
}

}  // namespace gg
