// Fused elementwise / small-reduction kernels: activation fwd/bwd, casts, filter packing,
// sigmoid cross-entropy fwd+bwd, first-frame MSE, TF-flavoured Adam, skinny linear layers
// (out_dim <= 4: d_h3_lin, dvideo_h4, d_final_fc) and the fused BasicLSTMCell step.
// All HBM- or latency-bound; 16 B vector accesses, grid-stride loops sized to the SM count.
#include <algorithm>

#include "common.cuh"

namespace gg {

constexpr int PW_THREADS = 256;
static inline int pw_blocks(int64_t n_vec) { return (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div64(n_vec, PW_THREADS), 148 * 8)); }
static inline bool al16(const void* p) { return ((uintptr_t)p % 16) == 0; }

// ---- activations ------------------------------------------------------------------------
template <typename TY, typename TD, typename TO, bool VEC>
__global__ void act_bwd_kernel(const TY* __restrict__ y, const TD* __restrict__ dy, TO* __restrict__ dx, int64_t n, int act, float ap) {
  pdl_grid_sync();
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  if (VEC) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n / 4; i += stride) {
      const float4 a = ld4(y + i * 4), g = ld4(dy + i * 4);
      float gv[4] = {g.x, g.y, g.z, g.w};
      const float yv[4] = {a.x, a.y, a.z, a.w};
      act_bwd_out_vec<4>(gv, yv, act, ap);
      st4(dx + i * 4, make_float4(gv[0], gv[1], gv[2], gv[3]));
    }
  } else {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
      stf(dx + i, ldf(dy + i) * act_grad_from_out(ldf(y + i), act, ap));
  }
}

template <typename TX, typename TY, bool VEC>
__global__ void act_fwd_kernel(const TX* __restrict__ x, TY* __restrict__ y, int64_t n, int act, float ap) {
  pdl_grid_sync();
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  if (VEC) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n / 4; i += stride) {
      const float4 a = ld4(x + i * 4);
      float v[4] = {a.x, a.y, a.z, a.w};
      act_fwd_vec<4>(v, act, ap);
      st4(y + i * 4, make_float4(v[0], v[1], v[2], v[3]));
    }
  } else {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) stf(y + i, act_fwd(ldf(x + i), act, ap));
  }
}

__global__ void axpby_kernel(const float* __restrict__ x, float a, float* __restrict__ y, float b, int64_t n) {
  pdl_grid_sync();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    y[i] = a * x[i] + (b == 0.f ? 0.f : b * y[i]);
}

// ---- losses --------------------------------------------------------------------------------
__global__ void sigmoid_ce_kernel(const float* __restrict__ logits, int64_t n, float target, float weight, float* __restrict__ loss_out,
                                  int accumulate, float* __restrict__ dlogits) {
  pdl_grid_sync();
  __shared__ float red[PW_THREADS];
  float acc = 0.f;
  const float inv = 1.f / (float)n;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    const float x = logits[i];
    // tf.nn.sigmoid_cross_entropy_with_logits: max(x,0) - x*z + log(1+exp(-|x|))
    acc += fmaxf(x, 0.f) - x * target + log1pf(expf(-fabsf(x)));
    if (dlogits) dlogits[i] = weight * inv * (1.f / (1.f + expf(-x)) - target);
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int s = PW_THREADS / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) loss_out[0] = (accumulate ? loss_out[0] : 0.f) + weight * red[0] * inv;
}

__global__ void mse_kernel(const float* __restrict__ a, int64_t as, const float* __restrict__ b, int64_t bs, int64_t rows, int64_t cols,
                           float scalar, float* __restrict__ loss_out, int accumulate, float* __restrict__ da) {
  pdl_grid_sync();
  __shared__ float red[PW_THREADS];
  float acc = 0.f;
  const int64_t n = rows * cols;
  const float inv = 1.f / (float)n;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    const int64_t r = i / cols, c = i - r * cols;
    const float d = a[r * as + c] - b[r * bs + c];
    acc += d * d;
    if (da) da[r * cols + c] = scalar * 2.f * d * inv;
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int s = PW_THREADS / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) loss_out[0] = (accumulate ? loss_out[0] : 0.f) + scalar * red[0] * inv;
}

// ---- L2 + L1 distance to a constant target (latent search: z_space_finder.py:258-292) -------------------------
// loss (+)= w2 * mean((a-t)^2) + w1 * mean|a-t|;   da = (2 w2 (a-t) + w1 sign(a-t)) / n   (tf.abs' = sign, 0 at 0).
// Blocks write their partial sums to `partial`; the block that takes the last ticket adds them in block order, so the
// loss does not depend on the order blocks finish in.  `ticket` must be zero on entry and is zero again on exit.
constexpr int DIST_MAX_BLOCKS = 296;
template <typename TA>
__global__ void __launch_bounds__(PW_THREADS)
distance_loss_kernel(const TA* __restrict__ a, const float* __restrict__ t, int64_t n, float w2, float w1, float* __restrict__ loss_out,
                     int accumulate, TA* __restrict__ da, float* __restrict__ partial, unsigned* __restrict__ ticket) {
  pdl_grid_sync();
  __shared__ float red2[PW_THREADS / 32], red1[PW_THREADS / 32];
  const float inv = 1.f / (float)n;
  float s2 = 0.f, s1 = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float d = ldf(a + i) - __ldg(t + i);
    s2 += d * d;
    s1 += fabsf(d);
    if (da) stf(da + i, (2.f * w2 * d + w1 * (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f))) * inv);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
  }
  if ((threadIdx.x & 31) == 0) { red2[threadIdx.x >> 5] = s2; red1[threadIdx.x >> 5] = s1; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float b2 = 0.f, b1 = 0.f;
    for (int w = 0; w < PW_THREADS / 32; ++w) { b2 += red2[w]; b1 += red1[w]; }
    partial[2 * blockIdx.x] = b2;
    partial[2 * blockIdx.x + 1] = b1;
    __threadfence();
    if (atomicAdd(ticket, 1u) == gridDim.x - 1) {
      __threadfence();
      float t2 = 0.f, t1 = 0.f;
      for (unsigned b = 0; b < gridDim.x; ++b) { t2 += __ldcg(partial + 2 * b); t1 += __ldcg(partial + 2 * b + 1); }
      loss_out[0] = (accumulate ? loss_out[0] : 0.f) + (w2 * t2 + w1 * t1) * inv;
      *ticket = 0u;
    }
  }
}

// ---- TF Adam (model.py:153-156) ----------------------------------------------
__global__ void __launch_bounds__(PW_THREADS)
adam_kernel(float* __restrict__ p, bf16* __restrict__ pb, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, int64_t n,
            float lr_t, float b1, float b2, float eps, float gs) {
  pdl_grid_sync();
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t n4 = n / 4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 pp = reinterpret_cast<float4*>(p)[i], gg4 = __ldg(reinterpret_cast<const float4*>(g) + i);
    float4 mm = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
    float* P = &pp.x; float* G = &gg4.x; float* M = &mm.x; float* V = &vv.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float gr = G[k] * gs;
      M[k] = b1 * M[k] + (1.f - b1) * gr;
      V[k] = b2 * V[k] + (1.f - b2) * gr * gr;
      P[k] -= lr_t * M[k] / (sqrtf(V[k]) + eps);
    }
    reinterpret_cast<float4*>(p)[i] = pp;
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
    if (pb != nullptr) st4(pb + 4 * i, pp);          // bf16 shadow: the tensor-core kernels' filter operand, refreshed in the same pass
  }
  for (int64_t i = n4 * 4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float gr = g[i] * gs;
    const float mi = b1 * m[i] + (1.f - b1) * gr, vi = b2 * v[i] + (1.f - b2) * gr * gr;
    m[i] = mi; v[i] = vi;
    const float pi = p[i] - lr_t * mi / (sqrtf(vi) + eps);
    p[i] = pi;
    if (pb != nullptr) pb[i] = __float2bfloat16_rn(pi);
  }
}

// device-side step counter so that a captured CUDA graph advances Adam's bias correction:
// state[0] = t (int32), state[1] = lr_t (float bits).  Double precision like the host formula.
__global__ void adam_tick_kernel(int* __restrict__ state, float lr, float b1, float b2) {
  pdl_grid_sync();
  const int t = state[0] + 1;
  state[0] = t;
  const double lr_t = (double)lr * sqrt(1.0 - pow((double)b2, (double)t)) / (1.0 - pow((double)b1, (double)t));
  reinterpret_cast<float*>(state)[1] = (float)lr_t;
}
__global__ void __launch_bounds__(PW_THREADS)
adam_dev_kernel(float* __restrict__ p, bf16* __restrict__ pb, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                int64_t n, const int* __restrict__ state, float b1, float b2, float eps, float gs) {
  pdl_grid_sync();
  const float lr_t = reinterpret_cast<const float*>(state)[1];
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t n4 = n / 4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 pp = reinterpret_cast<float4*>(p)[i], gg4 = __ldg(reinterpret_cast<const float4*>(g) + i);
    float4 mm = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
    float* P = &pp.x; float* G = &gg4.x; float* M = &mm.x; float* V = &vv.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float gr = G[k] * gs;
      M[k] = b1 * M[k] + (1.f - b1) * gr;
      V[k] = b2 * V[k] + (1.f - b2) * gr * gr;
      P[k] -= lr_t * M[k] / (sqrtf(V[k]) + eps);
    }
    reinterpret_cast<float4*>(p)[i] = pp;
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
    if (pb != nullptr) st4(pb + 4 * i, pp);          // bf16 shadow: the tensor-core kernels' filter operand, refreshed in the same pass
  }
  for (int64_t i = n4 * 4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float gr = g[i] * gs;
    const float mi = b1 * m[i] + (1.f - b1) * gr, vi = b2 * v[i] + (1.f - b2) * gr * gr;
    m[i] = mi; v[i] = vi;
    const float pi = p[i] - lr_t * mi / (sqrtf(vi) + eps);
    p[i] = pi;
    if (pb != nullptr) pb[i] = __float2bfloat16_rn(pi);
  }
}

// ---- skinny linear (out_dim <= 4) ---------------------------------------------------------------
// fwd: y[r, n] = act(sum_k x[r,k] W[k,n] + b[n]); one CTA per row
template <typename TX, typename TY>
__global__ void __launch_bounds__(PW_THREADS)
skinny_fwd_kernel(const TX* __restrict__ x, const float* __restrict__ W, const float* __restrict__ bias, TY* __restrict__ y, int in_dim,
                  int out_dim, int act, float ap) {
  pdl_grid_sync();
  const int r = blockIdx.x;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  if (out_dim == 1 && (in_dim & 3) == 0 && ((uintptr_t)x % 16 == 0) && ((uintptr_t)W % 16 == 0)) {
    // the GEMV of d_h3_lin / dvideo_h4: 4 elements per thread per trip, all loads independent
    const TX* xr = x + (int64_t)r * in_dim;
#pragma unroll 8                                                // (8 trips for d_h3_lin's 8192 inputs: their 16 loads go out together)
    for (int k = threadIdx.x * 4; k < in_dim; k += blockDim.x * 4) {
      const float4 xv = ld4(xr + k), wv = ld4(W + k);
      acc[0] = fmaf(xv.x, wv.x, fmaf(xv.y, wv.y, fmaf(xv.z, wv.z, fmaf(xv.w, wv.w, acc[0]))));
    }
  } else {
    for (int k = threadIdx.x; k < in_dim; k += blockDim.x) {
      const float xv = ldf(x + (int64_t)r * in_dim + k);
      for (int n = 0; n < out_dim; ++n) acc[n] = fmaf(xv, __ldg(W + (int64_t)k * out_dim + n), acc[n]);
    }
  }
  __shared__ float red[4][PW_THREADS / 32];
  for (int n = 0; n < out_dim; ++n) {
    const float s = warp_sum(acc[n]);
    if ((threadIdx.x & 31) == 0) red[n][threadIdx.x >> 5] = s;
  }
  __syncthreads();
  if (threadIdx.x < out_dim) {
    float s = bias ? bias[threadIdx.x] : 0.f;
    for (int w = 0; w < PW_THREADS / 32; ++w) s += red[threadIdx.x][w];
    stf(y + (int64_t)r * out_dim + threadIdx.x, act_fwd(s, act, ap));
  }
}
// dgrad: dx[r,k] = sum_n dy[r,n] W[k,n]
template <typename TD, typename TO>
__global__ void skinny_dgrad_kernel(const TD* __restrict__ dy, const float* __restrict__ W, TO* __restrict__ dx, int64_t rows, int in_dim, int out_dim) {
  pdl_grid_sync();
  const int64_t total = rows * in_dim;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / in_dim;
    const int k = (int)(i - r * in_dim);
    float s = 0.f;
    for (int n = 0; n < out_dim; ++n) s = fmaf(ldf(dy + r * out_dim + n), __ldg(W + (int64_t)k * out_dim + n), s);
    stf(dx + i, s);
  }
}
// wgrad: dW[k,n] += sum_r x[r,k] dy[r,n]
template <typename TX, typename TD>
__global__ void skinny_wgrad_kernel(const TX* __restrict__ x, const TD* __restrict__ dy, float* __restrict__ dW, int rows, int in_dim, int out_dim) {
  pdl_grid_sync();
  // grid = (k blocks, row chunks): each CTA reduces a chunk of rows for 128 k's, then one atomic per element
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= in_dim) return;
  const int rchunk = (rows + gridDim.y - 1) / gridDim.y;
  const int r0 = blockIdx.y * rchunk, r1 = min(rows, r0 + rchunk);
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
  for (int r = r0; r < r1; ++r) {
    const float xv = ldf(x + (int64_t)r * in_dim + k);
    for (int n = 0; n < out_dim; ++n) acc[n] = fmaf(xv, ldf(dy + (int64_t)r * out_dim + n), acc[n]);
  }
  for (int n = 0; n < out_dim; ++n) atomicAdd(dW + (int64_t)k * out_dim + n, acc[n]);
}

// ---- BasicLSTMCell step (recurrent_DCGAN.py:199-200) --------------------------------
__device__ __forceinline__ float sigm(float x) { return 1.f / (1.f + expf(-x)); }

// one CTA per batch row; thread j < H produces the 4 gates of unit j
__global__ void lstm_fwd_kernel(const float* __restrict__ gx, const float* __restrict__ Wh, const float* __restrict__ c_prev,
                                const float* __restrict__ h_prev, float* __restrict__ c_out, float* __restrict__ h_out,
                                float* __restrict__ gates_out, int H, float fb) {
  pdl_grid_sync();
  extern __shared__ float hs[];
  const int b = blockIdx.x;
  for (int k = threadIdx.x; k < H; k += blockDim.x) hs[k] = h_prev[(int64_t)b * H + k];
  __syncthreads();
  for (int j = threadIdx.x; j < H; j += blockDim.x) {
    float g[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) g[q] = gx[(int64_t)b * 4 * H + q * H + j];
    for (int k = 0; k < H; ++k) {
      const float hv = hs[k];
      const float* wr = Wh + (int64_t)k * 4 * H + j;
#pragma unroll
      for (int q = 0; q < 4; ++q) g[q] = fmaf(hv, __ldg(wr + q * H), g[q]);
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) gates_out[(int64_t)b * 4 * H + q * H + j] = g[q];
    const float c = c_prev[(int64_t)b * H + j] * sigm(g[2] + fb) + sigm(g[0]) * tanhf(g[1]);
    c_out[(int64_t)b * H + j] = c;
    h_out[(int64_t)b * H + j] = tanhf(c) * sigm(g[3]);
  }
}

__global__ void lstm_bwd_kernel(const float* __restrict__ gates, const float* __restrict__ c_prev, const float* __restrict__ c_out,
                                const float* __restrict__ dh, const float* __restrict__ dc, const float* __restrict__ Wh,
                                float* __restrict__ dgates, float* __restrict__ dc_prev, float* __restrict__ dh_prev, int H, float fb) {
  pdl_grid_sync();
  extern __shared__ float dg[];  // [4H]
  const int b = blockIdx.x;
  for (int j = threadIdx.x; j < H; j += blockDim.x) {
    const float gi = gates[(int64_t)b * 4 * H + j], gj = gates[(int64_t)b * 4 * H + H + j];
    const float gf = gates[(int64_t)b * 4 * H + 2 * H + j], go = gates[(int64_t)b * 4 * H + 3 * H + j];
    const float si = sigm(gi), tj = tanhf(gj), sf = sigm(gf + fb), so = sigm(go);
    const float tc = tanhf(c_out[(int64_t)b * H + j]);
    const float dhv = dh[(int64_t)b * H + j];
    const float dct = (dc ? dc[(int64_t)b * H + j] : 0.f) + dhv * so * (1.f - tc * tc);
    const float d_o = dhv * tc * so * (1.f - so);
    const float d_i = dct * tj * si * (1.f - si);
    const float d_j = dct * si * (1.f - tj * tj);
    const float d_f = dct * c_prev[(int64_t)b * H + j] * sf * (1.f - sf);
    dc_prev[(int64_t)b * H + j] = dct * sf;
    dg[j] = d_i; dg[H + j] = d_j; dg[2 * H + j] = d_f; dg[3 * H + j] = d_o;
    dgates[(int64_t)b * 4 * H + j] = d_i; dgates[(int64_t)b * 4 * H + H + j] = d_j;
    dgates[(int64_t)b * 4 * H + 2 * H + j] = d_f; dgates[(int64_t)b * 4 * H + 3 * H + j] = d_o;
  }
  __syncthreads();
  for (int k = threadIdx.x; k < H; k += blockDim.x) {
    const float* wr = Wh + (int64_t)k * 4 * H;
    float s = 0.f;
    for (int q = 0; q < 4 * H; ++q) s = fmaf(dg[q], __ldg(wr + q), s);
    dh_prev[(int64_t)b * H + k] = s;
  }
}

}  // namespace gg

using namespace gg;

namespace gg {
// dx = dy * act'(y) and, from the same pass, the bias gradient of the layer that produced y: dbias[c] += sum_rows dx[row, c] for a
// channel-last fp32 tensor with C <= 4 channels (g_h4: tanh' of the generated image + the 3-channel bias gradient, model.py:321-324).
// A thread owns 4 consecutive elements per trip (channel of element e = e % C); block partials -> one atomic per channel per block.
__global__ void __launch_bounds__(PW_THREADS)
act_bwd_bias_kernel(const float* __restrict__ y, const float* __restrict__ dy, float* __restrict__ dx, int64_t n, int C, int act, float ap,
                    float* __restrict__ dbias) {
  pdl_grid_sync();
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n / 4; i += stride) {
    const float4 a = ld4(y + i * 4), g = ld4(dy + i * 4);
    float gv[4] = {g.x, g.y, g.z, g.w};
    const float yv[4] = {a.x, a.y, a.z, a.w};
    act_bwd_out_vec<4>(gv, yv, act, ap);
    st4(dx + i * 4, make_float4(gv[0], gv[1], gv[2], gv[3]));
    int c = (int)((i * 4) % C);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[q] += (q == c) ? gv[k] : 0.f;
      c = (c + 1 == C) ? 0 : c + 1;
    }
  }
  __shared__ float red[4][PW_THREADS / 32];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float v = warp_sum(acc[q]);
    if ((threadIdx.x & 31) == 0) red[q][threadIdx.x >> 5] = v;
  }
  __syncthreads();
  if ((int)threadIdx.x < C) {
    float v = 0.f;
    for (int w = 0; w < PW_THREADS / 32; ++w) v += red[threadIdx.x][w];
    atomicAdd(dbias + threadIdx.x, v);
  }
}
}  // namespace gg

extern "C" int gg_act_bwd_bias(const void* y, int32_t y_dt, const void* dy, int32_t dy_dt, void* dx, int32_t dx_dt, int64_t rows, int32_t C,
                               int32_t act, float ap, float* dbias, void* stream) {
  GG_REQUIRE(y && dy && dx && dbias && rows > 0 && C > 0, GG_ERR_INVALID, "act_bwd_bias: bad argument");
  const int64_t n = rows * C;
  if (y_dt == GG_F32 && dy_dt == GG_F32 && dx_dt == GG_F32 && C <= 4 && n % 4 == 0 && al16(y) && al16(dy) && al16(dx)) {
    Launch(pw_blocks(n / 4), PW_THREADS, 0, (cudaStream_t)stream)(act_bwd_bias_kernel, (const float*)y, (const float*)dy, (float*)dx, n, (int)C,
                                                                  (int)act, ap, dbias);
    return check_launch("act_bwd_bias");
  }
  int rc = gg_act_bwd(y, y_dt, dy, dy_dt, dx, dx_dt, n, act, ap, stream);       // any other case: the two separate launches
  if (rc) return rc;
  return gg_bias_grad(dx, dx_dt, dbias, rows, C, stream);
}

extern "C" int gg_act_bwd(const void* y, int32_t y_dt, const void* dy, int32_t dy_dt, void* dx, int32_t dx_dt, int64_t n, int32_t act,
                          float ap, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  GG_REQUIRE(y && dy && dx && n > 0, GG_ERR_INVALID, "act_bwd: bad argument");
  const bool vec = (n % 4 == 0) && al16(y) && al16(dy) && al16(dx);
  const int blocks = pw_blocks(vec ? n / 4 : n);
#define GG_AB(TY, TD, TO)                                                                                            \
  do {                                                                                                               \
    if (vec) Launch(blocks, PW_THREADS, 0, st)(act_bwd_kernel<TY, TD, TO, true>, (const TY*)y, (const TD*)dy, (TO*)dx, n, act, ap);  \
    else Launch(blocks, PW_THREADS, 0, st)(act_bwd_kernel<TY, TD, TO, false>, (const TY*)y, (const TD*)dy, (TO*)dx, n, act, ap);     \
  } while (0)
  const int key = (y_dt == GG_BF16 ? 4 : 0) | (dy_dt == GG_BF16 ? 2 : 0) | (dx_dt == GG_BF16 ? 1 : 0);
  switch (key) {
    case 0: GG_AB(float, float, float); break;
    case 1: GG_AB(float, float, bf16); break;
    case 2: GG_AB(float, bf16, float); break;
    case 3: GG_AB(float, bf16, bf16); break;
    case 4: GG_AB(bf16, float, float); break;
    case 5: GG_AB(bf16, float, bf16); break;
    case 6: GG_AB(bf16, bf16, float); break;
    default: GG_AB(bf16, bf16, bf16); break;
  }
  return check_launch("act_bwd");
}

static int act_fwd_impl(const void* x, int x_dt, void* y, int y_dt, int64_t n, int act, float ap, cudaStream_t st) {
  const bool vec = (n % 4 == 0) && al16(x) && al16(y);
  const int blocks = pw_blocks(vec ? n / 4 : n);
#define GG_AF(TX, TY)                                                                                      \
  do {                                                                                                     \
    if (vec) Launch(blocks, PW_THREADS, 0, st)(act_fwd_kernel<TX, TY, true>, (const TX*)x, (TY*)y, n, act, ap);  \
    else Launch(blocks, PW_THREADS, 0, st)(act_fwd_kernel<TX, TY, false>, (const TX*)x, (TY*)y, n, act, ap);     \
  } while (0)
  if (x_dt == GG_F32 && y_dt == GG_F32) GG_AF(float, float);
  else if (x_dt == GG_F32) GG_AF(float, bf16);
  else if (y_dt == GG_F32) GG_AF(bf16, float);
  else GG_AF(bf16, bf16);
  return check_launch("act_fwd");
}

extern "C" int gg_act_fwd(const void* x, int32_t x_dt, void* y, int32_t y_dt, int64_t n, int32_t act, float ap, void* stream) {
  GG_REQUIRE(x && y && n > 0, GG_ERR_INVALID, "act_fwd: bad argument");
  return act_fwd_impl(x, x_dt, y, y_dt, n, act, ap, (cudaStream_t)stream);
}

extern "C" int gg_cast(const void* src, int32_t s_dt, void* dst, int32_t d_dt, int64_t n, void* stream) {
  GG_REQUIRE(src && dst && n > 0, GG_ERR_INVALID, "cast: bad argument");
  return act_fwd_impl(src, s_dt, dst, d_dt, n, GG_ACT_NONE, 0.f, (cudaStream_t)stream);
}

extern "C" int gg_axpby(const float* x, float a, float* y, float b, int64_t n, void* stream) {
  GG_REQUIRE(x && y && n > 0, GG_ERR_INVALID, "axpby: bad argument");
  Launch(pw_blocks(n), PW_THREADS, 0, (cudaStream_t)stream)(axpby_kernel, x, a, y, b, n);
  return check_launch("axpby");
}

namespace gg {
struct ScalarSrcs { const float* p[8]; };
__global__ void gather_scalars_kernel(ScalarSrcs s, float* __restrict__ dst, int n) {
  pdl_grid_sync();
  const int i = threadIdx.x;
  if (i < n && s.p[i] != nullptr) dst[i] = *s.p[i];
}
}  // namespace gg

extern "C" int gg_gather_scalars(const float* const* srcs, int32_t n, float* dst, void* stream) {
  GG_REQUIRE(srcs && dst && n > 0 && n <= 8, GG_ERR_INVALID, "gather_scalars: bad argument (1..8 device scalars)");
  ScalarSrcs s;
  for (int i = 0; i < 8; ++i) s.p[i] = i < n ? srcs[i] : nullptr;
  Launch(1, 32, 0, (cudaStream_t)stream)(gather_scalars_kernel, s, dst, (int)n);
  return check_launch("gather_scalars");
}

extern "C" int gg_sigmoid_ce(const float* logits, int64_t n, float target, float weight, float* loss_out, int32_t accumulate,
                             float* dlogits, void* stream) {
  GG_REQUIRE(logits && loss_out && n > 0, GG_ERR_INVALID, "sigmoid_ce: bad argument");
  Launch(1, PW_THREADS, 0, (cudaStream_t)stream)(sigmoid_ce_kernel, logits, n, target, weight, loss_out, accumulate, dlogits);
  return check_launch("sigmoid_ce");
}

extern "C" int gg_mse(const float* a, int64_t as, const float* b, int64_t bs, int64_t rows, int64_t cols, float scalar, float* loss_out,
                      int32_t accumulate, float* da, void* stream) {
  GG_REQUIRE(a && b && loss_out && rows > 0 && cols > 0, GG_ERR_INVALID, "mse: bad argument");
  Launch(1, PW_THREADS, 0, (cudaStream_t)stream)(mse_kernel, a, as, b, bs, rows, cols, scalar, loss_out, accumulate, da);
  return check_launch("mse");
}

extern "C" size_t gg_distance_loss_workspace_bytes(void) { return (size_t)(2 * DIST_MAX_BLOCKS) * sizeof(float) + 16; }

extern "C" int gg_distance_loss(const void* a, int32_t a_dt, const float* target, int64_t n, float w_l2, float w_l1, float* loss_out,
                                int32_t accumulate, void* da, void* ws, size_t ws_bytes, void* stream) {
  GG_REQUIRE(a && target && loss_out && ws && n > 0, GG_ERR_INVALID, "distance_loss: bad argument");
  GG_REQUIRE(ws_bytes >= gg_distance_loss_workspace_bytes() && al16(ws), GG_ERR_WORKSPACE, "distance_loss: workspace too small or misaligned");
  GG_REQUIRE(a_dt == GG_F32 || a_dt == GG_BF16, GG_ERR_UNSUPPORTED, "distance_loss: dtype");
  unsigned* ticket = reinterpret_cast<unsigned*>(ws);                    // zeroed once by the caller; the kernel leaves it zero
  float* partial = reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + 16);
  const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div64(n, PW_THREADS * 4), DIST_MAX_BLOCKS));
  if (a_dt == GG_BF16)
    Launch(blocks, PW_THREADS, 0, (cudaStream_t)stream)(distance_loss_kernel<bf16>, (const bf16*)a, target, n, w_l2, w_l1, loss_out,
                                                        (int)accumulate, (bf16*)da, partial, ticket);
  else
    Launch(blocks, PW_THREADS, 0, (cudaStream_t)stream)(distance_loss_kernel<float>, (const float*)a, target, n, w_l2, w_l1, loss_out,
                                                        (int)accumulate, (float*)da, partial, ticket);
  return check_launch("distance_loss");
}

extern "C" int gg_adam(float* p, void* p_bf16, const float* g, float* m, float* v, int64_t n, float lr_t, float b1, float b2, float eps, float gs,
                       void* stream) {
  GG_REQUIRE(p && g && m && v && n > 0, GG_ERR_INVALID, "adam: bad argument");
  GG_REQUIRE(al16(p) && al16(g) && al16(m) && al16(v) && ((uintptr_t)p_bf16 % 8) == 0, GG_ERR_INVALID, "adam: buffers must be 16-byte aligned (bf16 shadow: 8)");
  Launch(pw_blocks(n / 4 + 1), PW_THREADS, 0, (cudaStream_t)stream)(adam_kernel, p, (bf16*)p_bf16, g, m, v, n, lr_t, b1, b2, eps, gs);
  return check_launch("adam");
}

extern "C" int gg_adam_graph(float* p, void* p_bf16, const float* g, float* m, float* v, int64_t n, int32_t* state, float lr, float b1, float b2,
                             float eps, float gs, void* stream) {
  GG_REQUIRE(p && g && m && v && state && n > 0, GG_ERR_INVALID, "adam_graph: bad argument");
  GG_REQUIRE(al16(p) && al16(g) && al16(m) && al16(v) && ((uintptr_t)p_bf16 % 8) == 0, GG_ERR_INVALID, "adam_graph: buffers must be 16-byte aligned (bf16 shadow: 8)");
  Launch(1, 1, 0, (cudaStream_t)stream)(adam_tick_kernel, state, lr, b1, b2);
  int rc = check_launch("adam_tick");
  if (rc) return rc;
  Launch(pw_blocks(n / 4 + 1), PW_THREADS, 0, (cudaStream_t)stream)(adam_dev_kernel, p, (bf16*)p_bf16, g, m, v, n, state, b1, b2, eps, gs);
  return check_launch("adam_dev");
}

// The two halves of gg_adam_graph as separate calls: the tick (one thread: t += 1, lr_t in double precision) depends only on the
// previous step of the same group, so a caller can issue it EARLY on another stream -- under the update's forward pass -- and
// keep the 2.5 us launch (+ gap) off the critical path in front of the Adam launch.
extern "C" int gg_adam_tick(int32_t* state, float lr, float b1, float b2, void* stream) {
  GG_REQUIRE(state, GG_ERR_INVALID, "adam_tick: bad argument");
  Launch(1, 1, 0, (cudaStream_t)stream)(adam_tick_kernel, state, lr, b1, b2);
  return check_launch("adam_tick");
}
extern "C" int gg_adam_apply(float* p, void* p_bf16, const float* g, float* m, float* v, int64_t n, const int32_t* state, float b1, float b2,
                             float eps, float gs, void* stream) {
  GG_REQUIRE(p && g && m && v && state && n > 0, GG_ERR_INVALID, "adam_apply: bad argument");
  GG_REQUIRE(al16(p) && al16(g) && al16(m) && al16(v) && ((uintptr_t)p_bf16 % 8) == 0, GG_ERR_INVALID, "adam_apply: buffers must be 16-byte aligned (bf16 shadow: 8)");
  Launch(pw_blocks(n / 4 + 1), PW_THREADS, 0, (cudaStream_t)stream)(adam_dev_kernel, p, (bf16*)p_bf16, g, m, v, n, state, b1, b2, eps, gs);
  return check_launch("adam_dev");
}

namespace gg {
// ---- thin-input linears: in_dim <= 128, wide output (g_h0_lin: z[B,100] -> [B,8192], model.py:304) -------------------
// The generic SIMT GEMM spends 12-14 us on this 105-MFLOP problem.  Here a thread owns ONE output column: the matrix
// row-slices it needs are coalesced across the block, the (tiny) x tile sits in shared memory transposed so that four
// rows come back per broadcast 128-bit load, and 32 rows are accumulated in registers.
constexpr int THIN_ROWS = 32, THIN_MAXK = 128, THIN_COLS = 64;   // block = 64 output columns x 4 reduction groups

// y[r][j] = act(sum_i x[r][i] W[i][j] + b[j]): thread (c, g) accumulates i = g, g+4, ... for column c and 32 rows; the
// four partial sums meet in shared memory.  8 matrix loads per thread are in flight before the first FMA.
// stats != nullptr (fp32 pre-norm output feeding a train-mode batch norm over channel = column % Cc): the block also adds
// (sum, sum of squares) of its 32 rows x 64 columns to the fp64 accumulators [groups][2][Cc] -- the statistics pass
// (memset + colsum, 8-13 us for g_h0_lin's [64, 8192]) disappears; rows of one block lie in one row group (host checks).
template <typename TX, typename TY>
__global__ void __launch_bounds__(256)
thin_fwd_kernel(const TX* __restrict__ x, const float* __restrict__ W, const float* __restrict__ bias, TY* __restrict__ y, int rows,
                int in_dim, int out_dim, int act, float ap, double* __restrict__ stats, int Cc, int groups) {
  pdl_grid_sync();
  __shared__ __align__(16) float xs[THIN_MAXK * THIN_ROWS];                 // [i][r]            16 KB
  __shared__ __align__(16) float red[4 * THIN_ROWS * THIN_COLS];            // [g][r][c]         32 KB
  const int r0 = blockIdx.y * THIN_ROWS;
  {
    const int r = threadIdx.x >> 3, l = threadIdx.x & 7;        // 8 threads per row, 32 rows
    float xv[THIN_MAXK / 8];                                    // all of the thread's loads first (<= 16), then the transposed stores
#pragma unroll
    for (int u = 0; u < THIN_MAXK / 8; ++u) {
      const int i = l + 8 * u;
      xv[u] = (i < in_dim && r0 + r < rows) ? ldf(x + (int64_t)(r0 + r) * in_dim + i) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < THIN_MAXK / 8; ++u) {
      const int i = l + 8 * u;
      if (i < in_dim) xs[i * THIN_ROWS + r] = xv[u];
    }
  }
  __syncthreads();
  const int c = threadIdx.x & (THIN_COLS - 1), g = threadIdx.x >> 6;
  const int j = blockIdx.x * THIN_COLS + c;
  const bool col_ok = j < out_dim;
  float acc[THIN_ROWS];
#pragma unroll
  for (int r = 0; r < THIN_ROWS; ++r) acc[r] = 0.f;
  {
    constexpr int NB = THIN_MAXK / 4;                            // all of this thread's matrix elements in ONE load batch
    float m[NB];
#pragma unroll
    for (int u = 0; u < NB; ++u) { const int i = g + 4 * u; m[u] = (col_ok && i < in_dim) ? __ldg(W + (int64_t)i * out_dim + j) : 0.f; }
#pragma unroll
    for (int u = 0; u < NB; ++u) {
      const int i = g + 4 * u;
      if (i < in_dim) {
        const float4* xr = reinterpret_cast<const float4*>(xs + i * THIN_ROWS);
#pragma unroll
        for (int q = 0; q < THIN_ROWS / 4; ++q) {
          const float4 v = xr[q];
          acc[4 * q] = fmaf(v.x, m[u], acc[4 * q]); acc[4 * q + 1] = fmaf(v.y, m[u], acc[4 * q + 1]);
          acc[4 * q + 2] = fmaf(v.z, m[u], acc[4 * q + 2]); acc[4 * q + 3] = fmaf(v.w, m[u], acc[4 * q + 3]);
        }
      }
    }
  }
#pragma unroll
  for (int r = 0; r < THIN_ROWS; ++r) red[(g * THIN_ROWS + r) * THIN_COLS + c] = acc[r];
  __syncthreads();
  const float b = (bias && col_ok) ? __ldg(bias + j) : 0.f;
  float o[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {                                  // this thread finishes rows g*8 .. g*8+7
    const int r = g * 8 + k;
    o[k] = b + red[(0 * THIN_ROWS + r) * THIN_COLS + c] + red[(1 * THIN_ROWS + r) * THIN_COLS + c] +
           red[(2 * THIN_ROWS + r) * THIN_COLS + c] + red[(3 * THIN_ROWS + r) * THIN_COLS + c];
  }
  act_fwd_vec<8>(o, act, ap);
  double s1 = 0.0, s2 = 0.0;                                     // fp64 from the first add on: 16 DFMAs per thread, and the
#pragma unroll                                                   // statistics do not depend on how rows are split over threads
  for (int k = 0; k < 8; ++k) {
    const int r = r0 + g * 8 + k;
    if (r < rows && col_ok) { stf(y + (int64_t)r * out_dim + j, o[k]); s1 += (double)o[k]; s2 += (double)o[k] * (double)o[k]; }
  }
  if (stats != nullptr) {                                        // uniform over the block
    double* sred = reinterpret_cast<double*>(xs);                // the x tile is dead: [2][4][64] doubles = 4 KB
    sred[g * THIN_COLS + c] = s1;
    sred[(4 + g) * THIN_COLS + c] = s2;
    __syncthreads();
    if (threadIdx.x < 2 * THIN_COLS) {
      const int which = threadIdx.x >> 6, cc = threadIdx.x & (THIN_COLS - 1), jj = blockIdx.x * THIN_COLS + cc;
      if (jj < out_dim) {
        const double* q = sred + which * 4 * THIN_COLS + cc;
        const double t = (q[0] + q[THIN_COLS]) + (q[2 * THIN_COLS] + q[3 * THIN_COLS]);
        const int grp = (int)(((int64_t)r0 * groups) / rows);
        atomicAdd(stats + bn_sum_index(0, groups, grp, which, Cc, jj % Cc), t);
      }
    }
  }
}

// dW[i][j] += sum_r x[r][i] * dy[r][j]  (blockIdx.y: chunk of 32 input features);  db[j] += sum_r dy[r][j].
// Thread (c, g) takes rows r = g, g+4, ...; partial sums meet in shared memory.
template <typename TX, typename TD>
__global__ void __launch_bounds__(256, 4)        // <= 64 registers: the 512 CTAs of g_h0_lin (128 column blocks x 4 feature chunks) run as ONE wave
thin_wgrad_kernel(const TX* __restrict__ x, const TD* __restrict__ dy, float* __restrict__ dW, float* __restrict__ db, int rows, int in_dim,
                  int out_dim) {
  pdl_grid_sync();
  extern __shared__ __align__(16) float thin_smem[];
  float* red = thin_smem;                                       // [g][i][c]  32 KB
  float* bred = red + 4 * THIN_ROWS * THIN_COLS;                // [g][c]      1 KB
  float* xsw = bred + 4 * THIN_COLS;                            // [r][32] : this block's 32 input features of every row
  const int i0 = blockIdx.y * THIN_ROWS;
  for (int e0 = threadIdx.x; e0 < rows * THIN_ROWS; e0 += 8 * 256) {      // eight loads in flight per thread (a load -> store loop is a
    float xv[8];                                                          // chain of L2 round trips: 8 of them at batch 64)
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int e = e0 + 256 * u, r = e >> 5, i = e & 31;
      xv[u] = (e < rows * THIN_ROWS && i0 + i < in_dim) ? ldf(x + (int64_t)r * in_dim + i0 + i) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int e = e0 + 256 * u;
      if (e < rows * THIN_ROWS) xsw[e] = xv[u];
    }
  }
  __syncthreads();
  const int c = threadIdx.x & (THIN_COLS - 1), g = threadIdx.x >> 6;
  const int j = blockIdx.x * THIN_COLS + c;
  const bool col_ok = j < out_dim;
  float acc[THIN_ROWS], bsum = 0.f;
#pragma unroll
  for (int i = 0; i < THIN_ROWS; ++i) acc[i] = 0.f;
  for (int rb = g; rb < rows; rb += 64) {
    float gv[16];                                                 // 16 rows per load batch (all of them at batch 64)
#pragma unroll
    for (int u = 0; u < 16; ++u) { const int r = rb + 4 * u; gv[u] = (col_ok && r < rows) ? ldf(dy + (int64_t)r * out_dim + j) : 0.f; }
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      const int r = rb + 4 * u;
      if (r < rows) {
        bsum += gv[u];
        const float4* xr = reinterpret_cast<const float4*>(xsw + r * THIN_ROWS);
#pragma unroll
        for (int q = 0; q < THIN_ROWS / 4; ++q) {
          const float4 v = xr[q];
          acc[4 * q] = fmaf(v.x, gv[u], acc[4 * q]); acc[4 * q + 1] = fmaf(v.y, gv[u], acc[4 * q + 1]);
          acc[4 * q + 2] = fmaf(v.z, gv[u], acc[4 * q + 2]); acc[4 * q + 3] = fmaf(v.w, gv[u], acc[4 * q + 3]);
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < THIN_ROWS; ++i) red[(g * THIN_ROWS + i) * THIN_COLS + c] = acc[i];
  bred[g * THIN_COLS + c] = bsum;
  __syncthreads();
  if (!col_ok) return;
#pragma unroll
  for (int k = 0; k < 8; ++k) {                                  // this thread finishes input features g*8 .. g*8+7
    const int i = g * 8 + k;
    if (i0 + i < in_dim) {
      const float sum = red[(0 * THIN_ROWS + i) * THIN_COLS + c] + red[(1 * THIN_ROWS + i) * THIN_COLS + c] +
                        red[(2 * THIN_ROWS + i) * THIN_COLS + c] + red[(3 * THIN_ROWS + i) * THIN_COLS + c];
      dW[(int64_t)(i0 + i) * out_dim + j] += sum;               // single writer per element
    }
  }
  if (db != nullptr && blockIdx.y == 0 && g == 0) db[j] += bred[c] + bred[THIN_COLS + c] + bred[2 * THIN_COLS + c] + bred[3 * THIN_COLS + c];
}

bool thin_linear_ok(int rows, int in_dim, int out_dim) { return in_dim <= THIN_MAXK && out_dim >= 256 && rows <= 1024; }

// a block's 32 rows must lie in one row group for the fused statistics
bool thin_linear_stats_ok(int rows, int in_dim, int out_dim, int Cc, int groups) {
  return thin_linear_ok(rows, in_dim, out_dim) && Cc > 0 && out_dim % Cc == 0 && groups >= 1 && rows % groups == 0 &&
         (groups == 1 || (rows / groups) % THIN_ROWS == 0);
}

int thin_linear_fwd(const void* x, int x_dt, const float* W, const float* bias, void* y, int y_dt, int rows, int in_dim, int out_dim,
                    int act, float ap, cudaStream_t st, double* stats = nullptr, int Cc = 0, int groups = 1) {
  const dim3 grid(ceil_div(out_dim, THIN_COLS), ceil_div(rows, THIN_ROWS));
#define GG_TF(TX, TY) Launch(grid, 256, 0, st)(thin_fwd_kernel<TX, TY>, (const TX*)x, W, bias, (TY*)y, rows, in_dim, out_dim, act, ap, stats, Cc, groups)
  if (x_dt == GG_F32 && y_dt == GG_F32) GG_TF(float, float);
  else if (x_dt == GG_F32) GG_TF(float, bf16);
  else if (y_dt == GG_F32) GG_TF(bf16, float);
  else GG_TF(bf16, bf16);
  return check_launch("thin_fwd");
}

// also accumulates the bias gradient when db != nullptr (same pass over dy)
int thin_linear_wgrad(const void* x, int x_dt, const void* dy, int dy_dt, float* dW, float* db, int rows, int in_dim, int out_dim,
                      cudaStream_t st) {
  const dim3 grid(ceil_div(out_dim, THIN_COLS), ceil_div(in_dim, THIN_ROWS));
  const size_t smem = ((size_t)rows * THIN_ROWS + 4 * THIN_ROWS * THIN_COLS + 4 * THIN_COLS) * sizeof(float);
  GG_REQUIRE(smem <= 160 * 1024, GG_ERR_UNSUPPORTED, "thin_wgrad: too many rows");
#define GG_TW(TX, TD)                                                                                                     \
  do {                                                                                                                    \
    static bool attr = false;                                                                                             \
    if (!attr) { cudaFuncSetAttribute(thin_wgrad_kernel<TX, TD>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024); attr = true; } \
    Launch(grid, 256, smem, st)(thin_wgrad_kernel<TX, TD>, (const TX*)x, (const TD*)dy, dW, db, rows, in_dim, out_dim);  \
  } while (0)
  if (x_dt == GG_F32 && dy_dt == GG_F32) GG_TW(float, float);
  else if (x_dt == GG_F32) GG_TW(float, bf16);
  else if (dy_dt == GG_F32) GG_TW(bf16, float);
  else GG_TW(bf16, bf16);
  return check_launch("thin_wgrad");
}

int skinny_linear_fwd(const void* x, int x_dt, const float* W, const float* bias, void* y, int y_dt, int rows, int in_dim, int out_dim,
                      int act, float ap, cudaStream_t st) {
#define GG_SF(TX, TY) Launch(rows, PW_THREADS, 0, st)(skinny_fwd_kernel<TX, TY>, (const TX*)x, W, bias, (TY*)y, in_dim, out_dim, act, ap)
  if (x_dt == GG_F32 && y_dt == GG_F32) GG_SF(float, float);
  else if (x_dt == GG_F32) GG_SF(float, bf16);
  else if (y_dt == GG_F32) GG_SF(bf16, float);
  else GG_SF(bf16, bf16);
  return check_launch("skinny_fwd");
}
int skinny_linear_dgrad(const void* dy, int dy_dt, const float* W, void* dx, int dx_dt, int rows, int in_dim, int out_dim, cudaStream_t st) {
  const int blocks = pw_blocks((int64_t)rows * in_dim);
#define GG_SD(TD, TO) Launch(blocks, PW_THREADS, 0, st)(skinny_dgrad_kernel<TD, TO>, (const TD*)dy, W, (TO*)dx, rows, in_dim, out_dim)
  if (dy_dt == GG_F32 && dx_dt == GG_F32) GG_SD(float, float);
  else if (dy_dt == GG_F32) GG_SD(float, bf16);
  else if (dx_dt == GG_F32) GG_SD(bf16, float);
  else GG_SD(bf16, bf16);
  return check_launch("skinny_dgrad");
}
int skinny_linear_wgrad(const void* x, int x_dt, const void* dy, int dy_dt, float* dW, int rows, int in_dim, int out_dim, cudaStream_t st) {
  const int kblocks = ceil_div(in_dim, 128);
  const int rchunks = std::max(1, std::min(ceil_div(rows, 8), ceil_div(148 * 2, kblocks)));
  const dim3 blocks(kblocks, rchunks);
#define GG_SW(TX, TD) Launch(blocks, 128, 0, st)(skinny_wgrad_kernel<TX, TD>, (const TX*)x, (const TD*)dy, dW, rows, in_dim, out_dim)
  if (x_dt == GG_F32 && dy_dt == GG_F32) GG_SW(float, float);
  else if (x_dt == GG_F32) GG_SW(float, bf16);
  else if (dy_dt == GG_F32) GG_SW(bf16, float);
  else GG_SW(bf16, bf16);
  return check_launch("skinny_wgrad");
}
}  // namespace gg

extern "C" int gg_lstm_step_fwd(const float* gx, const float* Wh, const float* c_prev, const float* h_prev, float* c_out, float* h_out,
                                float* gates_out, int32_t B, int32_t H, float fb, void* stream) {
  GG_REQUIRE(gx && Wh && c_prev && h_prev && c_out && h_out && gates_out && B > 0 && H > 0, GG_ERR_INVALID, "lstm_fwd: bad argument");
  GG_REQUIRE(H <= 4096, GG_ERR_UNSUPPORTED, "lstm_fwd: H > 4096 unsupported");
  const int threads = std::min(256, ((H + 31) / 32) * 32);
  Launch(B, threads, H * sizeof(float), (cudaStream_t)stream)(lstm_fwd_kernel, gx, Wh, c_prev, h_prev, c_out, h_out, gates_out, H, fb);
  return check_launch("lstm_fwd");
}

extern "C" int gg_lstm_step_bwd(const float* gates, const float* c_prev, const float* c_out, const float* dh, const float* dc,
                                const float* Wh, float* dgates, float* dc_prev, float* dh_prev, int32_t B, int32_t H, float fb,
                                void* stream) {
  GG_REQUIRE(gates && c_prev && c_out && dh && Wh && dgates && dc_prev && dh_prev && B > 0 && H > 0, GG_ERR_INVALID, "lstm_bwd: bad argument");
  GG_REQUIRE(H <= 2048, GG_ERR_UNSUPPORTED, "lstm_bwd: H > 2048 unsupported");
  const int threads = std::min(256, ((H + 31) / 32) * 32);
  Launch(B, threads, 4 * H * sizeof(float), (cudaStream_t)stream)(lstm_bwd_kernel, gates, c_prev, c_out, dh, dc, Wh, dgates, dc_prev, dh_prev, H, fb);
  return check_launch("lstm_bwd");
}
