// C-ABI entry points of libgifgan.so (declared in include/gifgan.h): error plumbing and
// dispatch between the SIMT fp32-accumulate kernels and the tcgen05 bf16 kernels.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <unordered_set>

#include "common.cuh"

namespace gg {
static thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};

bool pdl_enabled() {
  static const bool on = [] { const char* v = getenv("GG_PDL"); return !(v && v[0] == '0'); }();
  return on;
}

void prefer_max_shared(const void* kernel) {
  static const bool on = [] { const char* v = getenv("GG_CARVEOUT"); return v && v[0] == '1'; }();   // measured: no gain on B200 (1.812 vs 1.832 ms/step) -> opt-in
  if (!on) return;
  static std::mutex mu;
  static std::unordered_set<const void*> done;
  std::lock_guard<std::mutex> lock(mu);
  if (done.insert(kernel).second) {
    (void)cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    (void)cudaGetLastError();
  }
}

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// tapgemm_simt.cu
int simt_conv_down(const gg_conv_desc*, const void*, const float*, const float*, void*, cudaStream_t);
int simt_conv_up(const gg_conv_desc*, const void*, const float*, const float*, void*, cudaStream_t);
int simt_conv_wgrad(const gg_conv_desc*, const void*, const void*, float*, cudaStream_t);
int simt_linear_fwd(const void*, int, const float*, const float*, void*, int, int, int, int, int, float, cudaStream_t);
int simt_linear_dgrad(const void*, int, const float*, void*, int, int, int, int, cudaStream_t);
int simt_linear_wgrad(const void*, int, const void*, int, float*, int, int, int, cudaStream_t);
// pointwise.cu
int skinny_linear_fwd(const void*, int, const float*, const float*, void*, int, int, int, int, int, float, cudaStream_t);
int skinny_linear_dgrad(const void*, int, const float*, void*, int, int, int, int, cudaStream_t);
int skinny_linear_wgrad(const void*, int, const void*, int, float*, int, int, int, cudaStream_t);
bool thin_linear_ok(int rows, int in_dim, int out_dim);
int thin_linear_fwd(const void*, int, const float*, const float*, void*, int, int, int, int, int, float, cudaStream_t, double* stats = nullptr, int Cc = 0, int groups = 1);
bool thin_linear_stats_ok(int rows, int in_dim, int out_dim, int Cc, int groups);
int thin_linear_wgrad(const void*, int, const void*, int, float*, float*, int, int, int, cudaStream_t);
// conv_c3.cu
bool c3_applicable(const gg_conv_desc*);
int c3_conv_down(const gg_conv_desc*, const float*, const float*, const float*, void*, cudaStream_t);
int c3_conv_up(const gg_conv_desc*, const void*, const float*, const float*, float*, cudaStream_t);
int c3_conv_wgrad(const gg_conv_desc*, const float*, const void*, float*, cudaStream_t);
// conv_c3_mma.cu (bf16 mode: warp-level tensor-core MMAs; the fp32 parity mode keeps the SIMT kernels)
int c3m_conv_down(const gg_conv_desc*, const float*, const float*, const float*, void*, cudaStream_t, const float* pre = nullptr,
                  const float* mean = nullptr, const float* rstd = nullptr, const float* gamma = nullptr, const float* beta = nullptr,
                  double* sums = nullptr, int bact = 0, float bact_param = 0.f, int* fused = nullptr);
int c3m_conv_up(const gg_conv_desc*, const void*, const float*, const float*, float*, cudaStream_t);
int c3m_conv_wgrad(const gg_conv_desc*, const float*, const void*, float*, cudaStream_t, float* dbias = nullptr);
static inline bool c3m_applicable(const gg_conv_desc* d) { return c3_applicable(d) && d->small_dtype == GG_BF16; }
// tc_tapgemm.cu
}  // namespace gg
struct gg_bnbwd_args {
  const float *pre, *mean, *rstd, *gamma, *beta;
  double* sums;
  int act;
  float act_param;
  int groups;
};
namespace gg {
int tc_conv_down(const gg_conv_desc*, const void*, const void*, const float*, void*, cudaStream_t, double* stats = nullptr, int groups = 1, int* fused = nullptr,
                 const gg_bnbwd_args* bnb = nullptr);
int tc_conv_up(const gg_conv_desc*, const void*, const void*, const float*, void*, cudaStream_t, double* stats = nullptr, int groups = 1, int* fused = nullptr,
               const gg_bnbwd_args* bnb = nullptr);
// bn.cu
int bn_accumulate_stats(const void* x, int x_dt, int64_t rows, int C, int groups, double* sums, cudaStream_t st);
int tc_conv_wgrad(const gg_conv_desc*, const void*, const void*, float*, cudaStream_t);
bool tc_upcat_ok(const gg_conv_desc*);
int tc_conv_up_cat(const gg_conv_desc*, const void*, const void*, const float*, void*, cudaStream_t, double* stats = nullptr, int groups = 1, int* fused = nullptr,
                   const gg_bnbwd_args* bnb = nullptr);
void tc_set_repeat(int);
void tc_set_prof(void*);
int tc_set_workspace(void*, size_t);
// dp_allreduce.cu
int dp_ipc_export(const void*, void*, uint64_t*);
int dp_ipc_import(const void*, uint64_t, void**);
size_t dp_signal_bytes();
int dp_allreduce(void* const*, void* const*, void* const*, int, int, int64_t, int64_t, int, cudaStream_t);
size_t tc_workspace_bytes();
}  // namespace gg

using namespace gg;

extern "C" int gg_version(void) { return GG_VERSION; }
extern "C" const char* gg_last_error(void) { return g_err; }
extern "C" uint64_t gg_launch_count(void) { return g_launches.load(); }

// measurement hook: every tensor-core conv call launches its kernel n times (host planning amortised)
static int g_cabi_repeat = 1;
extern "C" void gg_debug_set_repeat(int n) { g_cabi_repeat = n < 1 ? 1 : n; tc_set_repeat(n); }
#define GG_REPEAT(call) do { int rc_ = GG_OK; for (int r_ = 0; r_ < g_cabi_repeat && rc_ == GG_OK; ++r_) rc_ = (call); return rc_; } while (0)
// measurement hook: device buffer of 512 x 8 uint64 receiving a per-CTA clock64 breakdown of tc_pixgemm (NULL = off)
extern "C" void gg_debug_set_prof(void* buf) { tc_set_prof(buf); }
// data-parallel gradient exchange over peer memory (see include/gifgan.h)
extern "C" int gg_ipc_export(const void* device_ptr, void* handle64, uint64_t* offset) {
  GG_REQUIRE(device_ptr && handle64 && offset, GG_ERR_INVALID, "gg_ipc_export: bad argument");
  return dp_ipc_export(device_ptr, handle64, offset);
}
extern "C" int gg_ipc_import(const void* handle64, uint64_t offset, void** device_ptr) {
  GG_REQUIRE(handle64 && device_ptr, GG_ERR_INVALID, "gg_ipc_import: bad argument");
  return dp_ipc_import(handle64, offset, device_ptr);
}
extern "C" size_t gg_dp_signal_bytes(void) { return dp_signal_bytes(); }
extern "C" int gg_dp_allreduce(void* const* grads, void* const* stage, void* const* signals, int32_t rank, int32_t world, int64_t lo, int64_t n,
                               int32_t wire_bf16, void* stream) {
  GG_REQUIRE(grads && stage && signals, GG_ERR_INVALID, "gg_dp_allreduce: bad argument");
  return dp_allreduce(grads, stage, signals, rank, world, lo, n, wire_bf16, (cudaStream_t)stream);
}
// workspace of the split-K tensor-core launches (see include/gifgan.h)
extern "C" size_t gg_workspace_bytes(void) { return tc_workspace_bytes(); }
extern "C" int gg_set_workspace(void* device_buf, size_t bytes) { return tc_set_workspace(device_buf, bytes); }

extern "C" int gg_device_arch(void) {
  int dev = 0, major = 0, minor = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { (void)cudaGetLastError(); set_error("no CUDA device"); return GG_ERR_CUDA; }
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  return major * 10 + minor;
}

extern "C" int gg_conv_down(const gg_conv_desc* d, const void* large, const void* w, const float* bias, void* small, void* stream) {
  GG_REQUIRE(d && large && w && small, GG_ERR_INVALID, "conv_down: null pointer");
  if (d->flags & GG_CONV_TENSOR_CORE) return tc_conv_down(d, large, w, bias, small, (cudaStream_t)stream);
  if (c3m_applicable(d)) GG_REPEAT(c3m_conv_down(d, (const float*)large, (const float*)w, bias, small, (cudaStream_t)stream));
  if (c3_applicable(d)) GG_REPEAT(c3_conv_down(d, (const float*)large, (const float*)w, bias, small, (cudaStream_t)stream));
  GG_REPEAT(simt_conv_down(d, large, (const float*)w, bias, small, (cudaStream_t)stream));
}
extern "C" int gg_conv_up(const gg_conv_desc* d, const void* small, const void* w, const float* bias, void* large, void* stream) {
  GG_REQUIRE(d && large && w && small, GG_ERR_INVALID, "conv_up: null pointer");
  if ((d->flags & GG_CONV_TENSOR_CORE) && tc_upcat_ok(d)) return tc_conv_up_cat(d, small, w, bias, large, (cudaStream_t)stream);
  if (d->flags & GG_CONV_TENSOR_CORE) return tc_conv_up(d, small, w, bias, large, (cudaStream_t)stream);
  if (c3m_applicable(d)) GG_REPEAT(c3m_conv_up(d, small, (const float*)w, bias, (float*)large, (cudaStream_t)stream));
  if (c3_applicable(d)) GG_REPEAT(c3_conv_up(d, small, (const float*)w, bias, (float*)large, (cudaStream_t)stream));
  GG_REPEAT(simt_conv_up(d, small, (const float*)w, bias, large, (cudaStream_t)stream));
}
extern "C" int gg_conv_wgrad(const gg_conv_desc* d, const void* large, const void* small, float* dw, void* stream) {
  GG_REQUIRE(d && large && dw && small, GG_ERR_INVALID, "conv_wgrad: null pointer");
  if (d->flags & GG_CONV_TENSOR_CORE) return tc_conv_wgrad(d, large, small, dw, (cudaStream_t)stream);
  if (c3m_applicable(d)) GG_REPEAT(c3m_conv_wgrad(d, (const float*)large, small, dw, (cudaStream_t)stream));
  if (c3_applicable(d)) GG_REPEAT(c3_conv_wgrad(d, (const float*)large, small, dw, (cudaStream_t)stream));
  GG_REPEAT(simt_conv_wgrad(d, large, small, dw, (cudaStream_t)stream));
}

// filter gradient + bias gradient (dbias[k] += sum over the small grid of small[., k]) of a `down` conv.  The image-side warp-MMA
// kernel produces both in one launch (the bias gradient rides in an unused row of its GEMM); otherwise two launches.
extern "C" int gg_conv_wgrad_bias(const gg_conv_desc* d, const void* large, const void* small, float* dw, float* dbias, void* stream) {
  GG_REQUIRE(d && large && dw && small && dbias, GG_ERR_INVALID, "conv_wgrad_bias: null pointer");
  if (!(d->flags & GG_CONV_TENSOR_CORE) && c3m_applicable(d) && g_cabi_repeat == 1)
    return c3m_conv_wgrad(d, (const float*)large, small, dw, (cudaStream_t)stream, dbias);
  int rc = gg_conv_wgrad(d, large, small, dw, stream);
  if (rc) return rc;
  return gg_bias_grad(small, d->small_dtype, dbias, (int64_t)d->N * d->Do * d->Ho * d->Wo, d->K, stream);
}

// conv + batch statistics of its (pre-norm) output: stats[groups][2][channels] += (sum, sum of squares) per row group.
// On the tensor-core path the statistics come out of the GEMM epilogue; otherwise a separate coalesced pass runs.
extern "C" int gg_conv_down_stats(const gg_conv_desc* d, const void* large, const void* w, const float* bias, void* small,
                                  double* stats, int32_t groups, void* stream) {
  GG_REQUIRE(d && large && w && small && stats && groups >= 1, GG_ERR_INVALID, "conv_down_stats: bad argument");
  int fused = 0;
  int rc;
  if (d->flags & GG_CONV_TENSOR_CORE) rc = tc_conv_down(d, large, w, bias, small, (cudaStream_t)stream, stats, groups, &fused);
  else rc = gg_conv_down(d, large, w, bias, small, stream);
  if (rc || fused) return rc;
  const int64_t rows = (int64_t)d->N * d->Do * d->Ho * d->Wo;
  return bn_accumulate_stats(small, d->small_dtype, rows, d->K, groups, stats, (cudaStream_t)stream);
}
extern "C" int gg_conv_up_stats(const gg_conv_desc* d, const void* small, const void* w, const float* bias, void* large,
                                double* stats, int32_t groups, void* stream) {
  GG_REQUIRE(d && large && w && small && stats && groups >= 1, GG_ERR_INVALID, "conv_up_stats: bad argument");
  int fused = 0;
  int rc;
  if ((d->flags & GG_CONV_TENSOR_CORE) && tc_upcat_ok(d)) rc = tc_conv_up_cat(d, small, w, bias, large, (cudaStream_t)stream, stats, groups, &fused);
  else if (d->flags & GG_CONV_TENSOR_CORE) rc = tc_conv_up(d, small, w, bias, large, (cudaStream_t)stream, stats, groups, &fused);
  else rc = gg_conv_up(d, small, w, bias, large, stream);
  if (rc || fused) return rc;
  const int64_t rows = (int64_t)d->N * d->D * d->H * d->W;
  return bn_accumulate_stats(large, d->large_dtype, rows, d->C, groups, stats, (cudaStream_t)stream);
}

// Input gradient of a conv / deconv whose INPUT came out of a train-mode batch norm (+activation): the launch that produces
// dy for that batch norm also accumulates its backward reductions (sum g, sum g*xhat per channel and row group) into `sums`
// from the accumulator registers, so gg_bn_bwd(train = 3) can skip its reduction pass over (pre, dy).  *fused = 1 when the
// kernel did it (tensor-core path, tiles aligned with the row groups); otherwise only the plain dgrad ran.
extern "C" int gg_conv_dgrad_bnbwd(const gg_conv_desc* d, int32_t up, const void* dy, const void* w, void* dx, const float* pre,
                                   const float* save_mean, const float* save_rstd, const float* gamma, const float* beta, int32_t act,
                                   float act_param, int32_t groups, double* sums, int32_t* fused, void* stream) {
  GG_REQUIRE(d && dy && w && dx && fused, GG_ERR_INVALID, "conv_dgrad_bnbwd: null pointer");
  *fused = 0;
  gg_conv_desc c = *d;
  c.act = GG_ACT_NONE;
  if (!(c.flags & GG_CONV_TENSOR_CORE) && !up && pre && save_mean && save_rstd && sums && groups == 1 && c3m_applicable(&c)) {
    // image-side layer (g_h4's dgrad produces dy of g_bn3): warp-MMA kernel with the reductions in its epilogue
    int f = 0, rc = GG_OK;
    for (int r = 0; r < g_cabi_repeat && rc == GG_OK; ++r)          // (measurement hook: gg_debug_set_repeat)
      rc = c3m_conv_down(&c, (const float*)dy, (const float*)w, nullptr, dx, (cudaStream_t)stream, pre, save_mean, save_rstd, gamma, beta, sums,
                         act, act_param, &f);
    *fused = f;
    return rc;
  }
  if (!(c.flags & GG_CONV_TENSOR_CORE) || !pre || !save_mean || !save_rstd || !sums)
    return up ? gg_conv_up(&c, dy, w, nullptr, dx, stream) : gg_conv_down(&c, dy, w, nullptr, dx, stream);
  gg_bnbwd_args b = {pre, save_mean, save_rstd, gamma, beta, sums, act, act_param, groups};
  int f = 0, rc;
  if (up && tc_upcat_ok(&c)) rc = tc_conv_up_cat(&c, dy, w, nullptr, dx, (cudaStream_t)stream, nullptr, 1, &f, &b);
  else if (up) rc = tc_conv_up(&c, dy, w, nullptr, dx, (cudaStream_t)stream, nullptr, 1, &f, &b);
  else rc = tc_conv_down(&c, dy, w, nullptr, dx, (cudaStream_t)stream, nullptr, 1, &f, &b);
  *fused = f;
  return rc;
}

static gg_conv_desc no_act(const gg_conv_desc* d) { gg_conv_desc c = *d; c.act = GG_ACT_NONE; return c; }

extern "C" int gg_conv2d_fwd(const gg_conv_desc* d, const void* x, const void* w, const float* b, void* y, void* s) { return gg_conv_down(d, x, w, b, y, s); }
extern "C" int gg_conv2d_dgrad(const gg_conv_desc* d, const void* dy, const void* w, void* dx, void* s) { GG_REQUIRE(d, GG_ERR_INVALID, "null desc"); gg_conv_desc c = no_act(d); return gg_conv_up(&c, dy, w, nullptr, dx, s); }
extern "C" int gg_conv2d_wgrad(const gg_conv_desc* d, const void* x, const void* dy, float* dw, void* s) { return gg_conv_wgrad(d, x, dy, dw, s); }
extern "C" int gg_deconv2d_fwd(const gg_conv_desc* d, const void* x, const void* w, const float* b, void* y, void* s) { return gg_conv_up(d, x, w, b, y, s); }
extern "C" int gg_deconv2d_dgrad(const gg_conv_desc* d, const void* dy, const void* w, void* dx, void* s) { GG_REQUIRE(d, GG_ERR_INVALID, "null desc"); gg_conv_desc c = no_act(d); return gg_conv_down(&c, dy, w, nullptr, dx, s); }
extern "C" int gg_deconv2d_wgrad(const gg_conv_desc* d, const void* x, const void* dy, float* dw, void* s) { return gg_conv_wgrad(d, dy, x, dw, s); }
extern "C" int gg_conv3d_fwd(const gg_conv_desc* d, const void* x, const void* w, const float* b, void* y, void* s) { return gg_conv_down(d, x, w, b, y, s); }
extern "C" int gg_conv3d_dgrad(const gg_conv_desc* d, const void* dy, const void* w, void* dx, void* s) { GG_REQUIRE(d, GG_ERR_INVALID, "null desc"); gg_conv_desc c = no_act(d); return gg_conv_up(&c, dy, w, nullptr, dx, s); }
extern "C" int gg_conv3d_wgrad(const gg_conv_desc* d, const void* x, const void* dy, float* dw, void* s) { return gg_conv_wgrad(d, x, dy, dw, s); }

extern "C" int gg_linear_fwd(const void* x, int32_t x_dt, const float* matrix, const float* bias, void* y, int32_t y_dt, int32_t rows,
                             int32_t in_dim, int32_t out_dim, int32_t act, float ap, void* stream) {
  GG_REQUIRE(x && matrix && y && rows > 0 && in_dim > 0 && out_dim > 0, GG_ERR_INVALID, "linear_fwd: bad argument");
  if (out_dim <= 4) return skinny_linear_fwd(x, x_dt, matrix, bias, y, y_dt, rows, in_dim, out_dim, act, ap, (cudaStream_t)stream);
  if (thin_linear_ok(rows, in_dim, out_dim)) return thin_linear_fwd(x, x_dt, matrix, bias, y, y_dt, rows, in_dim, out_dim, act, ap, (cudaStream_t)stream);
  return simt_linear_fwd(x, x_dt, matrix, bias, y, y_dt, rows, in_dim, out_dim, act, ap, (cudaStream_t)stream);
}
// linear feeding a train-mode batch norm over channel = column % Cc: y is the fp32 pre-norm tensor; stats[groups][2][Cc] (zeroed by
// the caller) receives (sum, sum of squares) per row group -- from the same launch on the thin path, else from a statistics pass
extern "C" int gg_linear_fwd_stats(const void* x, int32_t x_dt, const float* matrix, const float* bias, float* y, int32_t rows, int32_t in_dim,
                                   int32_t out_dim, int32_t Cc, int32_t groups, double* stats, void* stream) {
  GG_REQUIRE(x && matrix && y && stats && rows > 0 && in_dim > 0 && out_dim > 0 && Cc > 0 && out_dim % Cc == 0 && groups >= 1, GG_ERR_INVALID,
             "linear_fwd_stats: bad argument");
  if (thin_linear_stats_ok(rows, in_dim, out_dim, Cc, groups))
    return thin_linear_fwd(x, x_dt, matrix, bias, y, GG_F32, rows, in_dim, out_dim, GG_ACT_NONE, 0.f, (cudaStream_t)stream, stats, Cc, groups);
  int rc = gg_linear_fwd(x, x_dt, matrix, bias, y, GG_F32, rows, in_dim, out_dim, GG_ACT_NONE, 0.f, stream);
  if (rc) return rc;
  return bn_accumulate_stats(y, GG_F32, (int64_t)rows * (out_dim / Cc), Cc, groups, stats, (cudaStream_t)stream);
}
extern "C" int gg_linear_dgrad(const void* dy, int32_t dy_dt, const float* matrix, void* dx, int32_t dx_dt, int32_t rows, int32_t in_dim,
                               int32_t out_dim, void* stream) {
  GG_REQUIRE(dy && matrix && dx && rows > 0 && in_dim > 0 && out_dim > 0, GG_ERR_INVALID, "linear_dgrad: bad argument");
  if (out_dim <= 4) return skinny_linear_dgrad(dy, dy_dt, matrix, dx, dx_dt, rows, in_dim, out_dim, (cudaStream_t)stream);
  return simt_linear_dgrad(dy, dy_dt, matrix, dx, dx_dt, rows, in_dim, out_dim, (cudaStream_t)stream);
}
extern "C" int gg_linear_wgrad(const void* x, int32_t x_dt, const void* dy, int32_t dy_dt, float* dmatrix, float* dbias, int32_t rows,
                               int32_t in_dim, int32_t out_dim, void* stream) {
  GG_REQUIRE(x && dy && rows > 0 && in_dim > 0 && out_dim > 0, GG_ERR_INVALID, "linear_wgrad: bad argument");
  int rc = GG_OK;
  // one pass over dy gives dW and db; the kernel stages 32 input features of EVERY row in shared memory (128 B per row on
  // top of 33 KB), so above 960 rows (e.g. 64 clips x 16 frames through gvideo_0) the generic kernel takes over
  if (dmatrix && out_dim > 4 && rows <= 960 && thin_linear_ok(rows, in_dim, out_dim))
    return thin_linear_wgrad(x, x_dt, dy, dy_dt, dmatrix, dbias, rows, in_dim, out_dim, (cudaStream_t)stream);
  if (dmatrix) {
    if (out_dim <= 4) rc = skinny_linear_wgrad(x, x_dt, dy, dy_dt, dmatrix, rows, in_dim, out_dim, (cudaStream_t)stream);
    else rc = simt_linear_wgrad(x, x_dt, dy, dy_dt, dmatrix, rows, in_dim, out_dim, (cudaStream_t)stream);
  }
  if (rc == GG_OK && dbias) rc = gg_bias_grad(dy, dy_dt, dbias, rows, out_dim, stream);
  return rc;
}
