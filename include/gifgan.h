/* gifgan.h -- C ABI of libgifgan.so: the B200 (sm_100a) kernels behind gif-gan's
 * conv-GAN training step.
 *
 * The reference has no FFI: its operator boundary is the Python module
 * /root/reference/models/recurrent_z/ops.py (plus the tf.nn.* calls made directly by
 * models/recurrent_image/rnn_test/recurrent_DCGAN.py), each of which bottoms out in a
 * TensorFlow-0.12 op.  Every entry point below replaces one of those TF op call sites
 * (forward, and the two gradients TF's autodiff derives for it); the citation after
 * each prototype is the reference line it stands in for.  INTEGRATION.md shows the
 * ctypes stub a maintainer of the reference would add.
 *
 * Conventions
 *   - plain C types only; all pointers are DEVICE pointers owned by the caller
 *     (including workspaces); the library never allocates or frees device memory
 *     in a compute call, never synchronises, and enqueues all work on `stream`
 *     (a cudaStream_t passed as void*), so calls are CUDA-graph-capturable;
 *   - return 0 on success, a negative gg_status otherwise; gg_last_error() gives the
 *     thread-local message.  Unsupported shapes/dtypes are ERRORS, never fallbacks;
 *   - activations are NHWC / NDHWC, channel-contiguous; filters are
 *     [taps..., C_large, C_small] (see gg_conv_desc), linear Matrix is [in, out];
 *   - re-entrant: may be called from PyTorch's autograd thread.
 */
#ifndef GIFGAN_H_
#define GIFGAN_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GG_VERSION 100

typedef enum gg_status {
  GG_OK = 0,
  GG_ERR_INVALID = -1,      /* bad argument / inconsistent descriptor            */
  GG_ERR_UNSUPPORTED = -2,  /* shape or dtype the sm_100a kernels do not cover   */
  GG_ERR_CUDA = -3,         /* a CUDA runtime / driver call failed               */
  GG_ERR_WORKSPACE = -4     /* caller-provided workspace too small               */
} gg_status;

typedef enum gg_dtype { GG_F32 = 0, GG_BF16 = 1 } gg_dtype;

typedef enum gg_act {
  GG_ACT_NONE = 0,
  GG_ACT_RELU = 1,    /* tf.nn.relu                          model.py:307            */
  GG_ACT_LRELU = 2,   /* ops.lrelu = tf.maximum(x, a*x)      ops.py:103-104 (d/dx=1 at 0) */
  GG_ACT_TANH = 3,    /* tf.nn.tanh                          model.py:324            */
  GG_ACT_SIGMOID = 4, /* tf.nn.sigmoid                       model.py:344            */
  GG_ACT_TANH01 = 5   /* (tanh(x)+1)/2                       recurrent_DCGAN.py:225  */
} gg_act;

/* One strided-convolution relation between a LARGE grid [N,D,H,W,C] and a SMALL grid
 * [N,Do,Ho,Wo,K]:   i = s*o + t - pad_lo   per spatial dim, filter w[kd,kh,kw,C,K].
 *   conv2d / conv3d  : large = input  (C=Cin),  small = output (K=Cout), w is HWIO / DHWIO
 *   deconv2d         : large = output (C=Cout), small = input  (K=Cin),  w is [kh,kw,Cout,Cin]
 * i.e. the filter layout of ops.py is [taps, C_large, C_small] in both cases.
 * 2-D problems set D=Do=kd=sd=1, pd=0.  pad_lo follows TF 'SAME' (total//2).          */
typedef struct gg_conv_desc {
  int32_t N;
  int32_t D, H, W, C;     /* large grid and its channels */
  int32_t Do, Ho, Wo, K;  /* small grid and its channels */
  int32_t kd, kh, kw;
  int32_t sd, sh, sw;
  int32_t pd, ph, pw;     /* pad_lo */
  int32_t large_dtype;    /* gg_dtype of the large-side activation tensor */
  int32_t small_dtype;    /* gg_dtype of the small-side activation tensor */
  int32_t act;            /* gg_act fused into the epilogue of the op that WRITES an activation */
  float act_param;        /* lrelu leak */
  int32_t flags;          /* GG_CONV_* */
} gg_conv_desc;

#define GG_CONV_ACCUMULATE 1 /* wgrad: dw += (default for wgrad; dw must be initialised) */
#define GG_CONV_TENSOR_CORE 2 /* use the tcgen05 path: bf16 operand activations, `w` = the bf16 copy of the filter
                               * (same [taps,C,K] layout as the fp32 master; gg_cast or the shadow kept by gg_adam*) */

int gg_version(void);
const char* gg_last_error(void);
/* Number of SMs / compute capability of the current device: (major*10+minor) or <0.   */
int gg_device_arch(void);

/* ---- the three contractions of a strided-conv relation -------------------------------
 * SIMT fp32-accumulate path: w, bias, dw are fp32.  Tensor-core path
 * (GG_CONV_TENSOR_CORE): w is the bf16 copy of the SAME [taps,C,K] array -- no transposed or
 * re-packed filter exists; the kernels read it K-major (conv_up) or MN-major (conv_down)
 * through TMA tensor maps, and conv_up with 4 stride classes x 64 channels assembles its
 * N = 256 class-concatenated tile from four boxes of it.                                 */

/* small[o,k] = act( sum_{t,c} large[s*o+t-p, c] * w[t,c,k] + bias[k] )
 * replaces tf.nn.conv2d + bias_add (ops.py:57-60), tf.nn.conv3d (ops.py:70-73), and the
 * input-gradient of tf.nn.conv2d_transpose (ops.py:86).                                 */
int gg_conv_down(const gg_conv_desc* d, const void* large, const void* w, const float* bias,
                 void* small, void* stream);

/* large[i,c] = act( sum_{t,o: s*o+t-p=i} sum_k small[o,k] * w[t,c,k] + bias[c] )
 * replaces tf.nn.conv2d_transpose + bias_add (ops.py:86-95) and the input-gradient of
 * tf.nn.conv2d / tf.nn.conv3d.                                                          */
int gg_conv_up(const gg_conv_desc* d, const void* small, const void* w, const float* bias,
               void* large, void* stream);

/* dw[t,c,k] += sum_o large[s*o+t-p, c] * small[o,k]      (fp32, accumulating)
 * replaces the filter-gradient of tf.nn.conv2d / conv3d / conv2d_transpose.             */
int gg_conv_wgrad(const gg_conv_desc* d, const void* large, const void* small, float* dw,
                  void* stream);
/* the same plus the bias gradient of a `down` conv (ops.py:59 bias_add), dbias[k] += sum_o small[o,k]: one launch on the
 * image-side layer (d_h0_conv, model.py:273), where it rides in an unused row of the filter-gradient GEMM.              */
int gg_conv_wgrad_bias(const gg_conv_desc* d, const void* large, const void* small, float* dw, float* dbias,
                       void* stream);

/* conv + batch statistics of its pre-norm output in one call (what `batch_norm(conv2d(...))` needs, ops.py:10-24 after
 * ops.py:57/86): stats (gg_bn_workspace_bytes(channels, groups) bytes of fp64, [replicas][groups][2][channels], zero-initialised by the caller) += per-channel (sum, sum of squares)
 * over each of `groups` equal blocks of the batch.  On the tensor-core path the sums are produced in the GEMM epilogue
 * (the tile is already in shared memory); otherwise by a coalesced pass.  Consumed by gg_bn_fwd_train_stats.            */
int gg_conv_down_stats(const gg_conv_desc* d, const void* large, const void* w, const float* bias, void* small,
                       double* stats, int32_t groups, void* stream);
int gg_conv_up_stats(const gg_conv_desc* d, const void* small, const void* w, const float* bias, void* large,
                     double* stats, int32_t groups, void* stream);

/* Input gradient of a conv (up != 0: gg_conv_up) or deconv (up == 0: gg_conv_down), no bias / activation, whose INPUT was the
 * output of a train-mode batch norm + activation (`lrelu(d_bn1(conv2d(...)))` feeding the next conv2d, model.py:274-276): the
 * launch that writes dy for that batch norm also accumulates its backward reductions -- per channel and row group, sum g and
 * sum g*xhat with g = dy*act'(gamma*xhat+beta), xhat = (pre-mean)*rstd -- into sums ([groups][2][C] fp64, zeroed by the
 * caller) from the accumulator registers.  pre / save_mean / save_rstd / gamma / beta are that batch norm's (gg_bn_fwd_train).
 * *fused = 1 when the reductions were produced (tensor-core path, tiles aligned with the row groups): gg_bn_bwd(train = 3)
 * then runs its apply pass only; *fused = 0: only the plain dgrad ran.                                                     */
int gg_conv_dgrad_bnbwd(const gg_conv_desc* d, int32_t up, const void* dy, const void* w, void* dx, const float* pre,
                        const float* save_mean, const float* save_rstd, const float* gamma, const float* beta, int32_t act,
                        float act_param, int32_t groups, double* sums, int32_t* fused, void* stream);

/* Reference-named aliases (same arguments, fixed direction). */
int gg_conv2d_fwd(const gg_conv_desc* d, const void* x, const void* w, const float* b, void* y, void* stream);     /* ops.py:57  */
int gg_conv2d_dgrad(const gg_conv_desc* d, const void* dy, const void* w, void* dx, void* stream);                   /* grad of ops.py:57 wrt input_ */
int gg_conv2d_wgrad(const gg_conv_desc* d, const void* x, const void* dy, float* dw, void* stream);                  /* grad of ops.py:57 wrt w */
int gg_deconv2d_fwd(const gg_conv_desc* d, const void* x, const void* w, const float* b, void* y, void* stream);   /* ops.py:86  */
int gg_deconv2d_dgrad(const gg_conv_desc* d, const void* dy, const void* w, void* dx, void* stream);                 /* grad of ops.py:86 wrt input_ */
int gg_deconv2d_wgrad(const gg_conv_desc* d, const void* x, const void* dy, float* dw, void* stream);                /* grad of ops.py:86 wrt w */
int gg_conv3d_fwd(const gg_conv_desc* d, const void* x, const void* w, const float* b, void* y, void* stream);     /* ops.py:70  */
int gg_conv3d_dgrad(const gg_conv_desc* d, const void* dy, const void* w, void* dx, void* stream);
int gg_conv3d_wgrad(const gg_conv_desc* d, const void* x, const void* dy, float* dw, void* stream);

/* ---- linear (ops.py:106-117: tf.matmul(input_, Matrix) + bias) ---------------------- */
int gg_linear_fwd(const void* x, int32_t x_dtype, const float* matrix, const float* bias, void* y, int32_t y_dtype,
                  int32_t rows, int32_t in_dim, int32_t out_dim, int32_t act, float act_param, void* stream);
/* the same feeding a train-mode batch norm over channel = column % Cc (model.py:304-307: g_h0_lin -> reshape -> g_bn0):
 * y is the fp32 pre-norm tensor and stats[groups][2][Cc] (fp64, zeroed by the caller) receives (sum, sum of squares) per
 * row group -- for gg_bn_fwd_train_stats -- from the same launch where the thin kernel applies, else from a statistics pass */
int gg_linear_fwd_stats(const void* x, int32_t x_dtype, const float* matrix, const float* bias, float* y, int32_t rows, int32_t in_dim,
                        int32_t out_dim, int32_t Cc, int32_t groups, double* stats, void* stream);
int gg_linear_dgrad(const void* dy, int32_t dy_dtype, const float* matrix, void* dx, int32_t dx_dtype,
                    int32_t rows, int32_t in_dim, int32_t out_dim, void* stream);
/* dmatrix += x^T dy ; dbias += colsum(dy) (either may be NULL) */
int gg_linear_wgrad(const void* x, int32_t x_dtype, const void* dy, int32_t dy_dtype, float* dmatrix, float* dbias,
                    int32_t rows, int32_t in_dim, int32_t out_dim, void* stream);

/* ---- thin linear + train-mode batch norm + activation in one launch per direction --------------------------------------
 * The generator's input projection relu(g_bn0(reshape(linear(z, 8192), [-1,4,4,512]))) (model.py:304-307): x [rows, in_dim]
 * with in_dim <= 128 and rows <= 128, Matrix [in_dim, out_dim]; the batch norm normalises channel c = column % C over
 * rows * (out_dim / C) values.  A CTA owns whole channels (all their columns, all rows), so statistics and backward
 * reductions stay inside it.  gg_linear_bn_ok() != 0 says the shape is eligible (otherwise the calls return
 * GG_ERR_UNSUPPORTED and the caller composes gg_linear_* with gg_bn_*).
 * fwd: pre [rows, out_dim] fp32 (kept for backward), y = act(bn(pre)) in y_dtype, save_mean / save_rstd [C], EMAs updated.
 * bwd: dmatrix += x^T dpre, dgamma +=, dbeta += (NULL: skipped); dpre is rounded to bf16 first when round_bf16 != 0 (what
 *      the two-kernel path does when activations are bf16).  The producer's bias gradient is exactly zero in train mode. */
int gg_linear_bn_ok(int32_t rows, int32_t in_dim, int32_t out_dim, int32_t C, int32_t groups);
int gg_linear_bn_fwd(const void* x, int32_t x_dtype, const float* matrix, const float* bias, const float* gamma, const float* beta,
                     float* moving_mean, float* moving_var, float* pre, void* y, int32_t y_dtype, float* save_mean, float* save_rstd,
                     int32_t rows, int32_t in_dim, int32_t out_dim, int32_t C, float eps, float decay, int32_t act, float act_param,
                     void* stream);
int gg_linear_bn_bwd(const void* x, int32_t x_dtype, const float* pre, const void* dy, int32_t dy_dtype, const float* gamma, const float* beta,
                     const float* save_mean, const float* save_rstd, float* dmatrix, float* dgamma, float* dbeta, int32_t rows,
                     int32_t in_dim, int32_t out_dim, int32_t C, int32_t act, float act_param, int32_t round_bf16, void* stream);

/* ---- batch norm (ops.py:10-24: tf.contrib.layers.batch_norm; recurrent_DCGAN.py:190-191) ----
 * x is [rows, C] (all leading axes flattened).  groups>1 normalises `groups` equal row
 * blocks independently (the reference's separate D(real)/D(fake) calls batched into one
 * launch); EMA updates are then applied group after group, as sequential calls would.
 * gamma/beta may be NULL (affine-free rnn_test variant); moving_* may be NULL (no EMA).
 * save_mean/save_rstd: [groups, C] fp32, consumed by gg_bn_bwd.
 * ws: >= gg_bn_workspace_bytes(C, groups) bytes, caller-owned.                          */
/* fp64 accumulators [replicas][groups][2][C] (replicas = 1 in this build; the library may replicate them to spread per-CTA
 * atomics, and every kernel that reads them adds the replicas).  The same size is expected for the `stats` buffer of
 * gg_conv_down_stats / gg_conv_up_stats / gg_bn_fwd_train_stats. */
size_t gg_bn_workspace_bytes(int32_t C, int32_t groups);
int gg_bn_fwd_train(const void* x, int32_t x_dtype, void* y, int32_t y_dtype, int64_t rows, int32_t C, int32_t groups,
                    const float* gamma, const float* beta, float* moving_mean, float* moving_var,
                    float* save_mean, float* save_rstd, float eps, float decay, int32_t act, float act_param,
                    void* ws, size_t ws_bytes, void* stream);
/* gg_bn_fwd_train without the statistics pass: `stats` = the (replicated) [groups][2][C] sums written by gg_conv_*_stats */
int gg_bn_fwd_train_stats(const void* x, int32_t x_dtype, void* y, int32_t y_dtype, int64_t rows, int32_t C, int32_t groups,
                          const float* gamma, const float* beta, float* moving_mean, float* moving_var,
                          float* save_mean, float* save_rstd, float eps, float decay, int32_t act, float act_param,
                          const double* stats, void* stream);
int gg_bn_fwd_infer(const void* x, int32_t x_dtype, void* y, int32_t y_dtype, int64_t rows, int32_t C,
                    const float* gamma, const float* beta, const float* moving_mean, const float* moving_var,
                    float eps, int32_t act, float act_param, void* stream);
/* dx = d(loss)/dx given dy = d(loss)/d(act(bn(x))).  train=1: batch statistics
 * (gradient flows through mean/var); train=0: inference statistics in save_mean/save_rstd
 * as written by gg_bn_infer_stats.  dgamma/dbeta (+=, may be NULL).                      */
/* train: 0 = inference-mode statistics (dx = gamma*rstd*g), 1 = batch statistics, 2 = batch statistics with a workspace
 * the caller has already zeroed (saves one memset per layer when a whole update shares one zeroed arena), 3 = batch
 * statistics with the reductions already in ws (written by gg_conv_dgrad_bnbwd): apply pass only. */
int gg_bn_bwd(const void* x, int32_t x_dtype, const void* dy, int32_t dy_dtype, void* dx, int32_t dx_dtype,
              int64_t rows, int32_t C, int32_t groups, const float* gamma, const float* beta,
              const float* save_mean, const float* save_rstd, float* dgamma, float* dbeta,
              int32_t act, float act_param, int32_t train, void* ws, size_t ws_bytes, void* stream);
int gg_bn_infer_stats(const float* moving_mean, const float* moving_var, float eps, int32_t C,
                      float* save_mean, float* save_rstd, void* stream);

/* ---- pointwise / reductions -------------------------------------------------------- */
/* dx = dy * act'(.) evaluated from the activation OUTPUT y (lrelu, relu, tanh, sigmoid, tanh01) */
int gg_act_bwd(const void* y, int32_t y_dtype, const void* dy, int32_t dy_dtype, void* dx, int32_t dx_dtype,
               int64_t n, int32_t act, float act_param, void* stream);
int gg_act_fwd(const void* x, int32_t x_dtype, void* y, int32_t y_dtype, int64_t n, int32_t act, float act_param, void* stream);
/* gg_act_bwd over a channel-last [rows, C] tensor + gg_bias_grad of the result (db[c] += sum_rows dx[r,c]) -- one launch for fp32
 * tensors with C <= 4 (g_h4: tanh' of the generated image and the 3-channel bias gradient, model.py:321-324), else the two calls */
int gg_act_bwd_bias(const void* y, int32_t y_dtype, const void* dy, int32_t dy_dtype, void* dx, int32_t dx_dtype, int64_t rows,
                    int32_t C, int32_t act, float act_param, float* db, void* stream);
/* db[c] += sum_rows dy[r,c]   (gradient of tf.nn.bias_add, ops.py:60,95) */
int gg_bias_grad(const void* dy, int32_t dy_dtype, float* db, int64_t rows, int32_t C, void* stream);
int gg_cast(const void* src, int32_t src_dtype, void* dst, int32_t dst_dtype, int64_t n, void* stream);
/* y = a*x + b*y  (fp32; gradient accumulation where a tensor feeds two consumers) */
int gg_axpby(const float* x, float a, float* y, float b, int64_t n, void* stream);
/* dst[i] = *srcs[i], i < n <= 8: `srcs` is a HOST array of device pointers (NULL entries are skipped) -- the scalar losses
 * of a train step (model.py:241-243) collected with one launch for a single device->host read */
int gg_gather_scalars(const float* const* srcs, int32_t n, float* dst, void* stream);
/* ops.get_std (ops.py:125-128): sqrt(mean_f(var_batch(x[B,F]))) -> out[0]; ws >= 2*F*8 bytes */
int gg_get_std(const void* x, int32_t x_dtype, int64_t B, int64_t F, float* out, void* ws, size_t ws_bytes, void* stream);

/* ---- loss (model.py:121-126: reduce_mean(sigmoid_cross_entropy_with_logits)) --------
 * loss_out[0] (+)= weight * mean_i CE(logits_i, target);  dlogits_i = weight*(sigmoid(x_i)-target)/n
 * (dlogits may be NULL for forward-only).  accumulate!=0 adds into loss_out.             */
int gg_sigmoid_ce(const float* logits, int64_t n, float target, float weight, float* loss_out, int32_t accumulate,
                  float* dlogits, void* stream);
/* Discriminator loss head as two launches (model.py:277 `linear(reshape(h3,[B,-1]), 1, 'd_h3_lin')` + model.py:121-131 the
 * cross-entropy means; z_model_lib.py:416 dvideo_h4 likewise).
 * fwd: logits[r] = h[r,:].w + bias[0];  parts[1+i] = weight_i * mean_{seg_begin_i <= r < seg_end_i} CE(logits[r], target_i),
 *      parts[0] = sum_i parts[1+i];  dlogits[r] = weight_i (sigmoid(logits[r]) - target_i) / n_i   (dlogits may be NULL).
 *      The segment arrays are HOST arrays (read at call time), 1..4 segments.  ticket: 4 bytes of device memory zero-filled
 *      ONCE by the caller (the kernel hands it back zeroed).  Same numbers as gg_linear_fwd + gg_sigmoid_ce.
 * bwd: dW[k] += sum_r dlogits[r] h[r,k];  dbias[0] += sum_r dlogits[r];  dh[r,k] = dlogits[r] w[k]  (dh in h's dtype; dW /
 *      dbias / dh may be NULL).  When h is the output of a train-mode batch norm over channel = column % C (pre, save_mean,
 *      save_rstd, sums non-NULL; in_dim % C == 0, rows % groups == 0) the launch also accumulates that batch norm's backward
 *      reductions (sum g, sum g*xhat; g = dh * act'(gamma*xhat + beta)) into sums[groups][2][C] (zeroed by the caller) and sets
 *      *fused = 1, so gg_bn_bwd(train = 3) can skip its reduction pass -- as gg_conv_dgrad_bnbwd does for conv layers.
 * Needs in_dim % 64 == 0 (gg_loss_head_ok); otherwise GG_ERR_UNSUPPORTED and the caller composes the separate calls. */
int gg_loss_head_ok(int32_t rows, int32_t in_dim, int32_t nsegs);
int gg_loss_head_fwd(const void* h, int32_t h_dtype, const float* w, const float* bias, int32_t rows, int32_t in_dim,
                     const int32_t* seg_begin, const int32_t* seg_end, const float* seg_target, const float* seg_weight, int32_t nsegs,
                     float* logits, float* parts, float* dlogits, void* ticket, void* stream);
int gg_loss_head_bwd(const void* h, int32_t h_dtype, const float* dlogits, const float* w, int32_t rows, int32_t in_dim, float* dW,
                     float* dbias, void* dh, const float* pre, const float* save_mean, const float* save_rstd, const float* gamma,
                     const float* beta, int32_t act, float act_param, int32_t groups, int32_t C, double* sums, int32_t* fused,
                     void* stream);
/* z_model_lib.py:109-111: loss (+)= scalar*mean((a-b)^2); da = scalar*2(a-b)/n  (a strided rows) */
int gg_mse(const float* a, int64_t a_row_stride, const float* b, int64_t b_row_stride, int64_t rows, int64_t cols,
           float scalar, float* loss_out, int32_t accumulate, float* da, void* stream);

/* latent search (z_space_finder.py:258-292, discriminator_activation_optimizer.py:175-204): distance of a generated
 * tensor to a constant fp32 target,  loss_out[0] (+)= w_l2*mean((a-t)^2) + w_l1*mean|a-t|  and, when da != NULL,
 * da = (2*w_l2*(a-t) + w_l1*sign(a-t))/n in a's dtype (the mean over [1,2,3] followed by the mean over the batch is one
 * mean over all n elements).  ws: gg_distance_loss_workspace_bytes() bytes, 16-byte aligned, zero-filled ONCE by the
 * caller before the first call (the kernel leaves it ready for the next); block partials are added in a fixed order. */
size_t gg_distance_loss_workspace_bytes(void);
int gg_distance_loss(const void* a, int32_t a_dtype, const float* target, int64_t n, float w_l2, float w_l1, float* loss_out,
                     int32_t accumulate, void* da, void* ws, size_t ws_bytes, void* stream);

/* ---- input frames (z_model_lib.py:339-346 + utils.py:57-63: the tail of the reference's per-frame decode) --------
 * frames: n decoded uint8 3-channel frames [src_h, src_w, 3] in device memory (frame f at frames + f*frame_stride_bytes, rows
 * row_stride_bytes apart; OpenCV's VideoCapture order is BGR).  out[n, dst_h, dst_w, 3] fp32 =
 *   cv2.resize(frame, (dst_w, dst_h), INTER_LINEAR)  ->  (swap_rb: BGR -> RGB)  ->  x / 127.5 - 1
 * bit for bit: OpenCV's 8-bit fixed-point bilinear incl. its 2x INTER_AREA shortcut (restated in oracle/image_ops.py and pinned
 * against cv2.resize), the normalisation evaluated in float64 and rounded once like numpy's.  One launch per batch; the host
 * ships 1 byte per sample instead of 4.                                                                              */
int gg_frames_to_input(const uint8_t* frames, int32_t n, int32_t src_h, int32_t src_w, int64_t frame_stride_bytes,
                       int64_t row_stride_bytes, float* out, int32_t dst_h, int32_t dst_w, int32_t swap_rb, void* stream);

/* ---- optimiser (model.py:153-156: tf.train.AdamOptimizer(lr, beta1).minimize) -------
 * TF semantics: p -= lr_t * m / (sqrt(v) + eps) with lr_t = lr*sqrt(1-b2^t)/(1-b1^t)
 * computed by the caller.  One launch over a flat fp32 parameter group.
 * grad_scale multiplies g first (1/world_size after a sum all-reduce).
 * p_bf16 (may be NULL): bf16 shadow of p, rewritten in the same pass -- the filter operand of
 * the tensor-core kernels, so no separate re-pack launch follows an update.              */
int gg_adam(float* p, void* p_bf16, const float* g, float* m, float* v, int64_t n, float lr_t, float beta1, float beta2, float eps,
            float grad_scale, void* stream);

/* CUDA-graph-friendly variant: the step counter lives on the device.  state = FOUR int32 (16 bytes; two are used), zero-filled
 * by the caller before the first step: state[0] = t (advanced by this call), state[1] = lr_t (float bits) =
 * lr*sqrt(1-b2^t)/(1-b1^t) recomputed on the device in double precision.                                           */
int gg_adam_graph(float* p, void* p_bf16, const float* g, float* m, float* v, int64_t n, int32_t* state, float lr, float beta1, float beta2,
                  float eps, float grad_scale, void* stream);
/* the two halves of gg_adam_graph as separate calls: gg_adam_tick (t += 1, lr_t) depends only on the group's previous step, so
 * it can be issued early on another stream; gg_adam_apply is the Adam launch alone, reading lr_t from state[1].            */
int gg_adam_tick(int32_t* state, float lr, float beta1, float beta2, void* stream);
int gg_adam_apply(float* p, void* p_bf16, const float* g, float* m, float* v, int64_t n, const int32_t* state, float beta1, float beta2,
                  float eps, float grad_scale, void* stream);

/* ---- BasicLSTMCell (recurrent_DCGAN.py:199-200) --------------------------------------
 * gates_x = x_t @ Matrix[:in] + Bias is batched over T by the caller with gg_linear_fwd;
 * one fused launch per step does  gates = gates_x + h @ Wh;  i,j,f,o;  c',h'.
 * gates_out (post-sum pre-activation) is saved for backward.                             */
int gg_lstm_step_fwd(const float* gates_x, const float* Wh, const float* c_prev, const float* h_prev,
                     float* c_out, float* h_out, float* gates_out, int32_t B, int32_t H, float forget_bias, void* stream);
/* given dh (total gradient wrt h_t) and dc (wrt c_t from t+1): dgates[B,4H], dc_prev (overwritten),
 * dh_prev = dgates @ Wh^T (overwritten).                                                 */
int gg_lstm_step_bwd(const float* gates, const float* c_prev, const float* c_out, const float* dh, const float* dc,
                     const float* Wh, float* dgates, float* dc_prev, float* dh_prev,
                     int32_t B, int32_t H, float forget_bias, void* stream);

/* ---- workspace ------------------------------------------------------------------------
 * The tensor-core conv launches that split their K loop over several CTAs (the 4x4 / 8x8 layers at batch 64: few
 * output tiles, 100-200 K chunks each) exchange fp32 partial tiles through a caller-owned DEVICE buffer: 64 KB of
 * counters followed by a scratch ring.  The buffer must be 1024-byte aligned, ZERO-FILLED when handed over, live until
 * replaced (NULL / 0 = none) and at least gg_workspace_bytes() long for every eligible layer to split; without a
 * workspace those launches run unsplit.  One workspace per process (= per device).  The library never allocates
 * device memory itself, so a captured CUDA graph stays valid as long as the buffer does.
 * (SURVEY.md 8b proposed gg_workspace_bytes; there is no reference counterpart -- TensorFlow's allocator did this.)   */
size_t gg_workspace_bytes(void);
int gg_set_workspace(void* device_buf, size_t bytes);

/* ---- data-parallel gradient exchange (one box, 2..8 GPUs, one process per GPU) ----------
 * Replaces the reference's single-process training loop's implicit "one device" with the exchange north_star asks for
 * ("gg_allreduce_bucket"): the flat fp32 gradient buffer of every rank is mapped into every process with CUDA
 * IPC and ONE kernel per optimiser update sum-all-reduces a range of it in place over NVLink peer loads (two-shot:
 * reduce-scatter in rank order, then all-gather; deterministic, bit-identical on every rank).  Host protocol:
 *   1. every rank: gg_ipc_export() of three device buffers -- grads (the flat gradient buffer), stage (at least
 *      ceil(largest range / world) + 4 floats, 16-byte aligned) and signals (gg_dp_signal_bytes() of ZEROED memory) -- and
 *      sends the (handle, offset) pairs to its peers by any host channel (torch.distributed object all-gather);
 *   2. every rank: gg_ipc_import() of each peer's three buffers (its own stay plain pointers);
 *   3. per update: gg_dp_allreduce(grads[world], stage[world], signals[world], rank, world, lo, n, wire_bf16, stream) on every
 *      rank, same arguments, same order of calls; lo (elements) 16-byte aligned.  Capturable into a CUDA graph (no host state).
 *      wire_bf16 = 1: the reduced sums (fp32 over the ranks' fp32 gradients) are rounded to bf16 before they are distributed --
 *      half the all-gather bytes; every rank, the reducing one included, ends up with the same rounded values.
 * Buffers from a VMM / expandable-segments pool cannot be exported (status GG_ERR_CUDA; callers fall back to NCCL).      */
int gg_ipc_export(const void* device_ptr, void* handle64, uint64_t* offset);
int gg_ipc_import(const void* handle64, uint64_t offset, void** device_ptr);
size_t gg_dp_signal_bytes(void);
int gg_dp_allreduce(void* const* grads, void* const* stage, void* const* signals, int32_t rank, int32_t world, int64_t lo, int64_t n,
                    int32_t wire_bf16, void* stream);

/* ---- introspection for tests/bench -------------------------------------------------- */
/* measurement hook (tools/, bench.py): tensor-core conv calls launch their kernel n times back to back */
void gg_debug_set_repeat(int n);
void gg_debug_set_prof(void* device_buf_512x8_u64);
/* number of kernels this library has launched on any stream since load (monotonic) */
uint64_t gg_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* GIFGAN_H_ */
