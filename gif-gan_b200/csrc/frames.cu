// Tail of the reference's frame decode on the GPU (z_model_lib.py:339-346, utils.py:57-63): decoded uint8 BGR frames ->
// cv2.resize(INTER_LINEAR) -> BGR to RGB -> x / 127.5 - 1 -> fp32 NHWC network input.  The reference does this per frame on the
// host inside the step loop and then feeds 4 bytes per sample through the placeholder; here the host hands over the DECODED BYTES
// (1 byte per sample over PCIe) and one launch produces the whole batch.
//
// Bit-exact with OpenCV's 8-bit fixed-point bilinear (resize.cpp: HResizeLinear / VResizeLinear, 11-bit coefficients; the exact
// 2x reduction shortcut to INTER_AREA) -- restated and pinned against cv2.resize itself in oracle/image_ops.py.  Integer / byte
// work, HBM-bound: a thread produces one output pixel (reads 4 source pixels = 12 bytes, writes 12 bytes); the coefficient tables
// of OpenCV are recomputed per thread in the same float / double arithmetic (no device tables: the library allocates nothing).
#include "common.cuh"

namespace gg {

constexpr int FR_THREADS = 256;

struct FrAxis { int s; int c0, c1; };

// resize.cpp: fx = (float)((dx + 0.5) * scale - 0.5); sx = cvFloor(fx); fx -= sx; [horizontal: clamp (sx, fx) at the borders];
// ialpha = saturate_cast<short>(cvRound((1 - fx) * 2048)), saturate_cast<short>(cvRound(fx * 2048))
__device__ __forceinline__ FrAxis fr_axis(int d, int src, double scale, bool clamp) {
  float f = (float)__dsub_rn(__dmul_rn((double)d + 0.5, scale), 0.5);        // separate multiply and subtract (no fused multiply-add)
  int s = (int)floorf(f);
  f = __fsub_rn(f, (float)s);
  if (clamp) {
    if (s < 0) { s = 0; f = 0.f; }
    if (s >= src - 1) { s = src - 1; f = 0.f; }
  }
  int c0 = __float2int_rn(__fmul_rn(__fsub_rn(1.f, f), 2048.f)), c1 = __float2int_rn(__fmul_rn(f, 2048.f));
  c0 = max(-32768, min(32767, c0));
  c1 = max(-32768, min(32767, c1));
  return FrAxis{s, c0, c1};
}

__global__ void __launch_bounds__(FR_THREADS)
frames_to_input_kernel(const uint8_t* __restrict__ src, int n, int H, int W, int64_t frame_stride, int64_t row_stride, float* __restrict__ dst,
                       int dh, int dw, int swap_rb) {
  pdl_grid_sync();
  // utils.py:63  x / 127.5 - 1.  in float64, rounded once to float32: a 256-entry table per block
  __shared__ float lut[256];
  for (int i = threadIdx.x; i < 256; i += FR_THREADS) lut[i] = (float)__dsub_rn(__ddiv_rn((double)i, 127.5), 1.0);
  __syncthreads();
  const bool area2 = (H == 2 * dh) && (W == 2 * dw);
  // cv::resize: inv_scale = (double)dst / src; scale = 1. / inv_scale
  const double sx_scale = __ddiv_rn(1.0, __ddiv_rn((double)dw, (double)W)), sy_scale = __ddiv_rn(1.0, __ddiv_rn((double)dh, (double)H));
  const int64_t total = (int64_t)n * dh * dw;
  for (int64_t i = (int64_t)blockIdx.x * FR_THREADS + threadIdx.x; i < total; i += (int64_t)gridDim.x * FR_THREADS) {
    const int x = (int)(i % dw);
    const int y = (int)((i / dw) % dh);
    const int64_t f = i / ((int64_t)dw * dh);
    const uint8_t* fr = src + f * frame_stride;
    int v[3];
    if (area2) {
      const uint8_t* p0 = fr + (int64_t)(2 * y) * row_stride + (int64_t)(2 * x) * 3;
      const uint8_t* p1 = p0 + row_stride;
#pragma unroll
      for (int c = 0; c < 3; ++c) v[c] = ((int)p0[c] + (int)p0[3 + c] + (int)p1[c] + (int)p1[3 + c] + 2) >> 2;
    } else {
      const FrAxis ax = fr_axis(x, W, sx_scale, true), ay = fr_axis(y, H, sy_scale, false);
      const int x1 = min(ax.s + 1, W - 1);
      const int y0 = max(0, min(H - 1, ay.s)), y1 = max(0, min(H - 1, ay.s + 1));       // vertical rule: clamp the ROW indices
      const uint8_t* r0 = fr + (int64_t)y0 * row_stride;
      const uint8_t* r1 = fr + (int64_t)y1 * row_stride;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const int h0 = (int)r0[ax.s * 3 + c] * ax.c0 + (int)r0[x1 * 3 + c] * ax.c1;
        const int h1 = (int)r1[ax.s * 3 + c] * ax.c0 + (int)r1[x1 * 3 + c] * ax.c1;
        const int o = (((ay.c0 * (h0 >> 4)) >> 16) + ((ay.c1 * (h1 >> 4)) >> 16) + 2) >> 2;
        v[c] = max(0, min(255, o));
      }
    }
    float* o = dst + i * 3;
    if (swap_rb) { o[0] = lut[v[2]]; o[1] = lut[v[1]]; o[2] = lut[v[0]]; }
    else { o[0] = lut[v[0]]; o[1] = lut[v[1]]; o[2] = lut[v[2]]; }
  }
}

// The 2x reduction (a 128 x 128 source for the 64 x 64 network input: OpenCV's INTER_AREA shortcut) with wide accesses: a thread
// produces FOUR output pixels of a row = 24 contiguous source bytes from each of two rows (three 8-byte loads per row) and 48
// output bytes (three 16-byte stores).  Same arithmetic as the byte-at-a-time path.
__global__ void __launch_bounds__(FR_THREADS)
frames_area2_vec_kernel(const uint8_t* __restrict__ src, int n, int64_t frame_stride, int64_t row_stride, float* __restrict__ dst, int dh,
                        int dw, int swap_rb) {
  pdl_grid_sync();
  __shared__ float lut[256];
  for (int i = threadIdx.x; i < 256; i += FR_THREADS) lut[i] = (float)__dsub_rn(__ddiv_rn((double)i, 127.5), 1.0);
  __syncthreads();
  const int qw = dw / 4;
  const int64_t total = (int64_t)n * dh * qw;
  for (int64_t i = (int64_t)blockIdx.x * FR_THREADS + threadIdx.x; i < total; i += (int64_t)gridDim.x * FR_THREADS) {
    const int xq = (int)(i % qw);
    const int y = (int)((i / qw) % dh);
    const int64_t f = i / ((int64_t)qw * dh);
    const uint8_t* p0 = src + f * frame_stride + (int64_t)(2 * y) * row_stride + (int64_t)xq * 24;
    const uint8_t* p1 = p0 + row_stride;
    uint2 a[3], b[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) { a[k] = __ldg(reinterpret_cast<const uint2*>(p0) + k); b[k] = __ldg(reinterpret_cast<const uint2*>(p1) + k); }
    const uint8_t* A = reinterpret_cast<const uint8_t*>(a);
    const uint8_t* B = reinterpret_cast<const uint8_t*>(b);
    float o[12];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const int v = ((int)A[6 * j + c] + (int)A[6 * j + 3 + c] + (int)B[6 * j + c] + (int)B[6 * j + 3 + c] + 2) >> 2;
        o[3 * j + (swap_rb ? 2 - c : c)] = lut[v];
      }
    float4* d4 = reinterpret_cast<float4*>(dst + i * 12);
    d4[0] = make_float4(o[0], o[1], o[2], o[3]);
    d4[1] = make_float4(o[4], o[5], o[6], o[7]);
    d4[2] = make_float4(o[8], o[9], o[10], o[11]);
  }
}

}  // namespace gg

using namespace gg;

extern "C" int gg_frames_to_input(const uint8_t* frames, int32_t n, int32_t src_h, int32_t src_w, int64_t frame_stride_bytes,
                                  int64_t row_stride_bytes, float* out, int32_t dst_h, int32_t dst_w, int32_t swap_rb, void* stream) {
  GG_REQUIRE(frames && out && n > 0 && src_h > 0 && src_w > 0 && dst_h > 0 && dst_w > 0, GG_ERR_INVALID, "frames_to_input: bad argument");
  GG_REQUIRE(row_stride_bytes >= (int64_t)src_w * 3 && frame_stride_bytes >= row_stride_bytes * (src_h - 1) + (int64_t)src_w * 3, GG_ERR_INVALID,
             "frames_to_input: strides smaller than a row / frame of 3-channel bytes");
  if (src_h == 2 * dst_h && src_w == 2 * dst_w && dst_w % 4 == 0 && row_stride_bytes % 8 == 0 && frame_stride_bytes % 8 == 0 &&
      (uintptr_t)frames % 8 == 0 && (uintptr_t)out % 16 == 0) {
    const int64_t items = (int64_t)n * dst_h * (dst_w / 4);
    const int vblocks = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div64(items, FR_THREADS), 148 * 8));
    Launch(vblocks, FR_THREADS, 0, (cudaStream_t)stream)(frames_area2_vec_kernel, frames, (int)n, frame_stride_bytes, row_stride_bytes, out,
                                                        (int)dst_h, (int)dst_w, (int)swap_rb);
    return check_launch("frames_to_input");
  }
  const int64_t total = (int64_t)n * dst_h * dst_w;
  const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div64(total, FR_THREADS), 148 * 8));
  Launch(blocks, FR_THREADS, 0, (cudaStream_t)stream)(frames_to_input_kernel, frames, (int)n, (int)src_h, (int)src_w, frame_stride_bytes,
                                                     row_stride_bytes, out, (int)dst_h, (int)dst_w, (int)swap_rb);
  return check_launch("frames_to_input");
}
