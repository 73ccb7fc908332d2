// tcgen05 / TMEM / TMA implicit-GEMM kernels (bf16 operands, fp32 accumulation in tensor memory)
// for the strided-conv relation of include/gifgan.h -- the dense contractions of gif-gan's hot
// path: tf.nn.conv2d / conv2d_transpose / conv3d forward, dgrad and wgrad
// (/root/reference/models/recurrent_z/ops.py:57,70,86 and their tf.gradients).
//
// Kernel 1  tc_pixgemm  (conv_down, conv_up):   D[pixel, n] = sum_tap sum_kc A_tap[pixel, kc] * W_tap[n, kc]
//   * A tile  = 128 pixels x 64 channels, fetched by ONE 5-D TMA box (64, bw, bh, bd, bn) from the NHWC
//     activation tensor; a tap is just a coordinate shift of the box and out-of-image pixels are
//     zero-filled by the TMA unit (this is the SAME padding).  Stride-2 forward convs read through one
//     tensor map per input-parity view (base offset + doubled strides), so every tap is a dense box.
//     conv_up (deconv fwd / conv dgrad) is decomposed by OUTPUT parity class: each class is a dense
//     stride-1 tap-GEMM over its own 4/6/6/9 taps -- no stride-inserted zero MACs.
//   * W tile  = BN output channels x 64 reduction channels (K-major) from the packed bf16 filter.
//   * both land in shared memory in the 128B-swizzled K-major layout that tcgen05.mma consumes directly.
//   * warp 0 = TMA producer, warp 1 = MMA issuer (single thread) + TMEM owner, warps 2-5 = epilogue
//     (tcgen05.ld 32x32b, +bias, activation, bf16/fp32 stores).  STAGES-deep mbarrier ring.
// Kernel 2  tc_wgrad:   dW[tap, c, k] += sum_pixels large[pixel*s + tap - p, c] * small[pixel, k]
//   * reduction runs over pixels, which is the slow axis of NHWC: both operands are MN-major for UMMA
//     (rows of 128 B = 64 channels per pixel), so the very same TMA boxes feed it, untransposed.
//   * the M tile (128) = two 64-channel "atoms" (tap, c-chunk); accumulators are reduced into the
//     fp32 gradient buffer with red.global.add (split over pixel ranges across CTAs).
#include <algorithm>
#include <mutex>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "tc_common.cuh"

namespace gg {

using namespace tc;

// ------------------------------------------------------------------------------------------------
// host: tensor-map encoding
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  });
  return fn;
}

int encode_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                     const uint32_t* box) {
  return encode_tmap(out, GG_BF16, base, rank, dims, strides_bytes, box);
}

int encode_tmap(CUtensorMap* out, int dtype, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                const uint32_t* box) {
  EncodeTiledFn enc = get_encode();
  GG_REQUIRE(enc != nullptr, GG_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  // cuTensorMapEncodeTiled is a DRIVER call: it needs a context current on the calling thread.  A thread that has only had its
  // device selected (an autograd worker before its first runtime call) has none yet -> CUDA_ERROR_INVALID_CONTEXT (201);
  // one runtime call binds the primary context.
  static thread_local bool ctx_bound = false;
  if (!ctx_bound) { (void)cudaFree(nullptr); ctx_bound = true; }
  cuuint64_t gdim[5], gstr[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) { gdim[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
  CUresult r = enc(out, dtype == GG_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank,
                   const_cast<void*>(base), gdim, gstr, bx, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  GG_REQUIRE(r == CUDA_SUCCESS, GG_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d): rank %d dims [%llu %llu %llu %llu %llu] box [%u %u %u %u %u]",
             (int)r, rank, (unsigned long long)gdim[0], (unsigned long long)(rank > 1 ? gdim[1] : 0), (unsigned long long)(rank > 2 ? gdim[2] : 0),
             (unsigned long long)(rank > 3 ? gdim[3] : 0), (unsigned long long)(rank > 4 ? gdim[4] : 0), bx[0], rank > 1 ? bx[1] : 0,
             rank > 2 ? bx[2] : 0, rank > 3 ? bx[3] : 0, rank > 4 ? bx[4] : 0);
  return GG_OK;
}

// ------------------------------------------------------------------------------------------------
// kernel 1: pixel-major tap GEMM
// ------------------------------------------------------------------------------------------------
constexpr int TC_MAX_TAPS = 32;
constexpr int TC_MAX_CLASSES = 8;
constexpr int TC_MAX_VIEWS = 8;
constexpr int TILE_M = 128;
constexpr int KCHUNK = 64;                 // bf16 elements per 128-byte swizzle row
constexpr int A_STAGE_BYTES = TILE_M * 128; // 16 KB
constexpr int TC_THREADS = 288;             // warps: 0 / 7 = TMA(A) even / odd chunks, 1 = MMA + TMEM owner, 2..5 = epilogue, 6 / 8 = TMA(B) even / odd chunks
constexpr int TC_WG_THREADS = 224;          // tc_wgrad: 0 = TMA(A), 1 = MMA, 2..5 = epilogue, 6 = TMA(B)

struct TcTap {
  int8_t view;              // which A tensor map
  int8_t od, oh, ow;        // box coordinate shift
  int16_t widx;             // tap index in the packed filter
  int16_t pad_;
};
struct TcCatOrigin { int8_t a0, b0, e0, pad_; };
struct TcClass {
  int Md, Mh, Mw;           // M grid (per image) of this class
  int od0, oh0, ow0;        // output coordinate = m * os + o0
  int tap_begin, tap_end;
  int tile_begin;           // first linear tile index of this class
  int tw, th, td, tn;       // number of tiles along each M axis
};
struct TcPixParams {
  CUtensorMap amap[TC_MAX_VIEWS];
  CUtensorMap bmap;
  CUtensorMap omap[TC_MAX_CLASSES];   // output tensor (per output-parity class for conv_up), box = (128 B of channels, bw, bh, bd, bn)
  int nclasses, ntiles_n;   // N tiles (output-channel tiles)
  int bw, bh, bd, bn;       // box extents: bw*bh*bd*bn == 128
  int BN;                   // output-channel tile (UMMA N): 64 / 128 / 256
  int R;                    // reduction channels per tap (multiple of 64)
  int Nout;                 // output channels
  int Mn;                   // images
  int OD, OH, OW;           // output tensor spatial dims
  int osd, osh, osw;
  int act;
  float act_param;
  int out_bf16;
  int stages;
  int cps;                    // 64-wide K chunks per pipeline stage (MMAs per commit = 4 * cps)
  int cat_C;                  // > 0: the N axis is the concatenation of the output parity classes, cat_C channels each (tc_conv_up_cat)
  int bmode;                  // B operand from the bf16 filter copy w[taps][C][K]: 0 = K-major rows (conv_up: N = C, reduce over K),
                              // 1 = MN-major atoms (conv_down: N = K, reduce over C), 2 = K-major, one box per parity class (up_cat)
  int ncat;                   // bmode 2: parity classes
  TcCatOrigin cat_origin[TC_MAX_TAPS];            // bmode 2: shift -> first tap (kd, kh, kw) of its class box (may be < 0: TMA zero-fills)
  int splitk;                 // S: the K loop of an output tile is split over S CTAs (blockIdx.x % S = rank); 1 = no split
  int sk_mode;                // how the S fp32 partials meet: 0 = thread-block cluster + distributed shared memory, 1 = through L2 (sk_*)
  CUtensorMap skmap;          // sk_mode 1: this launch's scratch as a 2-D fp32 tensor [tiles * S * BN/32 * 128 rows][32], box = (32, 128)
  unsigned int* sk_cnt;       // sk_mode 1: [tiles][2] arrive / depart counters, zero at rest
  unsigned long long* prof;   // optional per-CTA clock64 breakdown (tools/tc_sweep.py --prof), 8 slots per CTA
  double* stats;              // optional fused batch-norm statistics [groups][2][Nout] (fp32 output only)
  int stats_groups;
  // Optional fused batch-norm BACKWARD reductions (this launch is the dgrad that produces dy for a train-mode batch norm):
  // per channel and row group, sum g and sum g*xhat with g = dy * act'(gamma*xhat + beta), xhat = (pre - mean) * rstd --
  // what colsum_kernel<MODE 1> computes in a separate pass over (pre, dy).  Here dy is still in registers.
  const float* bnb_pre;       // fp32 pre-norm tensor of the consumer batch norm, same geometry as this launch's output; nullptr = off
  int epi2;                   // 1: a second set of four warps (the idle producers 0 / 6 / 7 and the MMA warp 1: TMEM lane quarters 0, 2, 3, 1)
                              //    takes every other 32-column block of the epilogue (not with split-K)
  int pre_tma;                // 1: the pre-norm tile is staged by TMA through pmap (box = 32 fp32 columns x the tile's 128 pixels)
  uint32_t pre_off;           // pre_tma: byte offset (from the 1 KB-aligned smem base) of the two 16 KB staging buffers
  CUtensorMap pmap[TC_MAX_CLASSES];   // pre_tma: the pre-norm tensor seen like omap[] (per output-parity class)
  const float* bnb_mean;      // [groups][statC]
  const float* bnb_rstd;      // [groups][statC]
  const float* bnb_gamma;     // [statC] or nullptr
  const float* bnb_beta;      // [statC] or nullptr
  double* bnb_sums;           // [groups][2][statC], zeroed by the caller
  int bnb_act;
  float bnb_act_param;
  int8_t cat_o[TC_MAX_CLASSES][4];   // cat: output origin (od0, oh0, ow0) of each parity class
  TcClass cls[TC_MAX_CLASSES];
  TcTap taps[TC_MAX_TAPS];
};

template <bool OUT_BF16>
__device__ __forceinline__ void store_row32(void* out, int64_t elem_off, const float (&v)[32]) {
  if (OUT_BF16) {
    bf16* dst = reinterpret_cast<bf16*>(out) + elem_off;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      uint4 u;
      __nv_bfloat162 p0 = __floats2bfloat162_rn(v[q * 8 + 0], v[q * 8 + 1]), p1 = __floats2bfloat162_rn(v[q * 8 + 2], v[q * 8 + 3]);
      __nv_bfloat162 p2 = __floats2bfloat162_rn(v[q * 8 + 4], v[q * 8 + 5]), p3 = __floats2bfloat162_rn(v[q * 8 + 6], v[q * 8 + 7]);
      u.x = *reinterpret_cast<uint32_t*>(&p0); u.y = *reinterpret_cast<uint32_t*>(&p1);
      u.z = *reinterpret_cast<uint32_t*>(&p2); u.w = *reinterpret_cast<uint32_t*>(&p3);
      reinterpret_cast<uint4*>(dst)[q] = u;
    }
  } else {
    float* dst = reinterpret_cast<float*>(out) + elem_off;
#pragma unroll
    for (int q = 0; q < 8; ++q) reinterpret_cast<float4*>(dst)[q] = make_float4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
  }
}

__global__ void __launch_bounds__(TC_THREADS, 1)
tc_pixgemm_kernel(const __grid_constant__ TcPixParams p, const float* __restrict__ bias, void* __restrict__ out) {
  extern __shared__ uint8_t smem_raw[];
  const long long t_begin = clock64();
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b_chunk_bytes = p.BN * 128;
  const int stage_bytes = p.cps * (A_STAGE_BYTES + b_chunk_bytes);    // [A chunk 0..cps-1 | B chunk 0..cps-1]
  const uint32_t bar_base = smem_base + p.stages * stage_bytes;     // full[s], empty[s], tmem_full, tmem_slot
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (p.stages + s); };
  const uint32_t tmem_full_bar = bar_base + 8u * (2 * p.stages);
  const uint32_t tmem_slot = bar_base + 8u * (2 * p.stages + 1);
  const uint32_t sk_bar = bar_base + 8u * (2 * p.stages + 2);          // split-K through L2: the partials of my slice have landed
  const uint32_t pre_bar0 = bar_base + 8u * (2 * p.stages + 3);        // fused batch-norm backward: pre-norm block landed (buffer 0)
  const uint32_t pre_bar1 = bar_base + 8u * (2 * p.stages + 4) + 16u;  // (buffer 1; sits behind the row mask)

  // ---- which tile ------------------------------------------------------------------------
  const int S = p.splitk;                      // cluster (S, 1, 1): rank in cluster == blockIdx.x % S
  const int tile = (int)blockIdx.x / S;
  const int rank = (int)blockIdx.x - tile * S;
  int ci = 0;
#pragma unroll 1
  for (int c = 1; c < p.nclasses; ++c) if (tile >= p.cls[c].tile_begin) ci = c;
  const TcClass& C = p.cls[ci];
  int t = tile - C.tile_begin;
  const int nt = t % p.ntiles_n; t /= p.ntiles_n;
  const int tw = t % C.tw; t /= C.tw;
  const int th = t % C.th; t /= C.th;
  const int td = t % C.td; t /= C.td;
  const int tn = t;
  const int mw0 = tw * p.bw, mh0 = th * p.bh, md0 = td * p.bd, mn0 = tn * p.bn;
  const int n0 = nt * p.BN;
  const int kchunks = p.R / KCHUNK;
  // split-K: rank r of the cluster reduces the chunk range [c_begin, c_begin + total_chunks) of the tile's K loop
  const int tile_chunks = (C.tap_end - C.tap_begin) * kchunks;
  const int per_rank = (tile_chunks + S - 1) / S;
  const int c_begin = rank * per_rank;
  const int total_chunks = min(per_rank, tile_chunks - c_begin);       // >= 1 (host guarantees (S - 1) * per_rank < tile_chunks)
  const int iters = (total_chunks + p.cps - 1) / p.cps;                // pipeline stages to run; the last may be partial
  const uint32_t tmem_cols = p.BN < 32 ? 32 : p.BN;   // power of two >= 32

  if (warp == 0 && lane == 0) {
    // full: one arrival per producer warp that touches the stage (2 operands x cps chunk slots)
    for (int s = 0; s < p.stages; ++s) { mbar_init(full_bar(s), 2 * p.cps); mbar_init(empty_bar(s), 1); }
    mbar_init(tmem_full_bar, 1);
    mbar_init(sk_bar, 1);
    mbar_init(pre_bar0, 1);
    mbar_init(pre_bar1, 1);
    fence_barrier_init();
    tma_prefetch_desc(&p.bmap);
  }
  if (warp == 1) tmem_alloc(tmem_slot, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  unsigned long long* prof = (p.prof != nullptr && blockIdx.x < 512) ? p.prof + 8 * blockIdx.x : nullptr;
  pdl_grid_sync();            // everything above overlapped the previous kernel's tail; global memory from here on
  const long long t_setup = clock64();

  // The TMA-producer and MMA-issuer roles are one-instruction-stream roles.  Their loops run with the whole warp
  // converged and elect a lane only around the issue (tc::elect_one): loop control and descriptor arithmetic stay
  // on the uniform datapath.  A and B operands have their own producer warps.
  // Role loops.  They are instruction-bound (one warp, dependent uniform-datapath ops: ~400 cycles per TMA issue were
  // measured), so each operand has TWO producer warps taking alternate 64-wide K chunks, and the MMA warp issues a
  // whole stage (cps chunks = 4 * cps MMAs) per barrier wait.  Chunk c lives in stage c / cps, slot c % cps.
  const int cps = p.cps, nst = p.stages;
  if (warp == 0 || warp == 7 || warp == 6 || warp == 8) {
    // ===== TMA producers: warps 0 / 7 -> A operand (activation boxes), warps 6 / 8 -> B operand (filter tiles) =====
    const bool isA = (warp == 0 || warp == 7);
    const int j = (warp == 0 || warp == 6) ? 0 : 1;                 // parity of the chunks this warp loads
    const int slot = (cps == 2) ? j : 0;
    const int sstep = 2 / cps;                                      // stages advanced per trip
    const uint32_t chunk_bytes = isA ? (uint32_t)A_STAGE_BYTES : (uint32_t)b_chunk_bytes;
    const uint32_t slot_off = (isA ? 0u : (uint32_t)(cps * A_STAGE_BYTES)) + (uint32_t)slot * chunk_bytes;
    int s = (cps == 2) ? 0 : j;
    uint32_t ph = 1;                                                // fresh barriers: waiting on parity 1 passes immediately
    if (s >= nst) { s -= nst; ph ^= 1u; }
    int t = C.tap_begin + (c_begin + j) / kchunks, kc = ((c_begin + j) % kchunks) * KCHUNK;
    const int nstage_total = iters;                                 // stages of this tile
    // trips: one per chunk of parity j; with cps == 2 a final odd stage still needs this warp's arrival on the barrier
    const int my_chunks = (total_chunks - j + 1) / 2;
    const int my_trips = (cps == 2) ? nstage_total : my_chunks;
#pragma unroll 1
    for (int i = 0; i < my_trips; ++i) {
      const uint32_t fb = bar_base + 8u * s, eb = bar_base + 8u * (nst + s);
      const uint32_t dst = smem_base + (uint32_t)s * (uint32_t)stage_bytes + slot_off;
      mbar_wait(eb, ph);
      if (i < my_chunks) {
        if (isA) {
          const TcTap tap = p.taps[t];
          if (elect_one()) {
            mbar_expect_tx(fb, chunk_bytes);
            tma_load_5d(dst, &p.amap[tap.view], fb, kc, mw0 + tap.ow, mh0 + tap.oh, md0 + tap.od, mn0);
          }
        } else {
          const int widx = p.taps[t].widx;
          if (elect_one()) {
            mbar_expect_tx(fb, chunk_bytes);
            if (p.bmode == 0) {
              tma_load_3d(dst, &p.bmap, fb, kc, n0, widx);
            } else if (p.bmode == 1) {
              tma_load_4d(dst, &p.bmap, fb, 0, kc, n0 >> 6, widx);           // BN/64 atoms of [64 reduction rows][64 n]
            } else {
              // all parity classes of the shift in ONE box: class (ad, ah, aw) uses tap (a0 + ad, b0 + ah, e0 + aw), so the
              // classes are a (sd, sh, sw) box of the filter's tap axes; taps outside [0, k) are zero-filled by TMA
              const TcCatOrigin o = p.cat_origin[widx];
              tma_load_5d(dst, &p.bmap, fb, kc, 0, o.e0, o.b0, o.a0);
            }
          }
        }
      } else if (elect_one()) {
        mbar_arrive(fb);                                            // odd tail: this slot stays empty
      }
      kc += 2 * KCHUNK;
      while (kc >= p.R) { kc -= p.R; ++t; }
      s += sstep;
      if (s >= nst) { s -= nst; ph ^= 1u; }
    }
    if (prof && warp == 0 && lane == 0) prof[0] = (unsigned long long)(clock64() - t_setup);
  } else if (warp == 1) {
    // ===== MMA issuer =====
    const bool b_mn = (p.bmode == 1);
    const uint32_t idesc = make_idesc_bf16(TILE_M, p.BN, 0, b_mn ? 1 : 0);
    // descriptor = constant high word | (address >> 4): only the low word changes per stage / chunk / k-step
    const uint64_t desc_hi = make_smem_desc(0, 16, 1024);
    // MN-major B (conv_down reads w[tap][c][k] as it lies): 64-wide n atoms 8 KB apart (LBO), 8-row k groups 1 KB apart
    // (SBO); one K=16 step = 16 rows = 2 KB = +128 descriptor units (as both operands of tc_wgrad_kernel)
    const uint64_t desc_hi_b = b_mn ? make_smem_desc(0, 8192, 1024) : desc_hi;
    const uint32_t bstep = b_mn ? 128u : 2u;
    const uint32_t b_off = (uint32_t)(cps * A_STAGE_BYTES);
    int s = 0, left = total_chunks;
    uint32_t ph = 0;
    uint32_t a_addr = smem_base, fb = bar_base, eb = bar_base + 8u * nst;
    uint32_t acc = 0u;
    long long wait_cycles = 0;
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
      if (prof) {
        const long long w0 = clock64();
        mbar_wait(fb, ph);
        wait_cycles += clock64() - w0;
      } else {
        mbar_wait(fb, ph);
      }
      tc_fence_after();
      const bool two = (cps == 2) && (left >= 2);
      left -= cps;
      if (elect_one()) {
        const uint64_t ad = desc_hi | (uint64_t)((a_addr & 0x3FFFFu) >> 4);
        const uint64_t bd = desc_hi_b | (uint64_t)(((a_addr + b_off) & 0x3FFFFu) >> 4);
        umma_bf16(tmem_base, ad, bd, idesc, acc);
#pragma unroll
        for (int k = 1; k < KCHUNK / 16; ++k)                // +32 B per K=16 step -> +2 in 16-byte units (K-major)
          umma_bf16(tmem_base, ad + 2u * k, bd + bstep * k, idesc, 1u);
        if (two) {
          const uint64_t ad2 = ad + (uint64_t)(A_STAGE_BYTES >> 4), bd2 = bd + (uint64_t)(b_chunk_bytes >> 4);
#pragma unroll
          for (int k = 0; k < KCHUNK / 16; ++k) umma_bf16(tmem_base, ad2 + 2u * k, bd2 + bstep * k, idesc, 1u);
        }
        umma_commit(eb);                                   // frees this smem stage when the MMAs retire
        if (it == iters - 1) umma_commit(tmem_full_bar);    // accumulator complete
      }
      acc = 1u;
      a_addr += stage_bytes; fb += 8u; eb += 8u;
      if (++s == nst) { s = 0; ph ^= 1u; a_addr = smem_base; fb = bar_base; eb = bar_base + 8u * nst; }
    }
    if (prof && lane == 0) { prof[2] = (unsigned long long)(clock64() - t_setup); prof[3] = (unsigned long long)wait_cycles; }
  }

  // ===== epilogue: warps 2..5 own TMEM lane quarters (warp % 4) =====
  // With epi2 the producer warps 0 / 6 / 7 and the MMA warp 1 -- idle once their loops are done, and between them they own all
  // four TMEM lane quarters as well -- form a second set that takes the odd 32-column blocks: the epilogue is a per-warp
  // dependent chain (one warp per scheduler, ~900 instructions per block with the fused batch-norm backward at IPC ~0.3), so
  // a second warp on each scheduler nearly halves it.
  const bool use_b = (S == 1) && p.epi2 != 0;
  const bool epiA = (warp >= 2 && warp <= 5);
  const bool epiB = use_b && (warp == 0 || warp == 1 || warp == 6 || warp == 7);
  const bool epi = epiA || epiB;
  const int eset = epiB ? 1 : 0;
  const uint32_t n_epi = use_b ? 256u : 128u;
  const int q = warp & 3;
  const int row = q * 32 + lane;                 // accumulator row = pixel within the tile
  const uint32_t row_off = (uint32_t)row * 128u;
  const uint32_t sw = (uint32_t)(row & 7);
  const uint32_t mask_smem = bar_base + 8u * (2 * p.stages + 4);        // 4 x u32 after the barriers (16-byte aligned)
  const uint32_t bnb_consts = mask_smem + 32u;                          // float4 {rstd, -mean*rstd, gamma, beta} per column of the N tile
  const uint32_t bnb_part = bnb_consts + 16u * (uint32_t)p.BN;          // float [4 warps][2][BN]: per-warp partial reductions
  long long t_acc = 0;
  int r_iw = 0, r_ih = 0, r_id = 0, r_in = 0;
  uint32_t vmask_row = 0;
  if (epi) {
    const int iw = row % p.bw, ih = (row / p.bw) % p.bh, id = (row / (p.bw * p.bh)) % p.bd, in = row / (p.bw * p.bh * p.bd);
    // rows of a partial tile that fall outside the M grid must not enter the fused batch statistics
    const bool row_valid = (mw0 + iw) < C.Mw && (mh0 + ih) < C.Mh && (md0 + id) < C.Md && (mn0 + in) < p.Mn;
    r_iw = iw; r_ih = ih; r_id = id; r_in = in; vmask_row = row_valid ? 1u : 0u;
    const uint32_t vmask = __ballot_sync(0xffffffffu, row_valid);
    if (lane == 0) asm volatile("st.shared.u32 [%0], %1;" ::"r"(mask_smem + 4u * q), "r"(vmask) : "memory");
    if (p.bnb_pre != nullptr) {
      const int statC = p.cat_C ? p.cat_C : p.Nout;
      const int grp = (int)(((long long)mn0 * p.stats_groups) / p.Mn);
      for (int col = epiA ? (warp - 2) * 32 + lane : p.BN; col < p.BN; col += 128) {
        const int ch = (n0 + col) % statC;
        const float mu = __ldg(p.bnb_mean + grp * statC + ch), rs = __ldg(p.bnb_rstd + grp * statC + ch);
        const float ga = p.bnb_gamma ? __ldg(p.bnb_gamma + ch) : 1.f, be = p.bnb_beta ? __ldg(p.bnb_beta + ch) : 0.f;
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(bnb_consts + 16u * (uint32_t)col), "f"(rs), "f"(-mu * rs), "f"(ga), "f"(be) : "memory");
      }
      asm volatile("bar.sync 2, %0;" ::"r"(n_epi) : "memory");      // the epilogue warps (both sets)
    }
    if (lane == 0) mbar_wait(tmem_full_bar, 0);   // one polling lane per warp: the spin must not steal issue slots
    __syncwarp();                                 // from the producer / MMA threads that share these schedulers
    tc_fence_after();
    t_acc = clock64();
  }

  // ---- split-K: push-style reduce-scatter of the S partial accumulators through distributed shared memory ----
  // Rank d of the cluster FINISHES the column slice [d*ncol, (d+1)*ncol) of the tile (ncol = BN / S): every rank reads its
  // fp32 partial from TMEM once and stores each 32-column block straight into the receive buffer of the rank that owns it
  // (st.shared::cluster; its own slice with the same instruction), R[source][block][128 rows][128 B, 16-byte chunk ^ (row & 7)]
  // at smem_base of the owner -- the pipeline stages, free on every rank once ALL ranks have seen their tmem_full
  // (cluster barrier 1).  After cluster barrier 2 a rank only reads its own shared memory: it sums the S partials of its
  // slice and runs the normal epilogue (bias, activation, statistics, TMA store) as if the tile were 128 x ncol.
  // (The first version pulled row slices from all ranks and let the leader alone run the epilogue: ~13 k cycles at S = 8.)
  const int ncol = p.BN / S;                       // columns this rank finishes
  const int ncol0 = rank * ncol;                   // first of them within the N tile
  if (S > 1 && p.sk_mode == 1) {
    // ---- the same reduce-scatter THROUGH L2 (no cluster): deterministic, and ~6 k cycles where the DSMEM exchange took ~20 k ----
    // Every rank parks its fp32 partial in global scratch with TMA stores (BN/32 boxes of [128 rows][32 columns], written from
    // the 128B-swizzled staging this kernel's fp32 epilogue uses anyway), counts itself in, waits until all S ranks of the tile
    // have done so (the launch is ONE wave with one CTA per SM, so they are co-resident; the spin is bounded) and TMA-loads the
    // boxes of ITS column slice from all S partials into R[source][block] -- exactly where the DSMEM variant receives them.
    // The last rank to leave re-arms the counters (they are zero whenever no launch is using them).
    if (epi) {
      for (int c0 = 0; c0 < p.BN; c0 += 32) {
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
        tmem_ld_wait();
        const uint32_t blk = smem_base + (uint32_t)(c0 >> 5) * (TILE_M * 128u) + row_off;
#pragma unroll
        for (int g = 0; g < 8; ++g)
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(blk + ((((uint32_t)g) ^ sw) << 4)), "r"(r[g * 4]), "r"(r[g * 4 + 1]),
                       "r"(r[g * 4 + 2]), "r"(r[g * 4 + 3]) : "memory");
      }
      fence_proxy_async();
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (warp == 2 && lane == 0) {
        const long long k0 = clock64();
        long long k1, k2, k3;
        const int nbk = p.BN >> 5, nb = ncol >> 5;     // 32-column boxes of the tile / of a slice
        for (int b = 0; b < nbk; ++b)
          tma_store_2d(&p.skmap, smem_base + (uint32_t)b * (TILE_M * 128u), 0, ((tile * S + rank) * nbk + b) * TILE_M);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");        // written (not only read): the staging is free again, too
        k1 = clock64();
        fence_proxy_async_all();
        __threadfence();
        unsigned int* cnt = p.sk_cnt + 2 * tile;
        atomicAdd(cnt, 1u);
        if (ld_acquire_gpu(cnt) < (uint32_t)S) {
          const long long t0 = clock64();
          while (ld_acquire_gpu(cnt) < (uint32_t)S) {
            __nanosleep(64);
            if (clock64() - t0 > 4000000000ll) __trap();
          }
        }
        k2 = clock64();
        __threadfence();
        fence_proxy_async_all();
        mbar_expect_tx(sk_bar, (uint32_t)(S * nb) * (TILE_M * 128u));
        for (int src = 0; src < S; ++src)
          for (int b = 0; b < nb; ++b)
            tma_load_2d(smem_base + (uint32_t)(src * nb + b) * (TILE_M * 128u), &p.skmap, sk_bar, 0, ((tile * S + src) * nbk + rank * nb + b) * TILE_M);
        if (atomicAdd(cnt + 1, 1u) == (unsigned int)(S - 1)) { cnt[0] = 0u; cnt[1] = 0u; __threadfence(); }
        if (prof) {   // exchange breakdown, 16 cycles per unit: [staged since accumulator-ready | stores complete | all ranks in | loads landed]
          mbar_wait(sk_bar, 0);
          k3 = clock64();
          auto u16 = [](long long c) { return (unsigned long long)min(65535ll, c >> 4); };
          prof[1] = u16(k0 - t_acc) | (u16(k1 - k0) << 16) | (u16(k2 - k1) << 32) | (u16(k3 - k2) << 48);
        }
      }
      if (lane == 0) mbar_wait(sk_bar, 0);
      __syncwarp();
    }
  } else if (S > 1) {
    if (epi) tc_fence_before();
    cluster_sync_all();                            // 1: every rank's MMAs have retired -> all pipeline buffers are free
    if (epi) {
      tc_fence_after();
      const uint32_t nb = (uint32_t)(ncol >> 5);   // 32-column blocks per slice
      for (int c0 = 0; c0 < p.BN; c0 += 32) {
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
        tmem_ld_wait();
        const uint32_t d = (uint32_t)(c0 / ncol), b = (uint32_t)((c0 % ncol) >> 5);
        const uint32_t dst = dsmem_addr(smem_base + ((uint32_t)rank * nb + b) * (TILE_M * 128u) + row_off, d);
#pragma unroll
        for (int g = 0; g < 8; ++g)
          asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + ((((uint32_t)g) ^ sw) << 4)), "r"(r[g * 4]), "r"(r[g * 4 + 1]),
                       "r"(r[g * 4 + 2]), "r"(r[g * 4 + 3]) : "memory");
      }
      tc_fence_before();
    }
    cluster_sync_all();                            // 2: all partials of my slice have arrived (release / acquire)
  }

  if (epi) {
    // Accumulator (TMEM, or the S partials of this rank's column slice in R when S > 1) -> (+bias, activation) -> shared
    // memory in the 128B-swizzled box layout -> TMA tensor store.  The TMA store clips partial tiles and, for conv_up,
    // scatters to the stride-s output parity view.  The staging of a split-K slice sits behind R (host checks the capacity).
    // Below, c0 counts columns of THIS RANK's slice; cg = ncol0 + c0 is the column within the N tile.
    const uint32_t out_base = (S > 1) ? smem_base + (uint32_t)p.BN * 512u : smem_base;
    // Fused batch-norm backward with pre_tma: the fp32 pre-norm tile of each 32-column block arrives by TMA (the box of the
    // output store, fp32 elements) in one of two 16 KB buffers, two blocks ahead of its use, and a thread reads ITS row with
    // eight conflict-free 16-byte shared loads.  (Reading the row straight from global memory made every LDG.128 touch 32
    // different lines: the epilogue was bound by L1 request throughput, ~2 k cycles per block -- profiles/r4n_*.)
    const bool pre_tma = p.bnb_pre != nullptr && p.pre_tma != 0;
    const uint32_t pre_buf = smem_base + p.pre_off;
    // one set: blocks alternate between the two buffers (two blocks ahead); two sets: each set owns one buffer and refills it
    // as soon as its four warps have read their rows (the block's arithmetic hides the load)
    auto issue_pre = [&](int blk, int buf) {           // one thread of the set
      const int colg = n0 + ncol0 + blk * 32;
      const int oc = p.cat_C ? colg / p.cat_C : ci, ch = p.cat_C ? colg % p.cat_C : colg;
      const uint32_t bar = buf ? pre_bar1 : pre_bar0;
      mbar_expect_tx(bar, TILE_M * 128u);
      tma_load_5d(pre_buf + (uint32_t)buf * (TILE_M * 128u), &p.pmap[oc], bar, ch, mw0, mh0, md0, mn0);
    };
    const bool pre_issuer = lane == 0 && (eset ? warp == 0 : warp == 2);
    const int c_first = use_b ? eset * 32 : 0, c_step = use_b ? 64 : 32;
    if (pre_tma && pre_issuer) {
      if (use_b) {
        if (c_first < ncol) issue_pre(c_first >> 5, eset);
      } else {
        issue_pre(0, 0);
        if (ncol > 32) issue_pre(1, 1);
      }
    }
    int kseq = 0;                                      // how many blocks this set has taken
    for (int c0 = c_first; c0 < ncol; c0 += c_step, ++kseq) {
      const int cg = ncol0 + c0;
      float v[32];
      if (S > 1) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0.f;
        for (int src = 0; src < S; ++src) {
          const uint32_t blk = smem_base + ((uint32_t)src * (uint32_t)(ncol >> 5) + (uint32_t)(c0 >> 5)) * (TILE_M * 128u) + row_off;
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            float4 t;
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(t.x), "=f"(t.y), "=f"(t.z), "=f"(t.w)
                         : "r"(blk + ((((uint32_t)g) ^ sw) << 4)) : "memory");
            v[g * 4] += t.x; v[g * 4 + 1] += t.y; v[g * 4 + 2] += t.z; v[g * 4 + 3] += t.w;
          }
        }
      } else {
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
      }
      if (bias != nullptr) {
        const float4* b4 = reinterpret_cast<const float4*>(bias + (p.cat_C ? (n0 + cg) % p.cat_C : n0 + cg));
#pragma unroll
        for (int j = 0; j < 8; ++j) { const float4 b = __ldg(b4 + j); v[4 * j] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w; }
      }
      act_fwd_vec<32>(v, p.act, p.act_param);
      if (p.out_bf16) {
        // 32 bf16 = 64 B = 4 chunks: half of the 128-byte row of column block c0/64
        const uint32_t blk = out_base + (uint32_t)(c0 >> 6) * (TILE_M * 128u) + row_off;
        const uint32_t ch0 = (uint32_t)((c0 & 63) >> 3);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          __nv_bfloat162 p0 = __floats2bfloat162_rn(v[g * 8 + 0], v[g * 8 + 1]), p1 = __floats2bfloat162_rn(v[g * 8 + 2], v[g * 8 + 3]);
          __nv_bfloat162 p2 = __floats2bfloat162_rn(v[g * 8 + 4], v[g * 8 + 5]), p3 = __floats2bfloat162_rn(v[g * 8 + 6], v[g * 8 + 7]);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(blk + (((ch0 + g) ^ sw) << 4)), "r"(*reinterpret_cast<uint32_t*>(&p0)),
                       "r"(*reinterpret_cast<uint32_t*>(&p1)), "r"(*reinterpret_cast<uint32_t*>(&p2)), "r"(*reinterpret_cast<uint32_t*>(&p3)) : "memory");
        }
      } else {
        // 32 fp32 = 128 B = the whole row of column block c0/32 (in place when the source is P)
        const uint32_t blk = out_base + (uint32_t)(c0 >> 5) * (TILE_M * 128u) + row_off;
#pragma unroll
        for (int g = 0; g < 8; ++g)
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(blk + ((((uint32_t)g) ^ sw) << 4)), "f"(v[g * 4]), "f"(v[g * 4 + 1]),
                       "f"(v[g * 4 + 2]), "f"(v[g * 4 + 3]) : "memory");
      }
      if (p.bnb_pre != nullptr) {
        // ---- fused batch-norm backward reductions over this 32-column block (see TcPixParams::bnb_*) ----
        const int statC = p.cat_C ? p.cat_C : p.Nout;
        const int colg = n0 + cg;                                   // first column of the block in the N axis
        const int oc = p.cat_C ? colg / p.cat_C : ci;               // parity class the block writes (cat: per block)
        const int o_d = p.cat_C ? p.cat_o[oc][0] : C.od0, o_h = p.cat_C ? p.cat_o[oc][1] : C.oh0, o_w = p.cat_C ? p.cat_o[oc][2] : C.ow0;
        const int ch0 = p.cat_C ? colg % p.cat_C : colg;
        const bool rv = (vmask_row != 0);
        float x[32];
        if (pre_tma) {
          const int buf = use_b ? eset : (kseq & 1);
          const uint32_t par = (uint32_t)(use_b ? (kseq & 1) : ((kseq >> 1) & 1));
          if (lane == 0) mbar_wait(buf ? pre_bar1 : pre_bar0, par);
          __syncwarp();
          const uint32_t rowp = pre_buf + (uint32_t)buf * (TILE_M * 128u) + row_off;
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            float4 t;
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(t.x), "=f"(t.y), "=f"(t.z), "=f"(t.w) : "r"(rowp + ((((uint32_t)g) ^ sw) << 4)) : "memory");
            x[4 * g] = t.x; x[4 * g + 1] = t.y; x[4 * g + 2] = t.z; x[4 * g + 3] = t.w;
          }
          if (use_b && c0 + 64 < ncol) {            // two sets: refill this set's buffer now
            if (eset) asm volatile("bar.sync 4, 128;" ::: "memory"); else asm volatile("bar.sync 3, 128;" ::: "memory");
            if (pre_issuer) issue_pre((c0 >> 5) + 2, eset);
          }
        } else {
          const long long pix = ((((long long)(mn0 + r_in) * p.OD + ((md0 + r_id) * p.osd + o_d)) * p.OH + ((mh0 + r_ih) * p.osh + o_h)) * p.OW +
                                 ((mw0 + r_iw) * p.osw + o_w));
          const float4* xp = reinterpret_cast<const float4*>(p.bnb_pre + pix * statC + ch0);
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            const float4 t = rv ? __ldg(xp + g) : make_float4(0.f, 0.f, 0.f, 0.f);
            x[4 * g] = t.x; x[4 * g + 1] = t.y; x[4 * g + 2] = t.z; x[4 * g + 3] = t.w;
          }
        }
        // xhat and the pre-activation u in place (x <- xhat, u kept in a second array only for the slow activations)
        if (p.bnb_act <= GG_ACT_LRELU) {
          // none / relu / lrelu: act'(u) is 1, slope or at_zero -- branch-free selects (ONE warp-uniform branch per block: a
          // per-element call of the general act_grad_from_pre was if-converted by ptxas into evaluating tanhf AND expf for
          // every element: 11 k cycles per 32-column block)
          const float slope = act_slope(p.bnb_act, p.bnb_act_param), at_zero = p.bnb_act == GG_ACT_RELU ? 0.f : 1.f;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float4 k;
            asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(k.x), "=f"(k.y), "=f"(k.z), "=f"(k.w) : "r"(bnb_consts + 16u * (uint32_t)(cg + j)));
            const float xh = fmaf(x[j], k.x, k.y);
            const float u = fmaf(k.z, xh, k.w);
            float gv = p.out_bf16 ? __bfloat162float(__float2bfloat16_rn(v[j])) : v[j];     // what the apply kernel will read back
            gv *= (u > 0.f ? 1.f : (u == 0.f ? at_zero : slope));
            gv = rv ? gv : 0.f;
            v[j] = gv;
            x[j] = gv * xh;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) {        // fully unrolled: a partially unrolled loop indexes v / x dynamically -> local memory
            float4 k;
            asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(k.x), "=f"(k.y), "=f"(k.z), "=f"(k.w) : "r"(bnb_consts + 16u * (uint32_t)(cg + j)));
            const float xh = fmaf(x[j], k.x, k.y);
            const float u = fmaf(k.z, xh, k.w);
            float gv = p.out_bf16 ? __bfloat162float(__float2bfloat16_rn(v[j])) : v[j];
            gv = rv ? gv * act_grad_from_pre(u, p.bnb_act, p.bnb_act_param) : 0.f;
            v[j] = gv;
            x[j] = gv * xh;
          }
        }
        const float s0 = warp_transpose_sum32(v, lane), s1 = warp_transpose_sum32(x, lane);   // lane = column of the block
        // per-warp partials [warp][2][BN] in shared memory; they meet below, ONE pair of fp64 atomics per channel per CTA
        // (atomics straight from the warps -- 8 * BN per CTA -- made the L2 atomic units the bottleneck: +24 us per launch)
        const uint32_t pp = bnb_part + 4u * (uint32_t)((q * 2) * p.BN + cg + lane);
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(pp), "f"(s0) : "memory");
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(pp + 4u * (uint32_t)p.BN), "f"(s1) : "memory");
        (void)ch0;
        if (pre_tma && !use_b && c0 + 64 < ncol) {      // one set: refill this block's buffer with block + 2 (all four warps are done with it)
          asm volatile("bar.sync 3, 128;" ::: "memory");
          if (pre_issuer) issue_pre((c0 >> 5) + 2, kseq & 1);
        }
      }
    }
    fence_proxy_async();                                       // generic-proxy smem writes -> visible to the TMA engine
    asm volatile("bar.sync 1, %0;" ::"r"(n_epi) : "memory");  // the epilogue warps (both sets)
    if (warp == 2 && lane == 0) {
      const int nblk = p.out_bf16 ? (ncol >> 6) : (ncol >> 5);
      const int cstep = p.out_bf16 ? 64 : 32;
      for (int b = 0; b < nblk; ++b) {
        const int col = n0 + ncol0 + b * cstep;                  // concatenated classes: column -> (class output map, channel)
        const int oc = p.cat_C ? col / p.cat_C : ci, ch = p.cat_C ? col % p.cat_C : col;
        tma_store_5d(&p.omap[oc], out_base + (uint32_t)b * (TILE_M * 128u), ch, mw0, mh0, md0, mn0);
      }
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    if (p.bnb_pre != nullptr && epiA) {
      // (the per-warp partials were published by the bar.sync before the TMA stores)
      const int statC = p.cat_C ? p.cat_C : p.Nout;
      const int uniq = p.cat_C ? p.cat_C : ncol;                    // distinct channels among the columns this rank finishes
      const int grp = (int)(((long long)mn0 * p.stats_groups) / p.Mn);
      for (int u = (warp - 2) * 32 + lane; u < uniq; u += 128) {
        float a0 = 0.f, a1 = 0.f;
        for (int col = ncol0 + u; col < ncol0 + ncol; col += uniq)  // cat: the parity classes of the channel
#pragma unroll
          for (int w4 = 0; w4 < 4; ++w4) {
            float t0, t1;
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(t0) : "r"(bnb_part + 4u * (uint32_t)((w4 * 2) * p.BN + col)));
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(t1) : "r"(bnb_part + 4u * (uint32_t)((w4 * 2 + 1) * p.BN + col)));
            a0 += t0; a1 += t1;
          }
        const int schan = (n0 + ncol0 + u) % statC;
        atomicAdd(p.bnb_sums + bn_sum_index(0, p.stats_groups, grp, 0, statC, schan), (double)a0);
        atomicAdd(p.bnb_sums + bn_sum_index(0, p.stats_groups, grp, 1, statC, schan), (double)a1);
      }
    }
    if (p.stats != nullptr && epiA) {
      // Fused batch-norm statistics (tf.nn.moments of the pre-norm tensor): the fp32 tile is in shared memory;
      // epilogue thread t sums column t over the tile's valid rows (a warp reads 32 consecutive floats of one
      // swizzled row per step: conflict-free) and adds (sum, sum of squares) to the fp64 accumulators of its group.
      uint32_t m[4];
      asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(m[0]), "=r"(m[1]), "=r"(m[2]), "=r"(m[3]) : "r"(mask_smem));
      const int grp = (int)(((long long)mn0 * p.stats_groups) / p.Mn);
      const int statC = p.cat_C ? p.cat_C : p.Nout;          // every parity class of a channel feeds the same statistic
      const int uniq = p.cat_C ? p.cat_C : ncol;             // distinct channels among the columns this rank finishes
      for (int u = (warp - 2) * 32 + lane; u < uniq; u += 128) {
        float s = 0.f, s2 = 0.f;
        for (int col = u; col < ncol; col += uniq) {         // (slice-local column) cat: the classes of the channel are folded BEFORE the atomics
          const uint32_t base = out_base + (uint32_t)(col >> 5) * (TILE_M * 128u) + (uint32_t)((col & 3) << 2);
          const uint32_t ch = (uint32_t)((col & 31) >> 2);
#pragma unroll
          for (int w4 = 0; w4 < 4; ++w4) {
            const uint32_t mrow = m[w4];
#pragma unroll 8
            for (int b = 0; b < 32; ++b) {
              const int r = w4 * 32 + b;
              float v;
              asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(base + (uint32_t)r * 128u + ((ch ^ (uint32_t)(r & 7)) << 4)));
              if ((mrow >> b) & 1u) { s += v; s2 = fmaf(v, v, s2); }
            }
          }
        }
        const int R = bn_replicas(statC, p.stats_groups);     // replicated accumulators: spread the same-address atomics
        const int rep = tile & (R - 1), schan = (n0 + ncol0 + u) % statC;
        atomicAdd(p.stats + bn_sum_index(rep, p.stats_groups, grp, 0, statC, schan), (double)s);
        atomicAdd(p.stats + bn_sum_index(rep, p.stats_groups, grp, 1, statC, schan), (double)s2);
      }
    }
    if (warp == 2 && lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // smem may be released after this
    tc_fence_before();
    if (prof && warp == 2 && lane == 0) { prof[4] = (unsigned long long)(t_acc - t_setup); prof[5] = (unsigned long long)(clock64() - t_setup); }
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
  if (prof && threadIdx.x == 0) { prof[6] = (unsigned long long)(clock64() - t_setup); prof[7] = (unsigned long long)(t_setup - t_begin); }
}

// ------------------------------------------------------------------------------------------------
// kernel 2: wgrad (MN-major operands, reduction over pixels)
// ------------------------------------------------------------------------------------------------
constexpr int WG_PIX = 64;                      // pixels (reduction) per stage
constexpr int WG_ATOM_BYTES = WG_PIX * 128;     // one 64-channel atom x 64 pixels = 8 KB

struct WgAtom {           // one 64-row half of an M tile: (tap, channel chunk of the large tensor)
  int8_t view, od, oh, ow;
  int16_t widx;           // tap index in dw
  int16_t c0;             // first channel (multiple of 64); -1 = padding atom
};
struct TcWgradParams {
  CUtensorMap lmap[TC_MAX_VIEWS];   // large tensor views (same maps as conv_down's A operand, box = WG_PIX pixels)
  CUtensorMap smap;                 // small tensor
  CUtensorMap dwmap;                // dw as a 2-D fp32 tensor [taps*C, K], box = (32 columns = 128 B, 64 rows), for the reduce-store
  int natoms, mtiles, ntiles_n, splits;
  int BN;                           // k-channel tile (multiple of 64, <= 256)
  int bw, bh, bd, bn;               // pixel box: product == WG_PIX
  int tw, th, td, tn;               // pixel tiles along each axis of the small grid
  int ptiles, ptiles_per_split;
  int C, K;
  int stages;
  int pps;                          // 64-pixel tiles per pipeline stage (MMAs per commit = 4 * pps)
  WgAtom atoms[2 * 208];            // up to 27 taps x 8 chunks (512 ch) = 216 -> mtiles <= 208
};

__global__ void __launch_bounds__(TC_WG_THREADS, 1)
tc_wgrad_kernel(const __grid_constant__ TcWgradParams p, float* __restrict__ dw) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nb = p.BN / 64;                                    // B atoms per stage
  const int sub_bytes = (2 + nb) * WG_ATOM_BYTES;              // one 64-pixel tile: [A atom 0 | A atom 1 | B atoms]
  const int stage_bytes = p.pps * sub_bytes;
  const uint32_t bar_base = smem_base + p.stages * stage_bytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (p.stages + s); };
  const uint32_t tmem_full_bar = bar_base + 8u * (2 * p.stages);
  const uint32_t tmem_slot = bar_base + 8u * (2 * p.stages + 1);

  int t = blockIdx.x;
  const int split = t % p.splits; t /= p.splits;
  const int nt = t % p.ntiles_n; t /= p.ntiles_n;
  const int mt = t;
  const WgAtom a0 = p.atoms[2 * mt], a1 = p.atoms[2 * mt + 1];
  const int k0 = nt * p.BN;
  const int pt_begin = split * p.ptiles_per_split;
  const int pt_end = min(p.ptiles, pt_begin + p.ptiles_per_split);
  const int npt = pt_end - pt_begin;                           // 64-pixel tiles of this split
  const int iters = (npt + p.pps - 1) / p.pps;                 // pipeline stages (the last may be partial)
  const uint32_t tmem_cols = p.BN < 32 ? 32 : p.BN;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(full_bar(s), 2); mbar_init(empty_bar(s), 1); }   // full: A + B producers
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
    tma_prefetch_desc(&p.smap);
  }
  if (warp == 1) tmem_alloc(tmem_slot, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  pdl_grid_sync();
  if (iters <= 0) {   // (cannot happen with the host's split computation; keep the teardown uniform)
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, tmem_cols);
    return;
  }

  // pixel-tile coordinates of the first tile of this split (one division chain, outside the loops)
  int tiw, tih, tid_, tin;
  {
    int pt = pt_begin;
    tiw = pt % p.tw; pt /= p.tw;
    tih = pt % p.th; pt /= p.th;
    tid_ = pt % p.td; pt /= p.td;
    tin = pt;
  }
  if (warp == 0 || warp == 6) {
    // warp 0: the two large-tensor atoms (A operand); warp 6: the small-tensor atoms (B operand).
    // Whole-warp loops, one elected lane around the TMA issue (see tc_pixgemm_kernel).
    const bool isA = warp == 0;
    const bool has1 = a1.c0 >= 0;
    const void* m0 = &p.lmap[a0.view];
    const void* m1 = &p.lmap[a1.view];
    const uint32_t bytes = isA ? (uint32_t)((has1 ? 2 : 1) * WG_ATOM_BYTES) : (uint32_t)(nb * WG_ATOM_BYTES);
    int s = 0;
    uint32_t ph = 1;
    uint32_t dst = smem_base, fb = bar_base, eb = bar_base + 8u * p.stages;
    const int nst = p.stages, pps = p.pps;
    int iw = tiw, ih = tih, id = tid_, in = tin;
    int left = npt;
    for (int it = 0; it < iters; ++it) {
      const int nsub = left < pps ? left : pps;
      left -= nsub;
      mbar_wait(eb, ph);
      const bool issue = elect_one();
      if (issue) mbar_expect_tx(fb, bytes * (uint32_t)nsub);
      for (int u = 0; u < nsub; ++u) {
        const int w0 = iw * p.bw, h0 = ih * p.bh, d0 = id * p.bd, n0 = in * p.bn;
        const uint32_t sub = dst + u * sub_bytes;
        if (issue) {
          if (isA) {
            tma_load_5d(sub, m0, fb, a0.c0, w0 + a0.ow, h0 + a0.oh, d0 + a0.od, n0);
            if (has1) tma_load_5d(sub + WG_ATOM_BYTES, m1, fb, a1.c0, w0 + a1.ow, h0 + a1.oh, d0 + a1.od, n0);
          } else {
            for (int j = 0; j < nb; ++j) tma_load_5d(sub + (2 + j) * WG_ATOM_BYTES, &p.smap, fb, k0 + j * 64, w0, h0, d0, n0);
          }
        }
        if (++iw == p.tw) { iw = 0; if (++ih == p.th) { ih = 0; if (++id == p.td) { id = 0; ++in; } } }
      }
      __syncwarp();
      dst += stage_bytes; fb += 8u; eb += 8u;
      if (++s == nst) { s = 0; ph ^= 1u; dst = smem_base; fb = bar_base; eb = bar_base + 8u * nst; }
    }
  } else if (warp == 1) {
    // A: M = 128 = 2 atoms (LBO = atom stride), MN-major; B: N = BN = nb atoms, MN-major.
    const uint32_t idesc = make_idesc_bf16(TILE_M, p.BN, 1, 1);
    const uint64_t desc_hi = make_smem_desc(0, WG_ATOM_BYTES, 1024);
    int s = 0;
    uint32_t ph = 0;
    uint32_t a_addr = smem_base, fb = bar_base, eb = bar_base + 8u * p.stages;
    const int nst = p.stages, pps = p.pps;
    int left = npt;
    for (int it = 0; it < iters; ++it) {
      const int nsub = left < pps ? left : pps;
      left -= nsub;
      mbar_wait(fb, ph);
      tc_fence_after();
      if (elect_one()) {
        for (int u = 0; u < nsub; ++u) {
          const uint32_t sub = a_addr + u * sub_bytes;
          const uint64_t ad = desc_hi | (uint64_t)((sub & 0x3FFFFu) >> 4);
          const uint64_t bd = desc_hi | (uint64_t)(((sub + 2 * WG_ATOM_BYTES) & 0x3FFFFu) >> 4);
#pragma unroll
          for (int k = 0; k < WG_PIX / 16; ++k)   // 16 pixels (K) per MMA = two 8-row groups: +2048 B = +128 units per step
            umma_bf16(tmem_base, ad + 128u * k, bd + 128u * k, idesc, (it > 0 || u > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(eb);
        if (it == iters - 1) umma_commit(tmem_full_bar);
      }
      a_addr += stage_bytes; fb += 8u; eb += 8u;
      if (++s == nst) { s = 0; ph ^= 1u; a_addr = smem_base; fb = bar_base; eb = bar_base + 8u * nst; }
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;                 // 0..127: atom (row >> 6), channel (row & 63)
    if (lane == 0) mbar_wait(tmem_full_bar, 0);
    __syncwarp();
    tc_fence_after();
    // Accumulator -> shared memory (128B-swizzled boxes: [BN/32 column blocks][128 rows][128 B], the same staging as the
    // fp32 epilogue of tc_pixgemm) -> TMA reduce-store (cp.reduce.async.bulk.tensor .add): the += into the flat gradient
    // buffer is done by the L2 on whole 128-byte lines, one instruction per 64 x 32 box, instead of 64 scattered 16-byte
    // red.global per thread (which bounded this kernel: 3 splits x 13 MB of 16-byte atomics for d_h3).
    // All pipeline stages are free once tmem_full fired.
    const uint32_t row_off = (uint32_t)row * 128u;
    const uint32_t sw = (uint32_t)(row & 7);
    for (int c0 = 0; c0 < p.BN; c0 += 32) {
      uint32_t r[32];
      tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
      tmem_ld_wait();
      const uint32_t blk = smem_base + (uint32_t)(c0 >> 5) * (TILE_M * 128u) + row_off;
#pragma unroll
      for (int g = 0; g < 8; ++g)
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(blk + ((((uint32_t)g) ^ sw) << 4)), "r"(r[g * 4]), "r"(r[g * 4 + 1]),
                     "r"(r[g * 4 + 2]), "r"(r[g * 4 + 3]) : "memory");
    }
    fence_proxy_async();                                       // generic-proxy smem writes -> visible to the TMA engine
    asm volatile("bar.sync 1, 128;" ::: "memory");            // the four epilogue warps
    if (warp == 2 && lane == 0) {
      const int nblk = p.BN >> 5;
      for (int h = 0; h < 2; ++h) {
        const WgAtom a = h ? a1 : a0;
        if (a.c0 < 0) continue;                                // padding atom of an odd atom count
        const int row0 = (int)a.widx * p.C + (int)a.c0;        // first dw row of this atom
        for (int b = 0; b < nblk; ++b)
          tma_reduce_add_2d(&p.dwmap, smem_base + (uint32_t)b * (TILE_M * 128u) + (uint32_t)h * (64u * 128u), k0 + b * 32, row0);
      }
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // complete (not only read) before the CTA retires
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------------
// host: plans
// ------------------------------------------------------------------------------------------------
}  // namespace gg
// fused batch-norm backward reductions requested by the caller of a dgrad launch (see TcPixParams::bnb_*)
struct gg_bnbwd_args {
  const float *pre, *mean, *rstd, *gamma, *beta;
  double* sums;
  int act;
  float act_param;
  int groups;
};
namespace gg {
static void set_bnb(TcPixParams& p, const gg_bnbwd_args* b, int N, int* fused) {
  if (b == nullptr || b->pre == nullptr || b->sums == nullptr) return;
  if (b->groups < 1 || N % b->groups != 0 || (N / b->groups) % p.bn != 0 || ((uintptr_t)b->pre % 16) != 0) return;   // a tile must not straddle row groups
  p.bnb_pre = b->pre; p.bnb_mean = b->mean; p.bnb_rstd = b->rstd; p.bnb_gamma = b->gamma; p.bnb_beta = b->beta;
  p.bnb_sums = b->sums; p.bnb_act = b->act; p.bnb_act_param = b->act_param; p.stats_groups = b->groups;
  if (fused) *fused = 1;
}

// The pre-norm tensor of a fused batch-norm backward, seen through the geometry of output map k (same dims, same element
// strides, fp32): pmap[k].  `estr` are the ELEMENT strides of dims 1..4, `base_elems` the element offset of the class view.
static int encode_pre_like(TcPixParams& p, const gg_bnbwd_args* b, int k, const uint64_t* odims, const uint64_t* estr, uint64_t base_elems) {
  if (b == nullptr || b->pre == nullptr || b->sums == nullptr) return GG_OK;
  static const bool enabled = [] { const char* v = getenv("GG_BNB_TMA"); return !(v && v[0] == '0'); }();   // A/B switch
  if (!enabled) return GG_OK;
  const uint64_t pstr[4] = {estr[0] * 4, estr[1] * 4, estr[2] * 4, estr[3] * 4};
  const uint32_t pbox[5] = {32, (uint32_t)p.bw, (uint32_t)p.bh, (uint32_t)p.bd, (uint32_t)p.bn};
  int rc = encode_tmap(&p.pmap[k], GG_F32, (const char*)b->pre + base_elems * 4, 5, odims, pstr, pbox);
  if (rc) return rc;
  p.pre_tma = 1;
  return GG_OK;
}

static int pow2floor(int x) { int p = 1; while (p * 2 <= x) p *= 2; return p; }

static void pick_box(int Mw, int Mh, int Md, int pixels, int* bw, int* bh, int* bd, int* bn) {
  *bw = std::min(pow2floor(Mw), pixels);
  *bh = std::min(pow2floor(Mh), pixels / *bw);
  *bd = std::min(pow2floor(Md), pixels / (*bw * *bh));
  *bn = pixels / (*bw * *bh * *bd);
}

// need_large / need_small: which activation tensors are GEMM OPERANDS (must be bf16); an output may be fp32 or bf16
static int check_tc(const gg_conv_desc* d, const void* a, const void* b, bool need_large, bool need_small) {
  GG_REQUIRE((!need_large || d->large_dtype == GG_BF16) && (!need_small || d->small_dtype == GG_BF16), GG_ERR_UNSUPPORTED,
             "tensor-core path needs bf16 operand activations");
  GG_REQUIRE(d->C % 64 == 0 && d->K % 64 == 0, GG_ERR_UNSUPPORTED, "tensor-core path needs channel counts that are multiples of 64 (C=%d K=%d)", d->C, d->K);
  GG_REQUIRE(d->kd * d->kh * d->kw <= TC_MAX_TAPS && d->sd * d->sh * d->sw <= TC_MAX_VIEWS, GG_ERR_UNSUPPORTED, "tensor-core path: too many taps / stride classes");
  GG_REQUIRE(((uintptr_t)a % 16 == 0) && ((uintptr_t)b % 16 == 0), GG_ERR_INVALID, "tensor-core path needs 16-byte aligned tensors");
  return GG_OK;
}

// tensor maps of the stride-parity views of the large tensor: view index = (ad*sh + ah)*sw + aw
static int make_large_views(const gg_conv_desc* d, const void* large, const uint32_t* box, CUtensorMap* maps) {
  for (int ad = 0; ad < d->sd; ++ad)
    for (int ah = 0; ah < d->sh; ++ah)
      for (int aw = 0; aw < d->sw; ++aw) {
        const int64_t base_elems = (((int64_t)ad * d->H + ah) * d->W + aw) * d->C;
        const uint64_t dims[5] = {(uint64_t)d->C, (uint64_t)((d->W - aw + d->sw - 1) / d->sw), (uint64_t)((d->H - ah + d->sh - 1) / d->sh),
                                  (uint64_t)((d->D - ad + d->sd - 1) / d->sd), (uint64_t)d->N};
        const uint64_t str[4] = {(uint64_t)d->sw * d->C * 2, (uint64_t)d->sh * d->W * d->C * 2, (uint64_t)d->sd * d->H * d->W * d->C * 2,
                                 (uint64_t)d->D * d->H * d->W * d->C * 2};
        if (dims[1] == 0 || dims[2] == 0 || dims[3] == 0) {   // empty view (grid smaller than the stride): never referenced
          memset(&maps[(ad * d->sh + ah) * d->sw + aw], 0, sizeof(CUtensorMap));
          continue;
        }
        int rc = encode_tmap_bf16(&maps[(ad * d->sh + ah) * d->sw + aw], (const bf16*)large + base_elems, 5, dims, str, box);
        if (rc) return rc;
      }
  return GG_OK;
}

static int make_small_map(const gg_conv_desc* d, const void* small, const uint32_t* box, CUtensorMap* map) {
  const uint64_t dims[5] = {(uint64_t)d->K, (uint64_t)d->Wo, (uint64_t)d->Ho, (uint64_t)d->Do, (uint64_t)d->N};
  const uint64_t str[4] = {(uint64_t)d->K * 2, (uint64_t)d->Wo * d->K * 2, (uint64_t)d->Ho * d->Wo * d->K * 2,
                           (uint64_t)d->Do * d->Ho * d->Wo * d->K * 2};
  return encode_tmap_bf16(map, small, 5, dims, str, box);
}

static inline int floordiv(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }
static inline int posmod(int a, int b) { int m = a % b; return m < 0 ? m + b : m; }

// ---- tuning / measurement hooks (tools/tc_sweep.py, bench.py): not part of the product path ----
static int g_repeat = 1;
static unsigned long long* g_prof = nullptr;      // device buffer [512][8] set by tc_set_prof
void tc_set_prof(void* buf) { g_prof = (unsigned long long*)buf; }                 // launches per call (amortises host planning when timing a kernel)
static int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return (v && *v) ? atoi(v) : dflt;
}
void tc_set_repeat(int n) { g_repeat = n < 1 ? 1 : n; }

// ---- split-K workspace (gg_set_workspace): [64 KB of arrive / depart counters, zero at rest | scratch ring for the fp32 partials] ----
struct TcWorkspace {
  char* base = nullptr;           // scratch ring
  size_t bytes = 0, head = 0;
  unsigned int* cnt = nullptr;    // counter ring
  size_t cnt_n = 0, cnt_head = 0;
};
static TcWorkspace g_ws;
static std::mutex g_ws_mu;
constexpr size_t WS_COUNTER_BYTES = 64 * 1024;
size_t tc_workspace_bytes() { return WS_COUNTER_BYTES + (size_t)96 * 1024 * 1024; }
int tc_set_workspace(void* ptr, size_t bytes) {
  std::lock_guard<std::mutex> lock(g_ws_mu);
  if (ptr == nullptr || bytes <= WS_COUNTER_BYTES + (1 << 20)) { g_ws = TcWorkspace(); return GG_OK; }
  GG_REQUIRE(((uintptr_t)ptr % 1024) == 0, GG_ERR_INVALID, "gg_set_workspace: the buffer must be 1024-byte aligned");
  g_ws.cnt = (unsigned int*)ptr; g_ws.cnt_n = WS_COUNTER_BYTES / 4; g_ws.cnt_head = 0;
  g_ws.base = (char*)ptr + WS_COUNTER_BYTES; g_ws.bytes = bytes - WS_COUNTER_BYTES; g_ws.head = 0;
  return GG_OK;
}

// Output-channel tile and split-K factor.  An M=128 tcgen05.mma from shared memory costs max(~76, N/2) cycles, so a
// wider N is cheaper per FLOP -- but at batch 64 the wide tiles leave most SMs idle.  Without split: the widest tile that
// still yields ~1 wave of CTAs.  The 4x4 / 8x8 layers (g_h1, d_h3: 8-32 M tiles, 100-200 K chunks per tile) then end up
// with N = 64 tiles at 42 % of the MMA rate; for them the K loop of a WIDE tile is split over S CTAs instead and the S
// fp32 partials are combined by a reduce-scatter: every rank finishes (bias, activation, statistics, store) BN / S columns.
// Two transports (tc_pixgemm_kernel), both parity-tested (tests/test_gpu_tc.py::test_tc_split_k), both OPT-IN:
//   * GG_TC_SPLITK_MODE=cluster: thread-block cluster + distributed shared memory, S in {2, 4}.  MEASURED (profiles/r2p_*):
//     the main loop halves as modelled (d_h3.down 34.3 k -> 18.2 k cycles to accumulator-ready) but the exchange costs
//     ~20 k cycles per CTA -- DSMEM moves ~17-21 B/clk per SM, a 128 x 256 fp32 partial is 128 KB (g_h1.up 17.5 -> 23.3 us).
//   * default mode (needs gg_set_workspace): through L2 with TMA stores / loads of 16 KB boxes and a per-tile arrive counter;
//     S in {2, 4, 8}; deterministic (fixed summation order), no cluster scheduling constraints.  MEASURED (profiles/
//     r4c_splitk_l2_sweep.md, per-CTA clock64): a 128 KB partial costs staging 2.1 k + TMA store 4.3 k (~30 B/clk per SM into
//     L2) + counter / skew 2.3 k + TMA load 3.2 k + 2 k of extra epilogue = ~14 k cycles on top of the normal epilogue, a 64 KB
//     partial ~9 k.  Alone, the split launches are 0-25 % faster (g_h1.down with statistics 23.7 -> 17.6 us at N = 128, S = 4;
//     g_h1.up / d_h3 within +-3 %); INSIDE the train step they lose (1.525 -> 1.569 ms): a one-wave launch that spins on its
//     tile mates shares the SMs with the filter-gradient stream, and each launch streams 8-16 MB of scratch through L2.
// GG_TC_SPLITK=auto lets the cycle model below choose, =2 / 4 / 8 forces S; unset = no split.  GG_TC_BN forces the tile.
static void pick_tile(int Nout, int64_t mtiles, int min_chunks, int max_chunks, bool out_bf16, bool allow_split, int* bn_out, int* split_out, int* mode_out) {
  int bn = Nout % 256 == 0 ? 256 : (Nout % 128 == 0 ? 128 : 64);
  while (bn > 64 && mtiles * (Nout / bn) < 120) bn /= 2;
  const int forced_bn = env_int("GG_TC_BN", 0);
  if (forced_bn > 0 && Nout % forced_bn == 0) bn = forced_bn;
  int S = 1;
  const char* sk = getenv("GG_TC_SPLITK");
  const char* skm = getenv("GG_TC_SPLITK_MODE");
  const bool cluster = skm && skm[0] == 'c';
  const bool auto_split = sk && sk[0] == 'a';
  const int forced_s = auto_split ? 0 : ((sk && *sk) ? std::max(1, atoi(sk)) : 1);
  const bool have_ws = g_ws.base != nullptr;
  if (!cluster && !have_ws) allow_split = false;
  auto cycles = [&](int b, int s) {        // one CTA: main loop of the heaviest class + fixed cost + the reduce-scatter
    const int mma = std::max(76, b / 2);
    const int xchg = cluster ? 1500 + 75 * b : 3000 + 14 * b;             // cluster: ~20 k cycles at b = 256; L2: store + counter + load of b * 512 bytes
    return (int64_t)ceil_div(max_chunks, s) * 4 * mma + 6000 + (s > 1 ? xchg : 0);
  };
  auto split_ok = [&](int b, int s) {
    if (Nout % b != 0 || b % (32 * s) != 0 || b / s < (out_bf16 ? 64 : 32)) return false;
    if (mtiles * (Nout / b) * s > 148) return false;                                   // one wave, one CTA per SM: the ranks of a tile are co-resident
    if (!cluster && (size_t)mtiles * (Nout / b) * s * b * 512 > g_ws.bytes) return false;
    if (cluster && s > 4) return false;
    return (int64_t)(s - 1) * ceil_div(min_chunks, s) < min_chunks && min_chunks / s >= 2;   // every rank has work
  };
  if (allow_split && forced_s != 1) {
    const int64_t tiles0 = mtiles * (Nout / bn);
    int64_t best = cycles(bn, 1) * ceil_div64(tiles0, 148);
    if (forced_s > 1) best = INT64_MAX;
    for (int b : {256, 128})
      for (int s2 : {2, 4, 8}) {
        if (forced_s > 1 && s2 != forced_s) continue;
        if (forced_bn > 0 && b != forced_bn) continue;
        if (!split_ok(b, s2)) continue;
        const int64_t c = cycles(b, s2);
        if (forced_s > 1 ? c < best : c * 100 < best * 85) { best = (forced_s > 1) ? c : c * 100 / 85; bn = b; S = s2; }
      }
  }
  *bn_out = bn; *split_out = S; *mode_out = (S > 1 && !cluster) ? 1 : 0;
}

static int launch_pix(TcPixParams& p, int total_tiles, const float* bias, void* out, cudaStream_t st) {
  const int chunk_bytes = A_STAGE_BYTES + p.BN * 128;
  // Issue-side measurements (tools/mma_bench.cu): a pipeline stage costs ~450 cycles of fixed issue/commit latency on
  // top of its MMAs, so a stage should carry >= 8 MMAs (two 64-wide K chunks): 128 x 128 x 16 then runs at ~80
  // cycles per MMA instead of ~137.  When the grid exceeds one wave, two co-resident CTAs per SM (<= ~100 KB each)
  // hide each other's prologue / epilogue and are preferred over deeper stages.
  const int budget = (total_tiles * p.splitk > 148 && p.splitk == 1) ? 100 * 1024 : 200 * 1024;
  p.cps = (budget / (2 * chunk_bytes)) >= 2 ? 2 : 1;
  p.cps = std::max(1, std::min(2, env_int("GG_TC_CPS", p.cps)));
  const int stage_bytes = p.cps * chunk_bytes;
  p.stages = std::max(2, std::min(8, budget / stage_bytes));
  // the epilogue stages the tile in the pipeline buffers; a split-K tile parks its fp32 partial there first (+ bf16 staging behind it)
  // split-K: S x (BN / S) columns of fp32 partials (= BN * 512 bytes) + the staging of this rank's slice behind them
  int out_bytes = p.splitk > 1 ? p.BN * 512 + TILE_M * (p.BN / p.splitk) * (p.out_bf16 ? 2 : 4) : TILE_M * p.BN * (p.out_bf16 ? 2 : 4);
  if (p.bnb_pre == nullptr || p.splitk > 1) p.pre_tma = 0;       // (split-K slices keep the direct loads)
  p.epi2 = (p.splitk == 1 && env_int("GG_TC_EPI2", 1) != 0) ? 1 : 0;
  if (p.pre_tma) {                                               // two 16 KB buffers for the pre-norm blocks, behind the output staging
    p.pre_off = (uint32_t)((out_bytes + 1023) & ~1023);
    out_bytes = (int)p.pre_off + 2 * TILE_M * 128;
  }
  while (p.stages * stage_bytes < out_bytes) ++p.stages;
  p.stages = std::max(1, std::min(p.stages, env_int("GG_TC_STAGES", p.stages)));
  GG_REQUIRE(p.stages * stage_bytes >= out_bytes, GG_ERR_INVALID, "tc_pixgemm: GG_TC_STAGES too small for the output tile");
  GG_REQUIRE((size_t)p.stages * stage_bytes + 2048 + 48 * 256 <= 227 * 1024, GG_ERR_INVALID, "tc_pixgemm: pipeline does not fit in shared memory");
  p.prof = g_prof;
  if (p.splitk > 1 && p.sk_mode == 1) {
    // scratch of this launch: the next chunk of the ring (launches that may run concurrently -- other streams, parallel
    // branches of a captured graph -- must not share one; stream order protects a chunk that comes round again)
    std::lock_guard<std::mutex> lock(g_ws_mu);
    const size_t need = (size_t)total_tiles * p.splitk * p.BN * 512;
    GG_REQUIRE(g_ws.base != nullptr && need <= g_ws.bytes && (size_t)2 * total_tiles <= g_ws.cnt_n, GG_ERR_INVALID, "tc_pixgemm: split-K workspace missing or too small");
    if (g_ws.head + need > g_ws.bytes) g_ws.head = 0;
    if (g_ws.cnt_head + 2 * (size_t)total_tiles > g_ws.cnt_n) g_ws.cnt_head = 0;
    const uint64_t kdims[2] = {32, (uint64_t)total_tiles * p.splitk * (p.BN / 32) * TILE_M};
    const uint64_t kstr[1] = {128};
    const uint32_t kbox[2] = {32, TILE_M};
    int rc = encode_tmap(&p.skmap, GG_F32, g_ws.base + g_ws.head, 2, kdims, kstr, kbox);
    if (rc) return rc;
    p.sk_cnt = g_ws.cnt + g_ws.cnt_head;
    g_ws.head += need;
    g_ws.cnt_head += 2 * (size_t)total_tiles;
  }
  const size_t smem = (size_t)p.stages * stage_bytes + 1024 + 8 * (2 * p.stages + 4) + 32 + (p.bnb_pre ? 48 * p.BN : 0);
  static std::once_flag once;
  std::call_once(once, [] { cudaFuncSetAttribute(tc_pixgemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); });
  int rc = GG_OK;
  for (int r = 0; r < g_repeat && rc == GG_OK; ++r) {
    Launch(total_tiles * p.splitk, TC_THREADS, smem, st, p.sk_mode == 1 ? 1 : p.splitk)(tc_pixgemm_kernel, p, bias, out);
    rc = check_launch("tc_pixgemm");
  }
  return rc;
}

// stats != nullptr: also accumulate per-channel (sum, sum of squares) of the fp32 output into stats[groups][2][channels].
// *fused is set to 1 when the kernel did it (otherwise the caller runs a separate statistics pass).
int tc_conv_down(const gg_conv_desc* d, const void* large, const void* w_bf, const float* bias, void* small, cudaStream_t st,
                 double* stats = nullptr, int groups = 1, int* fused = nullptr, const gg_bnbwd_args* bnb = nullptr) {
  int rc = check_tc(d, large, w_bf, true, false);
  if (rc) return rc;
  TcPixParams p;
  memset(&p, 0, sizeof(p));
  pick_box(d->Wo, d->Ho, d->Do, TILE_M, &p.bw, &p.bh, &p.bd, &p.bn);
  const uint32_t abox[5] = {64, (uint32_t)p.bw, (uint32_t)p.bh, (uint32_t)p.bd, (uint32_t)p.bn};
  rc = make_large_views(d, large, abox, p.amap);
  if (rc) return rc;
  TcClass& c = p.cls[0];
  c.Md = d->Do; c.Mh = d->Ho; c.Mw = d->Wo;
  c.tw = ceil_div(d->Wo, p.bw); c.th = ceil_div(d->Ho, p.bh); c.td = ceil_div(d->Do, p.bd); c.tn = ceil_div(d->N, p.bn);
  const int64_t mtiles = (int64_t)c.tw * c.th * c.td * c.tn;
  const int taps = d->kd * d->kh * d->kw;
  pick_tile(d->K, mtiles, taps * (d->C / KCHUNK), taps * (d->C / KCHUNK), d->small_dtype == GG_BF16, true, &p.BN, &p.splitk, &p.sk_mode);
  {  // B operand straight from the bf16 filter copy w[tap][c][k] (no transposed copy): the reduction index c is the ROW
     // axis, so the tile is MN-major -- BN/64 atoms of [64 c rows][64 k = 128 B], one 4-D box per K chunk
    const uint64_t bdims[4] = {64, (uint64_t)d->C, (uint64_t)(d->K / 64), (uint64_t)taps};
    const uint64_t bstr[3] = {(uint64_t)d->K * 2, 128, (uint64_t)d->C * d->K * 2};
    const uint32_t bbox[4] = {64, 64, (uint32_t)(p.BN / 64), 1};
    rc = encode_tmap_bf16(&p.bmap, w_bf, 4, bdims, bstr, bbox);
    if (rc) return rc;
    p.bmode = 1;
  }
  {  // output map: the small tensor, dense
    const uint64_t esz = d->small_dtype == GG_BF16 ? 2 : 4;
    const uint64_t odims[5] = {(uint64_t)d->K, (uint64_t)d->Wo, (uint64_t)d->Ho, (uint64_t)d->Do, (uint64_t)d->N};
    const uint64_t ostr[4] = {(uint64_t)d->K * esz, (uint64_t)d->Wo * d->K * esz, (uint64_t)d->Ho * d->Wo * d->K * esz,
                              (uint64_t)d->Do * d->Ho * d->Wo * d->K * esz};
    const uint32_t obox[5] = {(uint32_t)(128 / esz), (uint32_t)p.bw, (uint32_t)p.bh, (uint32_t)p.bd, (uint32_t)p.bn};
    rc = encode_tmap(&p.omap[0], d->small_dtype, small, 5, odims, ostr, obox);
    if (rc) return rc;
    const uint64_t estr[4] = {(uint64_t)d->K, (uint64_t)d->Wo * d->K, (uint64_t)d->Ho * d->Wo * d->K, (uint64_t)d->Do * d->Ho * d->Wo * d->K};
    rc = encode_pre_like(p, bnb, 0, odims, estr, 0);
    if (rc) return rc;
  }
  int t = 0;
  for (int a = 0; a < d->kd; ++a)
    for (int b = 0; b < d->kh; ++b)
      for (int e = 0; e < d->kw; ++e, ++t) {
        const int fd = a - d->pd, fh = b - d->ph, fw = e - d->pw;
        TcTap tp;
        tp.view = (int8_t)((posmod(fd, d->sd) * d->sh + posmod(fh, d->sh)) * d->sw + posmod(fw, d->sw));
        tp.od = (int8_t)floordiv(fd, d->sd); tp.oh = (int8_t)floordiv(fh, d->sh); tp.ow = (int8_t)floordiv(fw, d->sw);
        tp.widx = (int16_t)t; tp.pad_ = 0;
        p.taps[t] = tp;
      }
  c.tap_begin = 0; c.tap_end = taps; c.tile_begin = 0;
  p.nclasses = 1; p.ntiles_n = d->K / p.BN;
  p.R = d->C; p.Nout = d->K; p.Mn = d->N;
  p.OD = d->Do; p.OH = d->Ho; p.OW = d->Wo; p.osd = p.osh = p.osw = 1;
  p.act = d->act; p.act_param = d->act_param; p.out_bf16 = (d->small_dtype == GG_BF16);
  if (stats != nullptr && d->small_dtype == GG_F32 && groups >= 1 && d->N % groups == 0 && (d->N / groups) % p.bn == 0) {
    p.stats = stats; p.stats_groups = groups;
    if (fused) *fused = 1;
  }
  set_bnb(p, bnb, d->N, fused);
  return launch_pix(p, (int)(mtiles * p.ntiles_n), bias, small, st);
}

int tc_conv_up(const gg_conv_desc* d, const void* small, const void* w_ck, const float* bias, void* large, cudaStream_t st,
               double* stats = nullptr, int groups = 1, int* fused = nullptr, const gg_bnbwd_args* bnb = nullptr) {
  int rc = check_tc(d, small, w_ck, false, true);
  if (rc) return rc;
  TcPixParams p;
  memset(&p, 0, sizeof(p));
  const int Mw = ceil_div(d->W, d->sw), Mh = ceil_div(d->H, d->sh), Md = ceil_div(d->D, d->sd);
  pick_box(Mw, Mh, Md, TILE_M, &p.bw, &p.bh, &p.bd, &p.bn);
  const uint32_t abox[5] = {64, (uint32_t)p.bw, (uint32_t)p.bh, (uint32_t)p.bd, (uint32_t)p.bn};
  rc = make_small_map(d, small, abox, &p.amap[0]);
  if (rc) return rc;
  int ncls = 0, ntap = 0;
  int64_t tiles = 0, mt_max = 0;
  for (int ad = 0; ad < d->sd; ++ad)
    for (int ah = 0; ah < d->sh; ++ah)
      for (int aw = 0; aw < d->sw; ++aw) {
        TcClass c;
        memset(&c, 0, sizeof(c));
        c.Md = (d->D - ad + d->sd - 1) / d->sd; c.Mh = (d->H - ah + d->sh - 1) / d->sh; c.Mw = (d->W - aw + d->sw - 1) / d->sw;
        if (c.Md <= 0 || c.Mh <= 0 || c.Mw <= 0) continue;
        c.od0 = ad; c.oh0 = ah; c.ow0 = aw;
        c.tap_begin = ntap;
        for (int a = 0; a < d->kd; ++a) {
          if (posmod(ad + d->pd - a, d->sd) != 0) continue;
          for (int b = 0; b < d->kh; ++b) {
            if (posmod(ah + d->ph - b, d->sh) != 0) continue;
            for (int e = 0; e < d->kw; ++e) {
              if (posmod(aw + d->pw - e, d->sw) != 0) continue;
              TcTap tp;
              tp.view = 0;
              tp.od = (int8_t)floordiv(ad + d->pd - a, d->sd); tp.oh = (int8_t)floordiv(ah + d->ph - b, d->sh);
              tp.ow = (int8_t)floordiv(aw + d->pw - e, d->sw);
              tp.widx = (int16_t)((a * d->kh + b) * d->kw + e); tp.pad_ = 0;
              p.taps[ntap++] = tp;
            }
          }
        }
        c.tap_end = ntap;
        c.tw = ceil_div(c.Mw, p.bw); c.th = ceil_div(c.Mh, p.bh); c.td = ceil_div(c.Md, p.bd); c.tn = ceil_div(d->N, p.bn);
        const int64_t mt = (int64_t)c.tw * c.th * c.td * c.tn;
        mt_max = std::max(mt_max, mt);
        c.tile_begin = (int)tiles;   // scaled by ntiles_n below
        tiles += mt;
        {  // output map of this class: the stride-s parity view of the large tensor that the class writes
          const uint64_t esz = d->large_dtype == GG_BF16 ? 2 : 4;
          const uint64_t base_elems = (((uint64_t)ad * d->H + ah) * d->W + aw) * d->C;
          const uint64_t odims[5] = {(uint64_t)d->C, (uint64_t)c.Mw, (uint64_t)c.Mh, (uint64_t)c.Md, (uint64_t)d->N};
          const uint64_t ostr[4] = {(uint64_t)d->sw * d->C * esz, (uint64_t)d->sh * d->W * d->C * esz,
                                    (uint64_t)d->sd * d->H * d->W * d->C * esz, (uint64_t)d->D * d->H * d->W * d->C * esz};
          const uint32_t obox[5] = {(uint32_t)(128 / esz), (uint32_t)p.bw, (uint32_t)p.bh, (uint32_t)p.bd, (uint32_t)p.bn};
          rc = encode_tmap(&p.omap[ncls], d->large_dtype, (const char*)large + base_elems * esz, 5, odims, ostr, obox);
          if (rc) return rc;
          const uint64_t estr[4] = {(uint64_t)d->sw * d->C, (uint64_t)d->sh * d->W * d->C, (uint64_t)d->sd * d->H * d->W * d->C, (uint64_t)d->D * d->H * d->W * d->C};
          rc = encode_pre_like(p, bnb, ncls, odims, estr, base_elems);
          if (rc) return rc;
        }
        p.cls[ncls++] = c;
      }
  {
    int min_taps = TC_MAX_TAPS, max_taps = 0;
    for (int i = 0; i < ncls; ++i) {
      min_taps = std::min(min_taps, p.cls[i].tap_end - p.cls[i].tap_begin);
      max_taps = std::max(max_taps, p.cls[i].tap_end - p.cls[i].tap_begin);
    }
    pick_tile(d->C, tiles, min_taps * (d->K / KCHUNK), max_taps * (d->K / KCHUNK), d->large_dtype == GG_BF16, true, &p.BN, &p.splitk, &p.sk_mode);
  }
  p.ntiles_n = d->C / p.BN;
  for (int i = 0; i < ncls; ++i) p.cls[i].tile_begin *= p.ntiles_n;
  // a class without taps (stride > kernel) still has to write bias/zeros: the kernel handles iters == 0? no -> reject
  for (int i = 0; i < ncls; ++i)
    GG_REQUIRE(p.cls[i].tap_end > p.cls[i].tap_begin, GG_ERR_UNSUPPORTED, "tensor-core conv_up: an output parity class has no taps");
  const int taps = d->kd * d->kh * d->kw;
  const uint64_t bdims[3] = {(uint64_t)d->K, (uint64_t)d->C, (uint64_t)taps};
  const uint64_t bstr[2] = {(uint64_t)d->K * 2, (uint64_t)d->C * d->K * 2};
  const uint32_t bbox[3] = {64, (uint32_t)p.BN, 1};
  rc = encode_tmap_bf16(&p.bmap, w_ck, 3, bdims, bstr, bbox);
  if (rc) return rc;
  p.nclasses = ncls;
  p.R = d->K; p.Nout = d->C; p.Mn = d->N;
  p.OD = d->D; p.OH = d->H; p.OW = d->W; p.osd = d->sd; p.osh = d->sh; p.osw = d->sw;
  p.act = d->act; p.act_param = d->act_param; p.out_bf16 = (d->large_dtype == GG_BF16);
  if (stats != nullptr && d->large_dtype == GG_F32 && groups >= 1 && d->N % groups == 0 && (d->N / groups) % p.bn == 0) {
    p.stats = stats; p.stats_groups = groups;
    if (fused) *fused = 1;
  }
  set_bnb(p, bnb, d->N, fused);
  return launch_pix(p, (int)(tiles * p.ntiles_n), bias, large, st);
}

// ------------------------------------------------------------------------------------------------
// conv_up with the output parity classes concatenated along N  (64-channel outputs: d_h1 dgrad, g_h3 forward)
// ------------------------------------------------------------------------------------------------
// With N = 64 an M=128 tcgen05.mma runs at ~42 % of the tensor peak (its cost is max(~78, N/2) cycles), and the four
// parity classes of a 5x5 stride-2 deconvolution read the SAME nine shifted input boxes (shift (dp, dq) in {-1,0,1}^2)
// with different filter taps.  So: one tile = 128 small-grid positions x (4 classes x 64 channels) = N 256, K loop over
// the 9 shifts; the packed filter wcat[shift][class*C + c][k] holds w[tap(class, shift)][c][k], or zero where the class
// does not use the shift (25 of 36 (class, shift) pairs are real: 69 % useful MACs at the full N=256 rate).  A traffic
// drops from 25 to 9 boxes per K chunk and the launch is one wave of M/128 CTAs instead of four classes of them.
// No packed filter exists: the B tile of a shift is assembled by the TMA producer from four boxes of the plain bf16 copy
// (TcPixParams::cat_tap), a missing (class, shift) pair being a box at an out-of-range tap index, which TMA zero-fills.
constexpr int UPCAT_MAX_SHIFTS = 27;
struct UpcatPlan {
  int ncls, nshift;
  int cls_a[8][3];                       // class -> (ad, ah, aw)
  int shift[UPCAT_MAX_SHIFTS][3];        // shift -> (od, oh, ow)
  int tap[UPCAT_MAX_SHIFTS][8];          // (shift, class) -> filter tap index or -1
};

bool tc_upcat_ok(const gg_conv_desc* d) {
  static const bool enabled = [] { const char* v = getenv("GG_UPCAT"); return !(v && v[0] == '0'); }();   // A/B switch
  const int ncls = d->sd * d->sh * d->sw;
  return enabled && (d->flags & GG_CONV_TENSOR_CORE) && ncls > 1 && ncls <= 8 && ncls * d->C == 256 && d->K % 64 == 0 && d->D % d->sd == 0 &&
         d->H % d->sh == 0 && d->W % d->sw == 0 && d->small_dtype == GG_BF16;
}

static int upcat_plan(const gg_conv_desc* d, UpcatPlan* P) {
  memset(P, 0, sizeof(*P));
  for (int s = 0; s < UPCAT_MAX_SHIFTS; ++s) for (int c = 0; c < 8; ++c) P->tap[s][c] = -1;
  int ncls = 0;
  for (int ad = 0; ad < d->sd; ++ad)
    for (int ah = 0; ah < d->sh; ++ah)
      for (int aw = 0; aw < d->sw; ++aw) {
        P->cls_a[ncls][0] = ad; P->cls_a[ncls][1] = ah; P->cls_a[ncls][2] = aw;
        for (int a = 0; a < d->kd; ++a) {
          if (posmod(ad + d->pd - a, d->sd) != 0) continue;
          for (int b = 0; b < d->kh; ++b) {
            if (posmod(ah + d->ph - b, d->sh) != 0) continue;
            for (int e = 0; e < d->kw; ++e) {
              if (posmod(aw + d->pw - e, d->sw) != 0) continue;
              const int od = floordiv(ad + d->pd - a, d->sd), oh = floordiv(ah + d->ph - b, d->sh), ow = floordiv(aw + d->pw - e, d->sw);
              int s = 0;
              for (; s < P->nshift; ++s) if (P->shift[s][0] == od && P->shift[s][1] == oh && P->shift[s][2] == ow) break;
              if (s == P->nshift) {
                GG_REQUIRE(P->nshift < UPCAT_MAX_SHIFTS, GG_ERR_UNSUPPORTED, "upcat: too many shifts");
                P->shift[s][0] = od; P->shift[s][1] = oh; P->shift[s][2] = ow; ++P->nshift;
              }
              P->tap[s][ncls] = (a * d->kh + b) * d->kw + e;
            }
          }
        }
        ++ncls;
      }
  P->ncls = ncls;
  return GG_OK;
}

int tc_conv_up_cat(const gg_conv_desc* d, const void* small, const void* w_bf, const float* bias, void* large, cudaStream_t st,
                   double* stats = nullptr, int groups = 1, int* fused = nullptr, const gg_bnbwd_args* bnb = nullptr) {
  int rc = check_tc(d, small, w_bf, false, true);
  if (rc) return rc;
  GG_REQUIRE(tc_upcat_ok(d), GG_ERR_UNSUPPORTED, "conv_up (concatenated classes): shape not eligible");
  UpcatPlan P;
  rc = upcat_plan(d, &P);
  if (rc) return rc;
  GG_REQUIRE(P.nshift <= TC_MAX_TAPS && P.ncls <= TC_MAX_CLASSES, GG_ERR_UNSUPPORTED, "conv_up (concatenated classes): too many shifts");
  TcPixParams p;
  memset(&p, 0, sizeof(p));
  const int Mw = d->W / d->sw, Mh = d->H / d->sh, Md = d->D / d->sd;      // positions = the small grid's footprint of one class
  pick_box(Mw, Mh, Md, TILE_M, &p.bw, &p.bh, &p.bd, &p.bn);
  const uint32_t abox[5] = {64, (uint32_t)p.bw, (uint32_t)p.bh, (uint32_t)p.bd, (uint32_t)p.bn};
  rc = make_small_map(d, small, abox, &p.amap[0]);
  if (rc) return rc;
  TcClass& c = p.cls[0];
  c.Md = Md; c.Mh = Mh; c.Mw = Mw;
  c.tw = ceil_div(Mw, p.bw); c.th = ceil_div(Mh, p.bh); c.td = ceil_div(Md, p.bd); c.tn = ceil_div(d->N, p.bn);
  const int64_t mtiles = (int64_t)c.tw * c.th * c.td * c.tn;
  for (int s = 0; s < P.nshift; ++s) {
    TcTap tp;
    tp.view = 0; tp.od = (int8_t)P.shift[s][0]; tp.oh = (int8_t)P.shift[s][1]; tp.ow = (int8_t)P.shift[s][2];
    tp.widx = (int16_t)s; tp.pad_ = 0;
    p.taps[s] = tp;
  }
  c.tap_begin = 0; c.tap_end = P.nshift; c.tile_begin = 0;
  const uint64_t esz = d->large_dtype == GG_BF16 ? 2 : 4;
  for (int k = 0; k < P.ncls; ++k) {   // output map of class k: the stride-s parity view of the large tensor it writes
    const int ad = P.cls_a[k][0], ah = P.cls_a[k][1], aw = P.cls_a[k][2];
    const uint64_t base_elems = (((uint64_t)ad * d->H + ah) * d->W + aw) * d->C;
    const uint64_t odims[5] = {(uint64_t)d->C, (uint64_t)Mw, (uint64_t)Mh, (uint64_t)Md, (uint64_t)d->N};
    const uint64_t ostr[4] = {(uint64_t)d->sw * d->C * esz, (uint64_t)d->sh * d->W * d->C * esz, (uint64_t)d->sd * d->H * d->W * d->C * esz,
                              (uint64_t)d->D * d->H * d->W * d->C * esz};
    const uint32_t obox[5] = {(uint32_t)(128 / esz), (uint32_t)p.bw, (uint32_t)p.bh, (uint32_t)p.bd, (uint32_t)p.bn};
    rc = encode_tmap(&p.omap[k], d->large_dtype, (const char*)large + base_elems * esz, 5, odims, ostr, obox);
    if (rc) return rc;
    const uint64_t estr[4] = {(uint64_t)d->sw * d->C, (uint64_t)d->sh * d->W * d->C, (uint64_t)d->sd * d->H * d->W * d->C, (uint64_t)d->D * d->H * d->W * d->C};
    rc = encode_pre_like(p, bnb, k, odims, estr, base_elems);
    if (rc) return rc;
  }
  p.BN = P.ncls * d->C;                 // 256
  p.splitk = 1;
  p.cat_C = d->C;
  p.ntiles_n = 1;
  // B tile of shift s: class (ad, ah, aw) reads tap (pd + ad - sd*od, ph + ah - sh*oh, pw + aw - sw*ow) -- affine in the class,
  // so the tile is ONE 5-D box (64 k, C rows, sw, sh, sd) of the plain bf16 filter copy w[kd][kh][kw][C][K] starting at tap
  // (pd - sd*od, ph - sh*oh, pw - sw*ow); a class that does not use the shift falls outside [0, k) and is zero-filled by TMA.
  // Box order (aw fastest, then ah, ad) = class order of upcat_plan = row blocks of the N = ncls * C tile.
  const uint64_t bdims[5] = {(uint64_t)d->K, (uint64_t)d->C, (uint64_t)d->kw, (uint64_t)d->kh, (uint64_t)d->kd};
  const uint64_t bstr[4] = {(uint64_t)d->K * 2, (uint64_t)d->C * d->K * 2, (uint64_t)d->kw * d->C * d->K * 2, (uint64_t)d->kh * d->kw * d->C * d->K * 2};
  const uint32_t bbox[5] = {64, (uint32_t)d->C, (uint32_t)d->sw, (uint32_t)d->sh, (uint32_t)d->sd};
  rc = encode_tmap_bf16(&p.bmap, w_bf, 5, bdims, bstr, bbox);
  if (rc) return rc;
  p.bmode = 2; p.ncat = P.ncls;
  for (int s = 0; s < P.nshift; ++s) {
    TcCatOrigin o;
    o.a0 = (int8_t)(d->pd - d->sd * P.shift[s][0]); o.b0 = (int8_t)(d->ph - d->sh * P.shift[s][1]); o.e0 = (int8_t)(d->pw - d->sw * P.shift[s][2]);
    o.pad_ = 0;
    p.cat_origin[s] = o;
    for (int k = 0; k < P.ncls; ++k) {      // the plan's own table must agree with the affine form
      const int a = o.a0 + P.cls_a[k][0], b = o.b0 + P.cls_a[k][1], e = o.e0 + P.cls_a[k][2];
      const bool in = a >= 0 && a < d->kd && b >= 0 && b < d->kh && e >= 0 && e < d->kw;
      GG_REQUIRE((in ? (a * d->kh + b) * d->kw + e : -1) == P.tap[s][k], GG_ERR_INVALID, "conv_up (concatenated classes): tap table mismatch");
    }
  }
  p.nclasses = 1;
  p.R = d->K; p.Nout = p.BN; p.Mn = d->N;
  p.OD = d->D; p.OH = d->H; p.OW = d->W; p.osd = d->sd; p.osh = d->sh; p.osw = d->sw;
  p.act = d->act; p.act_param = d->act_param; p.out_bf16 = (d->large_dtype == GG_BF16);
  if (stats != nullptr && d->large_dtype == GG_F32 && groups >= 1 && d->N % groups == 0 && (d->N / groups) % p.bn == 0) {
    p.stats = stats; p.stats_groups = groups;
    if (fused) *fused = 1;
  }
  for (int k = 0; k < P.ncls; ++k) { p.cat_o[k][0] = (int8_t)P.cls_a[k][0]; p.cat_o[k][1] = (int8_t)P.cls_a[k][1]; p.cat_o[k][2] = (int8_t)P.cls_a[k][2]; p.cat_o[k][3] = 0; }
  set_bnb(p, bnb, d->N, fused);
  return launch_pix(p, (int)mtiles, bias, large, st);
}

int tc_conv_wgrad(const gg_conv_desc* d, const void* large, const void* small, float* dw, cudaStream_t st) {
  int rc = check_tc(d, large, small, true, true);
  if (rc) return rc;
  static TcWgradParams proto;   // large struct: build on the heap-free static under a lock, copy for the launch
  static std::mutex mu;
  std::lock_guard<std::mutex> lock(mu);
  TcWgradParams& p = proto;
  memset(&p, 0, sizeof(p));
  pick_box(d->Wo, d->Ho, d->Do, WG_PIX, &p.bw, &p.bh, &p.bd, &p.bn);
  const uint32_t box[5] = {64, (uint32_t)p.bw, (uint32_t)p.bh, (uint32_t)p.bd, (uint32_t)p.bn};
  rc = make_large_views(d, large, box, p.lmap);
  if (rc) return rc;
  rc = make_small_map(d, small, box, &p.smap);
  if (rc) return rc;
  p.tw = ceil_div(d->Wo, p.bw); p.th = ceil_div(d->Ho, p.bh); p.td = ceil_div(d->Do, p.bd); p.tn = ceil_div(d->N, p.bn);
  p.ptiles = p.tw * p.th * p.td * p.tn;
  const int cchunks = d->C / 64;
  int na = 0;
  for (int a = 0; a < d->kd; ++a)
    for (int b = 0; b < d->kh; ++b)
      for (int e = 0; e < d->kw; ++e)
        for (int cc = 0; cc < cchunks; ++cc) {
          const int fd = a - d->pd, fh = b - d->ph, fw = e - d->pw;
          WgAtom at;
          at.view = (int8_t)((posmod(fd, d->sd) * d->sh + posmod(fh, d->sh)) * d->sw + posmod(fw, d->sw));
          at.od = (int8_t)floordiv(fd, d->sd); at.oh = (int8_t)floordiv(fh, d->sh); at.ow = (int8_t)floordiv(fw, d->sw);
          at.widx = (int16_t)((a * d->kh + b) * d->kw + e);
          at.c0 = (int16_t)(cc * 64);
          GG_REQUIRE(na < 2 * 208, GG_ERR_UNSUPPORTED, "tensor-core wgrad: too many (tap, channel-chunk) atoms");
          p.atoms[na++] = at;
        }
  p.natoms = na;
  if (na & 1) { WgAtom pad; memset(&pad, 0, sizeof(pad)); pad.c0 = -1; p.atoms[na] = pad; }
  p.mtiles = (na + 1) / 2;
  p.BN = d->K % 256 == 0 ? 256 : (d->K % 128 == 0 ? 128 : 64);
  p.ntiles_n = d->K / p.BN;
  const int64_t tiles = (int64_t)p.mtiles * p.ntiles_n;
  // One CTA per SM (the pipeline takes ~190 KB of shared memory): size the pixel split for ONE full wave.  (The
  // earlier target of 2 x 148 CTAs ran as 2-3 waves, the last one nearly empty: d_h3 100 tiles x 3 = 300 CTAs.)
  // GG_DETERMINISTIC=1: no pixel split -- every dw element then receives exactly ONE reduce-add (onto the zero-filled or
  // previously accumulated gradient), so the filter gradients are bit-reproducible run to run; with splits the partial sums of
  // a tile land in L2 in CTA-finish order (last-bit differences that a GAN amplifies to a few % within 3-4 steps).
  const bool deterministic = env_int("GG_DETERMINISTIC", 0) != 0;
  const int64_t cta_target = deterministic ? 1 : env_int("GG_WG_CTAS", 148);
  int splits = (int)std::max<int64_t>(1, cta_target / tiles);
  splits = std::min(splits, std::max(1, p.ptiles / 4));
  p.ptiles_per_split = ceil_div(p.ptiles, splits);
  p.splits = ceil_div(p.ptiles, p.ptiles_per_split);
  p.C = d->C; p.K = d->K;
  {
    const int taps_total = d->kd * d->kh * d->kw;
    const uint64_t ddims[2] = {(uint64_t)d->K, (uint64_t)taps_total * d->C};
    const uint64_t dstr[1] = {(uint64_t)d->K * 4};
    const uint32_t dbox[2] = {32, 64};
    GG_REQUIRE(((uintptr_t)dw % 16) == 0, GG_ERR_INVALID, "tensor-core wgrad needs a 16-byte aligned gradient buffer");
    rc = encode_tmap(&p.dwmap, GG_F32, dw, 2, ddims, dstr, dbox);
    if (rc) return rc;
  }
  p.pps = std::max(1, std::min(4, env_int("GG_WG_PPS", 2)));      // 8 MMAs per commit (see launch_pix)
  const int stage_bytes = p.pps * (2 + p.BN / 64) * WG_ATOM_BYTES;
  p.stages = std::max(2, std::min(8, (200 * 1024) / stage_bytes));
  GG_REQUIRE((size_t)p.stages * stage_bytes + 2048 <= 227 * 1024, GG_ERR_INVALID, "tc_wgrad: pipeline does not fit in shared memory");
  const size_t smem = (size_t)p.stages * stage_bytes + 1024 + 8 * (2 * p.stages + 2);
  static std::once_flag once;
  std::call_once(once, [] { cudaFuncSetAttribute(tc_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); });
  int rcl = GG_OK;
  for (int r = 0; r < g_repeat && rcl == GG_OK; ++r) {
    Launch((unsigned)(tiles * p.splits), TC_WG_THREADS, smem, st)(tc_wgrad_kernel, p, dw);
    rcl = check_launch("tc_wgrad");
  }
  return rcl;
}

}  // namespace gg
