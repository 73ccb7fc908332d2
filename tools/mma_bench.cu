// Microbenchmark for the tcgen05 main loop (measurement tool, not product code):
//   how many cycles does ONE tcgen05.mma.cta_group::1.kind::f16 (M=128, N, K=16) cost when operands come from
//   shared memory (SS mode, 128B-swizzled K-major), alone and with a bulk-copy producer refilling the stages?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/mma_bench tools/mma_bench.cu
// Run  : tools/bin/mma_bench            (prints one line per configuration)
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "../gif-gan_b200/csrc/tc_common.cuh"

using namespace gg::tc;

namespace gg { void set_error(const char*, ...) {} std::atomic<uint64_t> g_launches{0}; }

constexpr int A_BYTES = 128 * 128;

struct Args {
  int N, iters, stages, mode;   // mode 0: MMA only; 1: + bulk-copy producer (A and B per stage); 2: producer only (no MMA)
  int commit_every, alt;        // (conv path) commit once per this many iterations; alternate between two accumulators
  int conv;                     // 1: converged-warp loops with elect.sync around the issue; 0: loops inside if (lane == 0)
  int kblocks;                  // MMAs (K=16) per stage: 4 = one 128-byte swizzle row
  const uint8_t* src;
  unsigned long long* out;      // per CTA: [issue_cycles, total_cycles]
};

__global__ void __launch_bounds__(192, 1) mma_bench_kernel(Args a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int stage_bytes = A_BYTES + a.N * 128;
  const uint32_t bar_base = smem_base + a.stages * stage_bytes;
  const uint32_t done_bar = bar_base + 8u * (2 * a.stages);
  const uint32_t tmem_slot = bar_base + 8u * (2 * a.stages + 1);
  if (threadIdx.x == 0) {
    for (int s = 0; s < a.stages; ++s) { mbar_init(bar_base + 8u * s, 1); mbar_init(bar_base + 8u * (a.stages + s), 1); }
    mbar_init(done_bar, 1);
    fence_barrier_init();
  }
  // zero the operand buffers (finite data)
  for (uint32_t i = threadIdx.x * 16; i < (uint32_t)(a.stages * stage_bytes); i += blockDim.x * 16)
    asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(smem_base + i), "r"(0x3c003c00u) : "memory");
  fence_proxy_async();
  const uint32_t tmem_cols = a.alt ? 2 * a.N : (a.N < 32 ? 32 : a.N);
  if (warp == 1) tmem_alloc(tmem_slot, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  const long long t0 = clock64();
  if (a.conv && warp == 0 && a.mode >= 1) {
    int s = 0; uint32_t ph = 1;
    const uint8_t* src = a.src + (size_t)blockIdx.x * stage_bytes;
    for (int it = 0; it < a.iters; ++it) {
      if (a.mode == 1) mbar_wait(bar_base + 8u * (a.stages + s), ph);
      if (a.mode == 2 && it >= a.stages) mbar_wait(bar_base + 8u * s, (uint32_t)(((it / a.stages) - 1) & 1));
      if (elect_one()) {
        mbar_expect_tx(bar_base + 8u * s, (uint32_t)stage_bytes);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_base + s * stage_bytes), "l"(src), "r"((uint32_t)stage_bytes), "r"(bar_base + 8u * s) : "memory");
      }
      __syncwarp();
      if (++s == a.stages) { s = 0; ph ^= 1u; }
    }
    if (a.mode == 2) {
      for (int j = 0; j < a.stages; ++j) { const int it = a.iters - a.stages + j; mbar_wait(bar_base + 8u * (it % a.stages), (uint32_t)((it / a.stages) & 1)); }
      if (lane == 0) { a.out[2 * blockIdx.x] = 0; a.out[2 * blockIdx.x + 1] = (unsigned long long)(clock64() - t0); }
    }
  } else if (a.conv && warp == 1 && a.mode != 2) {
    const uint32_t idesc = make_idesc_bf16(128, a.N, 0, 0);
    const uint64_t desc_hi = make_smem_desc(0, 16, 1024);
    int s = 0; uint32_t ph = 0;
    for (int it = 0; it < a.iters; ++it) {
      if (a.mode == 1) { mbar_wait(bar_base + 8u * s, ph); tc_fence_after(); }
      const uint32_t a_addr = smem_base + s * stage_bytes;
      const uint64_t ad = desc_hi | (uint64_t)((a_addr & 0x3FFFFu) >> 4);
      const uint64_t bd = desc_hi | (uint64_t)(((a_addr + A_BYTES) & 0x3FFFFu) >> 4);
      if (elect_one()) {
        const uint32_t alt = a.alt ? (uint32_t)a.N : 0u;
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tmem_base + (k & 1) * alt, ad + 2u * k, bd + 2u * k, idesc, (it > 0 || k > 1) ? 1u : 0u);
        for (int rep = 4; rep < a.kblocks; rep += 4) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(tmem_base + (k & 1) * alt, ad + 2u * k, bd + 2u * k, idesc, 1u);
        }
        if ((it + 1) % a.commit_every == 0) umma_commit(bar_base + 8u * (a.stages + s));
      }
      __syncwarp();
      if (++s == a.stages) { s = 0; ph ^= 1u; }
    }
    if (elect_one()) umma_commit(done_bar);
    __syncwarp();
    const long long t1 = clock64();
    mbar_wait(done_bar, 0);
    const long long t2 = clock64();
    if (lane == 0) { a.out[2 * blockIdx.x] = (unsigned long long)(t1 - t0); a.out[2 * blockIdx.x + 1] = (unsigned long long)(t2 - t0); }
  } else if (!a.conv && warp == 0 && lane == 0 && a.mode >= 1) {
    // producer: 1-D bulk copies global -> shared, completing on full[s]
    int s = 0; uint32_t ph = 1;
    const uint8_t* src = a.src + (size_t)blockIdx.x * stage_bytes;
    for (int it = 0; it < a.iters; ++it) {
      if (a.mode == 1) mbar_wait(bar_base + 8u * (a.stages + s), ph);
      if (a.mode == 2 && it >= a.stages) mbar_wait(bar_base + 8u * s, (uint32_t)(((it / a.stages) - 1) & 1));   // keep `stages` copies in flight
      mbar_expect_tx(bar_base + 8u * s, (uint32_t)stage_bytes);
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(smem_base + s * stage_bytes), "l"(src), "r"((uint32_t)stage_bytes), "r"(bar_base + 8u * s) : "memory");
      if (++s == a.stages) { s = 0; ph ^= 1u; }
    }
    if (a.mode == 2) {
      for (int j = 0; j < a.stages; ++j) { const int it = a.iters - a.stages + j; mbar_wait(bar_base + 8u * (it % a.stages), (uint32_t)((it / a.stages) & 1)); }
      a.out[2 * blockIdx.x] = 0; a.out[2 * blockIdx.x + 1] = (unsigned long long)(clock64() - t0);
    }
  } else if (!a.conv && warp == 1 && lane == 0 && a.mode != 2) {
    const uint32_t idesc = make_idesc_bf16(128, a.N, 0, 0);
    const uint64_t desc_hi = make_smem_desc(0, 16, 1024);
    int s = 0; uint32_t ph = 0;
    for (int it = 0; it < a.iters; ++it) {
      if (a.mode == 1) { mbar_wait(bar_base + 8u * s, ph); tc_fence_after(); }
      const uint32_t a_addr = smem_base + s * stage_bytes;
      const uint64_t ad = desc_hi | (uint64_t)((a_addr & 0x3FFFFu) >> 4);
      const uint64_t bd = desc_hi | (uint64_t)(((a_addr + A_BYTES) & 0x3FFFFu) >> 4);
      for (int k = 0; k < a.kblocks; ++k) umma_bf16(tmem_base, ad + 2u * (k & 3), bd + 2u * (k & 3), idesc, (it > 0 || k > 0) ? 1u : 0u);
      umma_commit(bar_base + 8u * (a.stages + s));
      if (++s == a.stages) { s = 0; ph ^= 1u; }
    }
    umma_commit(done_bar);
    const long long t1 = clock64();
    mbar_wait(done_bar, 0);
    const long long t2 = clock64();
    a.out[2 * blockIdx.x] = (unsigned long long)(t1 - t0);
    a.out[2 * blockIdx.x + 1] = (unsigned long long)(t2 - t0);
  }
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, tmem_cols); }
}

int main() {
  const int stages_max = 4;
  uint8_t* src; unsigned long long* out;
  cudaMalloc(&src, (size_t)148 * 64 * 1024);
  cudaMemset(src, 0, (size_t)148 * 64 * 1024);
  cudaMalloc(&out, 148 * 2 * sizeof(unsigned long long));
  cudaFuncSetAttribute(mma_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  unsigned long long h[296];
  const int Ns[3] = {64, 128, 256};
  struct V { int mode, kb, ce, alt; };
  const V variants[] = {{0, 4, 1, 0}, {0, 8, 1, 0}, {0, 16, 1, 0}, {0, 4, 4, 0}, {0, 4, 1000000, 0}, {0, 4, 1, 1}, {0, 8, 1, 1}, {1, 4, 1, 0}, {1, 8, 1, 0}, {2, 4, 1, 0}};
  for (int conv : {1})
  for (int grid : {148})
    for (const V& v : variants)
      for (int ni = 0; ni < 3; ++ni)
        {
          const int mode = v.mode, kb = v.kb;
          Args a; a.conv = conv; a.commit_every = v.ce; a.alt = v.alt; a.N = Ns[ni]; a.iters = 400; a.stages = stages_max; a.mode = mode; a.kblocks = kb; a.src = src; a.out = out;
          const int stage_bytes = A_BYTES + a.N * 128;
          const size_t smem = (size_t)a.stages * stage_bytes + 1024 + 8 * (2 * a.stages + 2) + 16;
          cudaMemset(out, 0, sizeof(h));
          cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
          mma_bench_kernel<<<grid, 192, smem>>>(a);   // warm-up
          cudaEventRecord(e0);
          mma_bench_kernel<<<grid, 192, smem>>>(a);
          cudaEventRecord(e1);
          cudaError_t err = cudaDeviceSynchronize();
          if (err != cudaSuccess) { printf("ERROR %s\n", cudaGetErrorString(err)); return 1; }
          float ms; cudaEventElapsedTime(&ms, e0, e1);
          cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
          double issue = 0, total = 0;
          for (int c = 0; c < grid; ++c) { issue += h[2 * c]; total += h[2 * c + 1]; }
          issue /= grid; total /= grid;
          const double nm = (double)a.iters * kb;
          printf("commit_every=%d alt=%d conv=%d grid=%3d mode=%d N=%3d mma_per_stage=%d stage_bytes=%6d : issue %.1f cyc/mma, total %.1f cyc/mma (%.0f cyc/stage, %.1f B/cyc/SM), kernel %.1f us\n",
                 v.ce > 1000 ? 0 : v.ce, v.alt, conv, grid, mode, a.N, kb, stage_bytes, issue / nm, total / nm, total / a.iters, (double)stage_bytes * a.iters / total, ms * 1e3);
        }
  return 0;
}
