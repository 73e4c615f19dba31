// Standalone sm_100a probe: tensor-pipe time of ONE tcgen05.mma (cta_group::1, kind::f16, bf16 operands, M = 128, K = 16)
// as a function of the N extent and of where the A operand lives (shared memory "SS" / tensor memory "TS").
// One CTA; one thread issues REPS back-to-back MMAs on operands that already sit in shared memory (contents are
// irrelevant), commits, and the wall time between the first issue and the commit's arrival is read with clock64.
// Answers: is a small-N MMA proportionally cheaper (math-bound, N/2 cycles), or does every instruction pay a floor
// (operand fetch from shared memory: A = 4 KB + B = 32*N bytes per instruction)?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o build/mma_rate_probe tools/mma_rate_probe.cu
// Run  : build/mma_rate_probe
#include <cuda.h>
#include <cstdio>
#include <cstdlib>
#include "../e2fgvi_b200/csrc/common.cuh"

using namespace e2f;

__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
}

__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(d), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}

// mode 0: SS, every MMA re-reads the same A (k slice 0) and B          mode 1: TS (A in TMEM columns 256..)
// mode 2: SS, A and B descriptors walk over 4 K slices and 3 stages like the conv main loop
// mode 3: SS pairs like the conv kernel: one N = 2n MMA + one N = n MMA per step (reports cycles per PAIR)
__global__ void __launch_bounds__(128, 1) rate_kernel(long long* out, int n, int mode, int reps) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  // 3 stages x (A 16 KB + B 32 KB)
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 3 * 49152);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 3 * 49152 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tbase = *tmem_slot;
  if (tid == 0) {
    const uint32_t idesc = idesc_bf16(128, n), idesc2 = idesc_bf16(128, 2 * n);
    const uint64_t da0 = umma_desc_sw128(smem_u32(smem), 16, 1024);
    const uint64_t db0 = umma_desc_sw128(smem_u32(smem) + 16384, 16, 1024);
    const long long t0 = clock64();
    for (int i = 0; i < reps; ++i) {
      if (mode == 0) {
        umma_f16(tbase, da0, db0, idesc, 1);
      } else if (mode == 1) {
        umma_ts(tbase, tbase + 256 + (i & 3) * 8, db0 + 2 * (i & 3), idesc, 1);
      } else if (mode == 2) {
        const uint32_t off = ((i >> 2) % 3) * (49152 >> 4) + 2 * (i & 3);
        umma_f16(tbase, da0 + off, db0 + off, idesc, 1);
      } else {
        const uint32_t off = ((i >> 2) % 3) * (49152 >> 4) + 2 * (i & 3);
        umma_f16(tbase, da0 + off, db0 + off, idesc2, 1);
        umma_f16(tbase, da0 + off, db0 + off, idesc, 1);
      }
    }
    umma_commit(bar);
    mbar_wait(bar, 0);
    const long long t1 = clock64();
    out[0] = t1 - t0;
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tbase, 512);
}


// mode 4: TWO issuing threads of one CTA (lane 0 of warps 0 and 1), `reps` MMAs each, disjoint accumulators.
// mode 5: every CTA of a grid with 2 CTAs per SM issues `reps` MMAs (same-operand SS); reports the slowest CTA.
__global__ void __launch_bounds__(128, 2) rate2_kernel(long long* out, int n, int mode, int reps) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 49152);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 2);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 49152 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (warp == 0) tmem_alloc(tmem_slot, 256);
  if (tid == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    fence_barrier_init();
  }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tbase = *tmem_slot;
  const int issuers = mode == 4 ? 2 : 1;
  if ((tid & 31) == 0 && warp < issuers) {
    const uint32_t idesc = idesc_bf16(128, n);
    const uint64_t da0 = umma_desc_sw128(smem_u32(smem), 16, 1024);
    const uint64_t db0 = umma_desc_sw128(smem_u32(smem) + 16384, 16, 1024);
    const long long t0 = clock64();
    for (int i = 0; i < reps; ++i) umma_f16(tbase + warp * 128, da0, db0, idesc, 1);
    umma_commit(&bar[warp]);
    mbar_wait(&bar[warp], 0);
    const long long t1 = clock64();
    out[blockIdx.x * 2 + warp] = t1 - t0;
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tbase, 256);
}

int main() {
  long long* d;
  cudaMalloc(&d, 8);
  const int smem_bytes = 3 * 49152 + 64 + 1024;
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  const int reps = 4096;
  const char* names[4] = {"SS same operands", "TS (A in TMEM)", "SS walking 3 stages", "SS pair N=2n + N=n"};
  for (int mode = 0; mode < 4; ++mode)
    for (int n = 16; n <= 256; n *= 2) {
      if (mode == 3 && n > 128) continue;
      long long best = 1LL << 60;
      for (int r = 0; r < 3; ++r) {
        rate_kernel<<<1, 128, smem_bytes>>>(d, n, mode, reps);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
          printf("kernel failed: %s\n", cudaGetErrorString(e));
          return 3;
        }
        long long h;
        cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
        if (h < best) best = h;
      }
      printf("MMA_RATE %-22s N=%3d : %7.1f cycles per %s (math: %d)\n", names[mode], n, static_cast<double>(best) / reps,
             mode == 3 ? "pair" : "MMA", mode == 3 ? (2 * n + n) / 2 : n / 2);
    }
  {
    long long* d2;
    cudaMalloc(&d2, 8 * 2 * 296);
    const int smem2 = 49152 + 64 + 1024;
    cudaFuncSetAttribute(rate2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2);
    for (int cfg = 0; cfg < 3; ++cfg) {
      const int mode = cfg == 0 ? 4 : 5, grid = cfg == 2 ? 296 : (cfg == 1 ? 148 : 1);
      for (int n = 64; n <= 128; n *= 2) {
        cudaMemset(d2, 0, 8 * 2 * 296);
        rate2_kernel<<<grid, 128, smem2>>>(d2, n, mode, reps);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
          printf("kernel failed: %s\n", cudaGetErrorString(e));
          return 3;
        }
        static long long h[2 * 296];
        cudaMemcpy(h, d2, sizeof(h), cudaMemcpyDeviceToHost);
        long long worst = 0;
        for (int i = 0; i < 2 * 296; ++i) worst = h[i] > worst ? h[i] : worst;
        printf("MMA_RATE %-34s N=%3d : %7.1f cycles per MMA per issuer\n",
               cfg == 0 ? "2 issuing threads, 1 CTA" : (cfg == 1 ? "1 CTA/SM x 148 (1 issuer each)" : "2 CTAs/SM x 148 (1 issuer each)"),
               n, static_cast<double>(worst) / reps);
      }
    }
  }
  return 0;
}
