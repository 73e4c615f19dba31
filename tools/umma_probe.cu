// Standalone sm_100a probe: validates the tcgen05 / TMA / TMEM building blocks (descriptor encodings,
// manual 128B swizzle vs TMA's, MN-major B operand, tcgen05.st/ld round trip) against a host GEMM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o build/umma_probe tools/umma_probe.cu
// Run  : build/umma_probe <mode> [swap]      (one mode per process so a trap cannot poison the next)
//   mode 0: A K-major (cp.async, manual swizzle)  B K-major via TMA SWIZZLE_128B
//   mode 1: A K-major (cp.async)                  B K-major (cp.async, manual swizzle)
//   mode 2: A K-major                             B MN-major ([k][n] rows, manual swizzle); swap=1 swaps LBO/SBO
//   mode 3: mode 1 + TMEM round trip (ld, x2, st, ld)
//   mode 4: mode 1 with N=64 instruction shape (only the first 64 columns are produced)
//   mode 5: A operand from TENSOR MEMORY (tcgen05.mma [d], [a_tmem], b_desc): A is written to TMEM columns 128.. by
//           tcgen05.st as packed fp16 pairs (column j of lane m holds A[m][2j] (low half) | A[m][2j+1] (high half));
//           B K-major in smem.  swap=1 tries the opposite half order.
//   mode 6: B rows gathered by TMA tile::gather4 (4 arbitrary rows per instruction, here in bit-reversed order) into
//           the K-major SWIZZLE_128B tile; tensor-map box {64, 1} (swap=0) or {64, 4} (swap=1).
#include <cuda.h>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "../e2fgvi_b200/csrc/common.cuh"

using namespace e2f;

constexpr int M = 128, N = 128, K = 128;

__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap tmapB, const __grid_constant__ CUtensorMap tmapG,
             const __half* __restrict__ A,
             const __half* __restrict__ B, float* __restrict__ D, int mode, int swap) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;            // 2 x [128 rows][64 halfs]  (16 KB each)
  uint8_t* sB = smem + 32768;    // same
  uint64_t* bar_tma = reinterpret_cast<uint64_t*>(smem + 65536);
  uint64_t* bar_mma = bar_tma + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_tma + 2);

  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) tmem_alloc(tmem_slot, 256);
  if (tid == 0) {
    mbar_init(bar_tma, 1);
    mbar_init(bar_mma, 1);
    fence_barrier_init();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tbase = *tmem_slot;

  // A: thread r owns row r
  for (int c = 0; c < 16; ++c)
    cp_async16(smem_u32(sA) + (c >> 3) * 16384 + sw128_offset(tid, c & 7), A + tid * K + c * 8);
  if (mode == 0) {
    if (tid == 0) {
      mbar_arrive_expect_tx(bar_tma, 32768);
      tma_load_2d(smem_u32(sB), &tmapB, bar_tma, 0, 0);
      tma_load_2d(smem_u32(sB) + 16384, &tmapB, bar_tma, 64, 0);
    }
  } else if (mode == 6) {
    if (tid == 0) {
      mbar_arrive_expect_tx(bar_tma, 32768);
      for (int half = 0; half < 2; ++half)
        for (int i = 0; i < 32; ++i) {            // smem rows 4i..4i+3 <- global rows rev7(4i+j)
          int r[4];
          for (int j = 0; j < 4; ++j) {
            const int n = 4 * i + j;
            r[j] = ((n & 1) << 6) | ((n & 2) << 4) | ((n & 4) << 2) | (n & 8) | ((n & 16) >> 2) | ((n & 32) >> 4) | ((n & 64) >> 6);
          }
          asm volatile(
              "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes"
              " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(smem_u32(sB) + half * 16384 + i * 512),
              "l"(reinterpret_cast<uint64_t>(&tmapG)), "r"(smem_u32(bar_tma)), "r"(half * 64), "r"(r[0]), "r"(r[1]),
              "r"(r[2]), "r"(r[3])
              : "memory");
        }
    }
  } else {
    // mode 1/3/4: B is [n][k] row-major, thread n owns row n.  mode 2: B is [k][n] row-major, thread k owns row k.
    for (int c = 0; c < 16; ++c)
      cp_async16(smem_u32(sB) + (c >> 3) * 16384 + sw128_offset(tid, c & 7), B + tid * 128 + c * 8);
  }
  cp_async_commit();
  cp_async_wait<0>();
  fence_proxy_async_smem();
  __syncthreads();

  if (mode == 5) {
    // thread m writes its row of A (128 halfs = 64 packed columns) to TMEM lanes [32*warp, +32), columns 128..191
    const uint32_t lane_base5 = static_cast<uint32_t>(warp * 32) << 16;
    for (int c = 0; c < 2; ++c) {
      uint32_t pk[32];
      for (int j = 0; j < 32; ++j) {
        const __half lo = A[tid * K + 2 * (c * 32 + j) + (swap ? 1 : 0)];
        const __half hi = A[tid * K + 2 * (c * 32 + j) + (swap ? 0 : 1)];
        pk[j] = static_cast<uint32_t>(__half_as_ushort(lo)) | (static_cast<uint32_t>(__half_as_ushort(hi)) << 16);
      }
      tmem_st32(tbase + lane_base5 + 128 + c * 32, pk);
    }
    tmem_st_wait();
    tc_fence_before_sync();
    __syncthreads();
  }
  if (tid == 0) {
    if (mode == 0 || mode == 6) mbar_wait(bar_tma, 0);
    tc_fence_after_sync();
    if (mode == 5) {
      const uint32_t idesc5 = umma_idesc_f16(128, 128, 0, 0);
      for (int k = 0; k < K / 16; ++k) {
        const uint64_t bd = umma_desc_sw128(smem_u32(sB) + (k >> 2) * 16384 + (k & 3) * 32, 16, 1024);
        const uint32_t a_tmem = tbase + 128 + k * 8;      // 16 halfs = 8 packed 32-bit columns per K step
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "setp.ne.b32 p, %4, 0;\n"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
            "}\n" ::"r"(tbase), "r"(a_tmem), "l"(bd), "r"(idesc5), "r"(k > 0 ? 1u : 0u)
            : "memory");
      }
      umma_commit(bar_mma);
    }
    const int n_inst = (mode == 4) ? 64 : 128;
    const uint32_t idesc = umma_idesc_f16(128, n_inst, 0, mode == 2 ? 1 : 0);
    for (int k = 0; k < (mode == 5 ? 0 : K / 16); ++k) {
      const uint64_t ad = umma_desc_sw128(smem_u32(sA) + (k >> 2) * 16384 + (k & 3) * 32, 16, 1024);
      uint64_t bd;
      if (mode == 2) {
        const uint32_t lbo = swap ? 1024 : 16384, sbo = swap ? 16384 : 1024;
        bd = umma_desc_sw128(smem_u32(sB) + k * 2048, lbo, sbo);
      } else {
        bd = umma_desc_sw128(smem_u32(sB) + (k >> 2) * 16384 + (k & 3) * 32, 16, 1024);
      }
      umma_f16(tbase, ad, bd, idesc, k > 0);
    }
    if (mode != 5) umma_commit(bar_mma);
  }
  mbar_wait(bar_mma, 0);
  tc_fence_after_sync();

  const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;
  uint32_t v[32];
  if (mode == 3) {
    for (int c = 0; c < 4; ++c) {
      tmem_ld32(tbase + lane_base + c * 32, v);
      tmem_ld_wait();
      for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(2.0f * __uint_as_float(v[i]));
      tmem_st32(tbase + lane_base + c * 32, v);
      tmem_st_wait();
    }
  }
  for (int c = 0; c < 4; ++c) {
    tmem_ld32(tbase + lane_base + c * 32, v);
    tmem_ld_wait();
    for (int i = 0; i < 32; ++i) D[tid * N + c * 32 + i] = __uint_as_float(v[i]);
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tbase, 256);
}

#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e = (x);                                                               \
    if (e != cudaSuccess) {                                                            \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__);   \
      return 2;                                                                        \
    }                                                                                  \
  } while (0)

typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
  const int mode = argc > 1 ? atoi(argv[1]) : 0;
  const int swap = argc > 2 ? atoi(argv[2]) : 0;
  std::vector<__half> hA(M * K), hB(128 * 128);
  std::vector<float> fA(M * K), fB(128 * 128);
  srand(1234);
  for (int i = 0; i < M * K; ++i) {
    float x = (rand() % 2001 - 1000) / 1000.0f;
    hA[i] = __float2half(x);
    fA[i] = __half2float(hA[i]);
  }
  for (int i = 0; i < 128 * 128; ++i) {
    float x = (rand() % 2001 - 1000) / 1000.0f;
    hB[i] = __float2half(x);
    fB[i] = __half2float(hB[i]);
  }
  __half *dA, *dB;
  float* dD;
  CK(cudaMalloc(&dA, M * K * 2));
  CK(cudaMalloc(&dB, 128 * 128 * 2));
  CK(cudaMalloc(&dD, M * N * 4));
  CK(cudaMemcpy(dA, hA.data(), M * K * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), 128 * 128 * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(dD, 0, M * N * 4));

  CUtensorMap tmap;
  memset(&tmap, 0, sizeof(tmap));
  {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn) {
      printf("no cuTensorMapEncodeTiled\n");
      return 2;
    }
    cuuint64_t dims[2] = {K, 128};          // innermost first
    cuuint64_t strides[1] = {K * 2};        // bytes, dims 1..
    cuuint32_t box[2] = {64, 128};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = reinterpret_cast<EncodeTiled>(fn)(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, dB, dims, strides, box,
                                                   estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      printf("cuTensorMapEncodeTiled failed %d\n", (int)r);
      return 2;
    }
  }
  CUtensorMap tmapG;
  memset(&tmapG, 0, sizeof(tmapG));
  {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    cuuint64_t dims[2] = {K, 128};
    cuuint64_t strides[1] = {K * 2};
    cuuint32_t box[2] = {64, swap ? 4u : 1u};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = reinterpret_cast<EncodeTiled>(fn)(&tmapG, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, dB, dims, strides, box,
                                                   estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      printf("cuTensorMapEncodeTiled (gather map) failed %d\n", (int)r);
      return 2;
    }
  }
  const int smem_bytes = 65536 + 64 + 1024;
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  probe_kernel<<<1, 128, smem_bytes>>>(tmap, tmapG, dA, dB, dD, mode, swap);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  std::vector<float> hD(M * N);
  CK(cudaMemcpy(hD.data(), dD, M * N * 4, cudaMemcpyDeviceToHost));

  const int ncols = (mode == 4) ? 64 : 128;
  double maxerr = 0, maxref = 0;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < ncols; ++n) {
      double ref = 0;
      int nn = n;
      if (mode == 6) nn = ((n & 1) << 6) | ((n & 2) << 4) | ((n & 4) << 2) | (n & 8) | ((n & 16) >> 2) | ((n & 32) >> 4) | ((n & 64) >> 6);
      for (int k = 0; k < K; ++k)
        ref += (double)fA[m * K + k] * (mode == 2 ? (double)fB[k * 128 + n] : (double)fB[nn * K + k]);
      if (mode == 3) ref *= 2.0;
      maxerr = fmax(maxerr, fabs(ref - hD[m * N + n]));
      maxref = fmax(maxref, fabs(ref));
    }
  printf("PROBE mode=%d swap=%d max_abs_err=%.6g max_ref=%.6g %s\n", mode, swap, maxerr, maxref,
         maxerr < 1e-2 ? "PASS" : "FAIL");
  return maxerr < 1e-2 ? 0 : 1;
}
