// Standalone sm_100a probe for the conv kernel's HALO mode: one TMA box {64 ch, 10 x, 18 y} (SWIZZLE_128B) holds the
// (8+2) x (16+2) input halo of an 8 x 16 output-pixel tile; the A operand of tap (dy, dx) is then the SAME shared-memory
// tile read through a UMMA descriptor whose start address is shifted by (dy*10 + dx) pixels (128 B each, NOT a multiple
// of the 1024-byte swizzle atom) and whose stride between 8-row groups (SBO) is one halo row = 1280 B.
// Question answered: does tcgen05.mma apply the 128B swizzle XOR on absolute shared-memory address bits (so a shifted
// start with base_offset = 0 just works, mode 0), or does it need base_offset = (start >> 7) & 7 (mode 1)?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o build/halo_probe tools/halo_probe.cu
// Run  : build/halo_probe <mode>
#include <cuda.h>
#include <cuda_bf16.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../e2fgvi_b200/csrc/common.cuh"

using namespace e2f;

constexpr int HW_H = 24, HW_W = 20, C = 64, NOUT = 64;
constexpr int TW = 8, TH = 16, HALO_W = TW + 2, HALO_H = TH + 2;
constexpr int A_BYTES = HALO_W * HALO_H * 128;          // 23040
constexpr int A_PAD = 23552;                            // rounded up to 1024
constexpr int W_TILE = NOUT * 128;                      // 8 KB per tap

__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
}

__device__ __forceinline__ void tma_load_4d(uint32_t smem_dst, const void* tmap, uint64_t* bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, float* __restrict__ D,
             int x0, int y0, int mode) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sW = smem + A_PAD;
  uint64_t* bar_tma = reinterpret_cast<uint64_t*>(sW + 9 * W_TILE);
  uint64_t* bar_mma = bar_tma + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_tma + 2);
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) tmem_alloc(tmem_slot, 64);
  if (tid == 0) {
    mbar_init(bar_tma, 1);
    mbar_init(bar_mma, 1);
    fence_barrier_init();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tbase = *tmem_slot;
  if (tid == 0) {
    mbar_arrive_expect_tx(bar_tma, A_BYTES + 9 * W_TILE);
    tma_load_4d(smem_u32(sA), &tmA, bar_tma, 0, x0 - 1, y0 - 1, 0);
    for (int t = 0; t < 9; ++t) tma_load_2d(smem_u32(sW) + t * W_TILE, &tmW, bar_tma, 0, t * NOUT);
    mbar_wait(bar_tma, 0);
    tc_fence_after_sync();
    const uint32_t idesc = idesc_bf16(128, NOUT);
    for (int tap = 0; tap < 9; ++tap) {
      const int dy = tap / 3, dx = tap % 3;
      const uint32_t a_addr = smem_u32(sA) + (dy * HALO_W + dx) * 128;
      uint64_t da = umma_desc_sw128(a_addr, 16, HALO_W * 128);
      if (mode == 1) da |= static_cast<uint64_t>((a_addr >> 7) & 7) << 49;
      const uint64_t dw = umma_desc_sw128(smem_u32(sW) + tap * W_TILE, 16, 1024);
      for (int k = 0; k < 4; ++k) umma_f16(tbase, da + 2 * k, dw + 2 * k, idesc, (tap | k) != 0);
    }
    umma_commit(bar_mma);
  }
  __syncwarp();
  mbar_wait(bar_mma, 0);
  tc_fence_after_sync();
  const uint32_t taddr = tbase + (static_cast<uint32_t>(warp * 32) << 16);
  for (int c = 0; c < NOUT / 32; ++c) {
    uint32_t v[32];
    tmem_ld32(taddr + c * 32, v);
    tmem_ld_wait();
    for (int i = 0; i < 32; ++i) D[tid * NOUT + c * 32 + i] = __uint_as_float(v[i]);
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tbase, 64);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
  const int mode = argc > 1 ? atoi(argv[1]) : 0;
  std::vector<float> x(HW_H * HW_W * C), w(9 * NOUT * C);
  std::vector<__nv_bfloat16> xb(x.size()), wb(w.size());
  uint32_t s = 12345;
  auto rnd = [&]() { s = s * 1664525u + 1013904223u; return static_cast<float>(static_cast<int>((s >> 20) % 9) - 4); };
  for (size_t i = 0; i < x.size(); ++i) { x[i] = rnd(); xb[i] = __float2bfloat16(x[i]); }
  for (size_t i = 0; i < w.size(); ++i) { w[i] = rnd(); wb[i] = __float2bfloat16(w[i]); }
  __nv_bfloat16 *dx, *dw;
  float* dD;
  cudaMalloc(&dx, xb.size() * 2);
  cudaMalloc(&dw, wb.size() * 2);
  cudaMalloc(&dD, 128 * NOUT * 4);
  cudaMemcpy(dx, xb.data(), xb.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dw, wb.data(), wb.size() * 2, cudaMemcpyHostToDevice);
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(fp);
  CUtensorMap tmA, tmW;
  {
    const cuuint64_t dims[4] = {C, HW_W, HW_H, 1};
    const cuuint64_t strides[3] = {C * 2, HW_W * C * 2, static_cast<cuuint64_t>(HW_H) * HW_W * C * 2};
    const cuuint32_t box[4] = {64, HALO_W, HALO_H, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, dx, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode A halo box: CUresult %d\n", static_cast<int>(r));
    if (r != CUDA_SUCCESS) return 2;
  }
  {
    const cuuint64_t dims[2] = {C, 9 * NOUT};
    const cuuint64_t strides[1] = {C * 2};
    const cuuint32_t box[2] = {64, NOUT};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&tmW, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dw, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode W: CUresult %d\n", static_cast<int>(r));
    if (r != CUDA_SUCCESS) return 2;
  }
  const int smem_bytes = A_PAD + 9 * W_TILE + 64 + 1024;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  const int cases[3][2] = {{3, 2}, {0, 0}, {HW_W - TW, HW_H - TH}};
  int bad_total = 0;
  for (int t = 0; t < 3; ++t) {
    const int x0 = cases[t][0], y0 = cases[t][1];
    cudaMemset(dD, 0xff, 128 * NOUT * 4);
    probe_kernel<<<1, 128, smem_bytes>>>(tmA, tmW, dD, x0, y0, mode);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("kernel failed: %s\n", cudaGetErrorString(e));
      return 3;
    }
    std::vector<float> d(128 * NOUT);
    cudaMemcpy(d.data(), dD, d.size() * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    double maxd = 0;
    for (int r = 0; r < 128; ++r) {
      const int y = y0 + r / TW, xx = x0 + r % TW;
      for (int co = 0; co < NOUT; ++co) {
        float acc = 0.f;
        for (int tap = 0; tap < 9; ++tap) {
          const int iy = y - 1 + tap / 3, ix = xx - 1 + tap % 3;
          if (iy < 0 || iy >= HW_H || ix < 0 || ix >= HW_W) continue;
          for (int c = 0; c < C; ++c) acc += x[(iy * HW_W + ix) * C + c] * w[(tap * NOUT + co) * C + c];
        }
        const double df = fabs(static_cast<double>(acc) - d[r * NOUT + co]);
        if (df > maxd) maxd = df;
        if (df > 1e-3) ++bad;
      }
    }
    printf("case %d (x0=%d y0=%d) mode %d: %d mismatches, max diff %.3f\n", t, x0, y0, mode, bad, maxd);
    bad_total += bad;
  }
  printf("HALO_PROBE mode=%d %s\n", mode, bad_total ? "FAIL" : "PASS");
  return bad_total ? 1 : 0;
}
