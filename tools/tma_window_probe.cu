// Standalone sm_100a probe: a 4-D TMA tensor map whose dimension-1 stride (one pixel = cin elements) is SMALLER than
// dimension 0's extent (64 elements = 64/cin consecutive pixels).  Each "row" of a box is then a sliding window of
// 64/cin horizontally adjacent pixels x cin channels — the im2col row of a small-channel conv, fetched with no
// materialisation.  Checks (a) cuTensorMapEncodeTiled accepts the overlapping strides, (b) the box lands as expected
// in SWIZZLE_128B order, (c) rows with a negative / overflowing y coordinate are zero-filled, (d) elementStrides on y.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o build/tma_window_probe tools/tma_window_probe.cu
// Run  : build/tma_window_probe <cin> <stride>
#include <cuda.h>
#include <cuda_bf16.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../e2fgvi_b200/csrc/common.cuh"

using namespace e2f;

__device__ __forceinline__ void tma_load_4d(uint32_t smem_dst, const void* tmap, uint64_t* bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap tm, uint16_t* __restrict__ out, int x0, int y0, int n) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 16384);
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar, 16384);
    tma_load_4d(smem_u32(smem), &tm, bar, 0, x0, y0, n);
  }
  mbar_wait(bar, 0);
  const int r = threadIdx.x;                       // row r of the 128-row tile, de-swizzled on the way out
  for (int c = 0; c < 8; ++c) {
    const uint4 v = *reinterpret_cast<const uint4*>(smem + sw128_offset(r, c));
    *reinterpret_cast<uint4*>(out + r * 64 + c * 8) = v;
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
  const int cin = argc > 1 ? atoi(argv[1]) : 8, stride = argc > 2 ? atoi(argv[2]) : 1;
  const int N = 2, H = 24, P = 80;                 // P: row pitch in pixels (data + zero gap)
  const size_t elems = static_cast<size_t>(N) * H * P * cin + 64;
  std::vector<uint16_t> h(elems);
  for (size_t i = 0; i < elems; ++i) h[i] = static_cast<uint16_t>(1 + (i * 2654435761u >> 7) % 60000);
  uint16_t *d_in, *d_out;
  cudaMalloc(&d_in, elems * 2);
  cudaMalloc(&d_out, 128 * 64 * 2);
  cudaMemcpy(d_in, h.data(), elems * 2, cudaMemcpyHostToDevice);
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(fp);
  CUtensorMap tm;
  // dim0: 64 consecutive elements; dim1: window start, one step = `stride` pixels; dim2: rows; dim3: images
  const cuuint64_t dims[4] = {64, static_cast<cuuint64_t>(P / stride), static_cast<cuuint64_t>(H), static_cast<cuuint64_t>(N)};
  const cuuint64_t strides[3] = {static_cast<cuuint64_t>(stride) * cin * 2, static_cast<cuuint64_t>(P) * cin * 2,
                                 static_cast<cuuint64_t>(H) * P * cin * 2};
  const cuuint32_t box[4] = {64, 16, static_cast<cuuint32_t>(8 * stride), 1};
  const cuuint32_t estr[4] = {1, 1, static_cast<cuuint32_t>(stride), 1};
  CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, d_in, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode: CUresult %d (cin=%d stride=%d dim1 stride %d B)\n", static_cast<int>(r), cin, stride, stride * cin * 2);
  if (r != CUDA_SUCCESS) return 2;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 + 1024 + 64);
  int bad_total = 0;
  const int cases[3][3] = {{5, 3, 1}, {0, -2 * 1, 0}, {40, H - 3, 1}};
  for (int t = 0; t < 3; ++t) {
    const int x0 = cases[t][0], y0 = cases[t][1] * (t == 1 ? stride : 1), n = cases[t][2];
    probe_kernel<<<1, 128, 16384 + 1024 + 64>>>(tm, d_out, x0, y0, n);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("kernel failed: %s\n", cudaGetErrorString(e));
      return 3;
    }
    std::vector<uint16_t> o(128 * 64);
    cudaMemcpy(o.data(), d_out, o.size() * 2, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int ry = 0; ry < 8; ++ry)
      for (int rx = 0; rx < 16; ++rx)
        for (int e2 = 0; e2 < 64; ++e2) {
          const int y = y0 + ry * stride, i = x0 + rx;
          uint16_t want = 0;
          if (y >= 0 && y < H && i >= 0 && i < P / stride)
            want = h[(static_cast<size_t>(n) * H + y) * P * cin + static_cast<size_t>(i) * stride * cin + e2];
          if (o[(ry * 16 + rx) * 64 + e2] != want) ++bad;
        }
    printf("case %d (x0=%d y0=%d n=%d): %d mismatches\n", t, x0, y0, n, bad);
    bad_total += bad;
  }
  printf("TMA_WINDOW_PROBE cin=%d stride=%d %s\n", cin, stride, bad_total ? "FAIL" : "PASS");
  return bad_total ? 1 : 0;
}
