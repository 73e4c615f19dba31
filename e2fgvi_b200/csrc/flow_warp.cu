// flow_warp: bilinear gather at (x + u, y + v)  — reference model/modules/flow_comp.py:345-383.
// HBM-bound: per output pixel 4 corner reads of C channels + one write.  NHWC layout makes every corner read a
// contiguous C*sizeof(T) run, so a warp reads whole 128B lines with 16-byte vectors.
// Algorithmic bytes per call: (2*C + 2) * H*W * 4  (fp32), i.e. x read once + out written once + flow.
#include "common.cuh"
#include "launch.h"

namespace e2f {

struct Corner {
  int off[4];     // pixel offsets (in pixels) of the 4 corners, clamped to be addressable
  float wgt[4];   // bilinear weight, 0 for corners outside the image (zeros padding)
};

// (px, py): absolute sample position.  pad_mode 0 = zeros, 1 = border (clamp the coordinate first).
__device__ __forceinline__ Corner make_corners(float px, float py, int H, int W, int pad_mode) {
  if (pad_mode == 1) {
    px = fminf(fmaxf(px, 0.f), static_cast<float>(W - 1));
    py = fminf(fmaxf(py, 0.f), static_cast<float>(H - 1));
  }
  // keep the float->int conversion defined for wild flows
  px = fminf(fmaxf(px, -4.f), static_cast<float>(W) + 4.f);
  py = fminf(fmaxf(py, -4.f), static_cast<float>(H) + 4.f);
  const float fx = floorf(px), fy = floorf(py);
  const float lx = px - fx, ly = py - fy;
  const int x0 = static_cast<int>(fx), y0 = static_cast<int>(fy);
  Corner c;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int dy = k >> 1, dx = k & 1;
    const int yy = y0 + dy, xx = x0 + dx;
    const bool in = (yy >= 0) && (yy < H) && (xx >= 0) && (xx < W);
    const float w = (dy ? ly : 1.f - ly) * (dx ? lx : 1.f - lx);
    c.wgt[k] = in ? w : 0.f;
    c.off[k] = min(max(yy, 0), H - 1) * W + min(max(xx, 0), W - 1);
  }
  return c;
}

// One thread per (pixel, 16-byte channel vector).
template <typename T>
__global__ void __launch_bounds__(256) flow_warp_nhwc_kernel(const T* __restrict__ x, const float2* __restrict__ flow,
                                                             T* __restrict__ out, int N, int H, int W, int C,
                                                             int pad_mode) {
  constexpr int VEC = 16 / sizeof(T);
  const int vecs = C / VEC;
  const long long total = static_cast<long long>(N) * H * W * vecs;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int v = static_cast<int>(idx % vecs);
  const long long pix = idx / vecs;
  const int xw = static_cast<int>(pix % W);
  const int yh = static_cast<int>((pix / W) % H);
  const long long n = pix / (static_cast<long long>(W) * H);
  const float2 f = __ldg(flow + pix);
  const Corner c = make_corners(static_cast<float>(xw) + f.x, static_cast<float>(yh) + f.y, H, W, pad_mode);
  const T* base = x + n * H * W * C + v * VEC;
  uint4 raw[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) raw[k] = __ldg(reinterpret_cast<const uint4*>(base + static_cast<long long>(c.off[k]) * C));
  uint4 res;
  if constexpr (sizeof(T) == 4) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float* p = reinterpret_cast<const float*>(&raw[k]);
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[i] = fmaf(c.wgt[k], p[i], acc[i]);
    }
    res = *reinterpret_cast<uint4*>(acc);
  } else {
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const __half2* p = reinterpret_cast<const __half2*>(&raw[k]);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 t = __half22float2(p[i]);
        acc[2 * i] = fmaf(c.wgt[k], t.x, acc[2 * i]);
        acc[2 * i + 1] = fmaf(c.wgt[k], t.y, acc[2 * i + 1]);
      }
    }
    uint32_t* r = reinterpret_cast<uint32_t*>(&res);
#pragma unroll
    for (int i = 0; i < 4; ++i) r[i] = pack_half2(acc[2 * i], acc[2 * i + 1]);
  }
  *reinterpret_cast<uint4*>(out + pix * C + v * VEC) = res;
}

// NCHW fp32, any (small) C: one thread per pixel, channel loop.  Used for 2-channel flows and 3-channel images.
__global__ void __launch_bounds__(256) flow_warp_nchw_kernel(const float* __restrict__ x, const float2* __restrict__ flow,
                                                             float* __restrict__ out, int N, int C, int H, int W,
                                                             int pad_mode) {
  const long long total = static_cast<long long>(N) * H * W;
  const long long pix = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (pix >= total) return;
  const int xw = static_cast<int>(pix % W);
  const int yh = static_cast<int>((pix / W) % H);
  const long long n = pix / (static_cast<long long>(W) * H);
  const float2 f = __ldg(flow + pix);
  const Corner c = make_corners(static_cast<float>(xw) + f.x, static_cast<float>(yh) + f.y, H, W, pad_mode);
  const long long plane = static_cast<long long>(H) * W;
  const float* xb = x + n * C * plane;
  float* ob = out + n * C * plane + static_cast<long long>(yh) * W + xw;
  for (int ch = 0; ch < C; ++ch) {
    const float* p = xb + ch * plane;
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) acc = fmaf(c.wgt[k], __ldg(p + c.off[k]), acc);
    ob[ch * plane] = acc;
  }
}

int launch_flow_warp_nhwc(const void* x, const float* flow, void* out, int n, int h, int w, int c, int dtype,
                          int pad_mode, cudaStream_t stream) {
  const int vec = dtype == 1 ? 8 : 4;
  const long long total = static_cast<long long>(n) * h * w * (c / vec);
  const int threads = 256;
  const unsigned blocks = static_cast<unsigned>((total + threads - 1) / threads);
  if (blocks == 0) return 0;
  if (dtype == 1)
    flow_warp_nhwc_kernel<__half><<<blocks, threads, 0, stream>>>(static_cast<const __half*>(x),
                                                                   reinterpret_cast<const float2*>(flow),
                                                                   static_cast<__half*>(out), n, h, w, c, pad_mode);
  else
    flow_warp_nhwc_kernel<float><<<blocks, threads, 0, stream>>>(static_cast<const float*>(x),
                                                                  reinterpret_cast<const float2*>(flow),
                                                                  static_cast<float*>(out), n, h, w, c, pad_mode);
  count_launch();
  return static_cast<int>(cudaGetLastError());
}

int launch_flow_warp_nchw(const float* x, const float* flow, float* out, int n, int c, int h, int w, int pad_mode,
                          cudaStream_t stream) {
  const long long total = static_cast<long long>(n) * h * w;
  const int threads = 256;
  const unsigned blocks = static_cast<unsigned>((total + threads - 1) / threads);
  if (blocks == 0) return 0;
  flow_warp_nchw_kernel<<<blocks, threads, 0, stream>>>(x, reinterpret_cast<const float2*>(flow), out, n, c, h, w,
                                                        pad_mode);
  count_launch();
  return static_cast<int>(cudaGetLastError());
}

}  // namespace e2f
