// Video-level driver kernels (reference test.py:132-179): everything the reference does on the
// host / in eager torch around each InpaintGenerator.forward call, as byte-exact HBM-bound kernels on uint8 frames.
//
//  * video_prepare_clip : gather the window's frames by id, uint8 -> float / 255 * 2 - 1, multiply by (1 - mask),
//                         mirror-pad to the model's modulus (test.py:139,152-165: to_tensors, imgs * (1 - masks),
//                         cat(x, flip(x))[:h + h_pad] twice) -> masked_frames [t][3][Hp][Wp] fp32 in one pass.
//  * video_compose      : prediction -> ((p + 1) / 2) * 255 -> uint8 (truncation) -> hole composite with the original
//                         frame (test.py:167-174) -> img [n_local][H][W][3] uint8.
//  * video_blend        : comp = first ? img : comp * 0.5 + img * 0.5 in fp32 (test.py:175-179); exact, because every
//                         value is an integer / 2^k with k <= 3.
//  * video_finalize     : fp32 -> uint8 truncation (test.py:195).
// Integer / byte work: results are bit-identical to the reference's numpy code (IEEE fp32 division and rounding
// replicated with __fdiv_rn / separate multiplies, no FMA contraction).
#include <cstdint>
#include "launch.h"

namespace e2f {
namespace video {

// one thread per output pixel (x fastest -> coalesced fp32 plane stores); 3 channels per thread
__global__ void __launch_bounds__(256) prepare_clip_kernel(const uint8_t* __restrict__ frames,
                                                           const uint8_t* __restrict__ masks,
                                                           const int* __restrict__ ids, float* __restrict__ out, int t,
                                                           int h, int w, int hp, int wp) {
  const long long total = static_cast<long long>(t) * hp * wp;
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int x = static_cast<int>(i % wp);
  const long long r = i / wp;
  const int y = static_cast<int>(r % hp), k = static_cast<int>(r / hp);
  // cat([x, flip(x)])[:h + pad]: row h + j is row h - 1 - j
  const int sy = y < h ? y : 2 * h - 1 - y, sx = x < w ? x : 2 * w - 1 - x;
  const long long src = (static_cast<long long>(__ldg(ids + k)) * h + sy) * w + sx;
  const float keep = 1.0f - static_cast<float>(__ldg(masks + src) != 0);
  const uint8_t* px = frames + src * 3;
  const long long plane = static_cast<long long>(hp) * wp;
  float* o = out + static_cast<long long>(k) * 3 * plane + static_cast<long long>(y) * wp + x;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float v = __fsub_rn(__fmul_rn(__fdiv_rn(static_cast<float>(__ldg(px + c)), 255.0f), 2.0f), 1.0f);
    o[c * plane] = __fmul_rn(v, keep);
  }
}

__global__ void __launch_bounds__(256) compose_kernel(const float* __restrict__ pred, const uint8_t* __restrict__ frames,
                                                      const uint8_t* __restrict__ masks, const int* __restrict__ ids,
                                                      uint8_t* __restrict__ img, int n_local, int h, int w, int hp,
                                                      int wp) {
  const long long total = static_cast<long long>(n_local) * h * w;
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int x = static_cast<int>(i % w);
  const long long r = i / w;
  const int y = static_cast<int>(r % h), k = static_cast<int>(r / h);
  const long long src = (static_cast<long long>(__ldg(ids + k)) * h + y) * w + x;
  const bool hole = __ldg(masks + src) != 0;
  const long long plane = static_cast<long long>(hp) * wp;
  const float* p = pred + static_cast<long long>(k) * 3 * plane + static_cast<long long>(y) * wp + x;
  uint8_t* o = img + i * 3;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    uint8_t v;
    if (hole) {
      const float f = __fmul_rn(__fdiv_rn(__fadd_rn(__ldg(p + c * plane), 1.0f), 2.0f), 255.0f);
      v = static_cast<uint8_t>(static_cast<int>(f));          // numpy astype(uint8): truncation; f is in [0, 255]
    } else {
      v = __ldg(frames + src * 3 + c);
    }
    o[c] = v;
  }
}

// one thread per byte of img
__global__ void __launch_bounds__(256) blend_kernel(const uint8_t* __restrict__ img, const int* __restrict__ ids,
                                                    const int* __restrict__ first, float* __restrict__ comp,
                                                    int n_local, long long frame_elems) {
  const long long total = static_cast<long long>(n_local) * frame_elems;
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int k = static_cast<int>(i / frame_elems);
  const long long e = i - static_cast<long long>(k) * frame_elems;
  float* c = comp + static_cast<long long>(__ldg(ids + k)) * frame_elems + e;
  const float v = static_cast<float>(__ldg(img + i));
  *c = __ldg(first + k) ? v : __fadd_rn(__fmul_rn(*c, 0.5f), __fmul_rn(v, 0.5f));
}

__global__ void __launch_bounds__(256) finalize_kernel(const float* __restrict__ comp, uint8_t* __restrict__ out,
                                                       long long count) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < count) out[i] = static_cast<uint8_t>(static_cast<int>(__ldg(comp + i)));
}

static unsigned blocks_for(long long total) { return static_cast<unsigned>((total + 255) / 256); }

}  // namespace video

int launch_video_prepare_clip(const uint8_t* frames, const uint8_t* masks, const int* ids, float* out, int t, int h,
                              int w, int hp, int wp, cudaStream_t stream) {
  const long long total = static_cast<long long>(t) * hp * wp;
  if (total == 0) return 0;
  video::prepare_clip_kernel<<<video::blocks_for(total), 256, 0, stream>>>(frames, masks, ids, out, t, h, w, hp, wp);
  count_launch();
  return static_cast<int>(cudaGetLastError());
}

int launch_video_compose(const float* pred, const uint8_t* frames, const uint8_t* masks, const int* ids, uint8_t* img,
                         int n_local, int h, int w, int hp, int wp, cudaStream_t stream) {
  const long long total = static_cast<long long>(n_local) * h * w;
  if (total == 0) return 0;
  video::compose_kernel<<<video::blocks_for(total), 256, 0, stream>>>(pred, frames, masks, ids, img, n_local, h, w, hp, wp);
  count_launch();
  return static_cast<int>(cudaGetLastError());
}

int launch_video_blend(const uint8_t* img, const int* ids, const int* first, float* comp, int n_local,
                       long long frame_elems, cudaStream_t stream) {
  const long long total = static_cast<long long>(n_local) * frame_elems;
  if (total == 0) return 0;
  video::blend_kernel<<<video::blocks_for(total), 256, 0, stream>>>(img, ids, first, comp, n_local, frame_elems);
  count_launch();
  return static_cast<int>(cudaGetLastError());
}

int launch_video_finalize(const float* comp, uint8_t* out, long long count, cudaStream_t stream) {
  if (count == 0) return 0;
  video::finalize_kernel<<<video::blocks_for(count), 256, 0, stream>>>(comp, out, count);
  count_launch();
  return static_cast<int>(cudaGetLastError());
}

}  // namespace e2f
