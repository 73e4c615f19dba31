// 3x3 / stride 1 / pad 1 convolutions of the path (encoder e2fgvi.py:75-109, decoder :143-150, offset heads
// feat_prop.py:20-28, backbones :73-77) as an implicit GEMM on tcgen05 with fp32-level accuracy (bf16 3-term split,
// see gemm.cu):   out[n,y,x,co] = act( sum_{tap,src,c} X_src[n, y+r-1, x+s-1, c] * W[co, tap, src, c] + b[co] ) (+ res)
//
//  * im2col is done by TMA: activations are NHWC bf16 (hi, lo); one 4-D box {64 ch, 16 x, 8 y, 1 n} per (tap, source,
//    64-channel chunk) lands directly as a 128-row K-major SWIZZLE_128B operand tile; negative / overflowing
//    coordinates are zero-filled by the TMA unit == the conv's zero padding.  Nothing is materialised.
//  * multi-source K: the channel concatenations in front of these convs (torch.cat at e2fgvi.py:103-108,
//    feat_prop.py:36,125,131-136) are never built — each concatenated tensor is its own TMA source.
//  * groups (encoder convs with groups 2/4/8): a tile's N range lives inside one group and its K chunks start at the
//    group's channel offset of every source; chunks that spill past a group's slice hit zero weights.
//  * small-channel sources (cin = 4 / 8 / 16 / 32: SPyNet's 7x7 convs, the 3-channel stem): "window-packed" K.  The
//    source is stored row-gapped, [N][H][W + pad][cin] with `pad` zero pixels in front of every row (the gap doubles
//    as the previous row's right padding).  The tensor map's pixel dimension has a stride of ONE pixel (x conv
//    stride) but dimension 0 spans 64 elements = 64/cin consecutive pixels, so every TMA row is the sliding window
//    [x - pad + g*PX, +PX) x cin — a whole slice of the kernel row per 64-wide K chunk instead of one tap padded from
//    cin to 64 channels (7x fewer K chunks for SPyNet's first conv, 12x for the stem).  Weights for taps past the
//    kernel width are zero; what those positions read is finite data of the same buffer.
//  * epilogue: + bias, LeakyReLU(slope), optional residual add, fp32 NHWC store and/or the bf16 (hi, lo) split of
//    the result — dense NHWC or row-gapped for a following window-packed conv (the zero gaps are written here).
// Pipeline = gemm.cu: persistent CTAs, TMA warp / MMA warp / 4 epilogue warps, double-buffered TMEM accumulator.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cstdlib>
#include "common.cuh"
#include "launch.h"

namespace e2f {
namespace conv {

constexpr int BM = 128, BK = 64;
constexpr int TILE_H = 8, TILE_W = 16;                  // 8 x 16 output pixels = 128 GEMM rows
constexpr int A_TILE = BM * BK * 2;
constexpr int EPI_WARPS = 8;                             // two warps per TMEM lane quarter, half of the tile's columns each
constexpr int MAX_COUT = 512;                            // bias staged in shared memory once per CTA
constexpr int EPI_STAGE = 2048;                          // per epilogue warp: 32 rows x 64 bytes store-transposition buffer
constexpr int THREADS = (2 + EPI_WARPS) * 32;
constexpr int MAX_SRC = 4;

// N tile: 128 output channels; 64 / 32 for the layers with <= 64 / <= 32 output channels per group (decoder, encoder
// conv 1, SPyNet, the 3-channel output conv), which would otherwise waste most of every MMA; 96 for the encoder's
// groups-of-96 conv (e2fgvi.py:86: 768 -> 384, groups 4), where a 128-wide tile would compute 25 % padding.
template <int BN>
struct Cfg {
  static constexpr int W_TILE = BN * BK * 2;
  static constexpr int STAGE = 2 * A_TILE + 2 * W_TILE;   // 64 KB (BN=128) / 48 KB (BN=64) / 40 KB (BN=32)
  static constexpr int STAGES = (BN >= 96) ? 3 : 4;
  static constexpr int ACC_COLS = 2 * BN;                 // accumulator: [Ah.Wh + Al.Wh | Ah.Wl], summed by the epilogue
  // double-buffered; tcgen05.alloc wants a power of two (BN = 96: 384 columns used of 512)
  static constexpr int TMEM_COLS = (4 * BN <= 128) ? 128 : (4 * BN <= 256) ? 256 : 512;
  static constexpr int SMEM = STAGES * STAGE + 256 + MAX_COUT * 4 + EPI_WARPS * EPI_STAGE + 1024;
};

struct Maps {
  CUtensorMap a_hi[MAX_SRC], a_lo[MAX_SRC], w_hi, w_lo;
};

struct Params {
  int N, H, W, Cout, groups;     // H, W: OUTPUT spatial size
  int ks, stride, pad;           // square kernel size (3 or 7), stride (1 or 2), zero padding
  int nsrc;
  int rows_px;             // 0: dense NHWC sources.  > 0: window-packed K, 64 / cin pixels per K chunk (one source)
  int rows_g;              // ... K chunks per kernel row = ceil(ks / rows_px)
  int out_lead, out_pitch; // split output rows: `out_lead` zero pixels, then W pixels; out_pitch = W + out_lead
  int out_tail;            // zero pixels after the last row of the split output (0 when dense)
  int cig[MAX_SRC];        // channels per group of each source
  int chunks[MAX_SRC];     // ceil(cig / 64)
  int chunks_total;        // sum of chunks
  // Tap table ("gather conv"): tap i reads the input at (y*stride + tap_dy[i], x*stride + tap_dx[i]); a plain k x k conv
  // lists its k*k taps with dy = ky - pad, dx = kx - pad.  Taps are grouped into PHASES: phase ph owns taps
  // [ph_tap0[ph], ph_tap0[ph+1]) and writes output pixel (y*ostep + ph_oy[ph], x*ostep + ph_ox[ph]) of an
  // out_H x out_W image — one phase for a conv; nine for the transposed 7x7 / stride-3 conv that SoftComp's
  // Linear + fold is (tfocal_transformer.py:65-72): output pixels with the same (y mod 3, x mod 3) share a tap set.
  int tile_w, tile_h;      // GEMM-grid pixels per tile (tile_w * tile_h <= 128; rows beyond are idle)
  int nphase, ostep, out_H, out_W;
  long long out_nstride;   // pixels between consecutive images of the fp32 output / residual (dense: out_H * out_W)
  long long osp_nstride;   // ... of the split outputs (dense: out_H * out_pitch)
  int8_t tap_dy[64], tap_dx[64];
  uint8_t ph_tap0[10], ph_oy[9], ph_ox[9];
  float slope;             // LeakyReLU negative slope (1 = identity)
  int epi_flags;           // EPI_TANH: tanh after the activation; EPI_NCHW: fp32 output stored [N][Cout][out_H][out_W]
  const float* bias;
  const float* bias_map;   // optional fp32 [out_H][out_W][Cout] added per output pixel (folded Linear bias + sc.bias) or null
  const float* residual;   // NHWC fp32 [N][out_H][out_W][Cout] or null
  float* out;              // NHWC fp32 or null
  __nv_bfloat16* out_hi;   // NHWC bf16 split of the result (operand format of the next conv) or null
  __nv_bfloat16* out_lo;
};

constexpr int EPI_TANH = 1, EPI_NCHW = 2;     // both only on the element-wise store path (Cout % 4 != 0: the 3-channel output conv)

__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(N >> 3) << 17) |
         (static_cast<uint32_t>(M >> 4) << 24);
}

__device__ __forceinline__ void tma_load_4d(uint32_t smem_dst, const void* tmap, uint64_t* bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

struct TileCoord {
  int n, y0, x0, g, co0;   // co0: first output channel of the tile (global index)
  int ph, oy, ox;          // phase and its output-pixel offset
};

template <int BN>
__device__ __forceinline__ TileCoord decode_tile(int tile, const Params& p, int tiles_y, int tiles_x, int tiles_ng) {
  TileCoord t;
  const int nt = tile % tiles_ng;
  int r = tile / tiles_ng;
  t.g = r % p.groups;
  r /= p.groups;
  t.ph = r % p.nphase;
  r /= p.nphase;
  const int tx = r % tiles_x;
  r /= tiles_x;
  const int ty = r % tiles_y;
  t.n = r / tiles_y;
  t.y0 = ty * p.tile_h;
  t.x0 = tx * p.tile_w;
  t.co0 = t.g * (p.Cout / p.groups) + nt * BN;
  t.oy = p.ph_oy[t.ph];
  t.ox = p.ph_ox[t.ph];
  return t;
}

// Epilogue of one 128-pixel x BN-channel tile for the calling thread's accumulator row `r` (TMEM lane): sums the two
// accumulator halves, + bias, LeakyReLU, optional residual, fp32 and/or bf16-split NHWC stores.  TW = tile width in
// pixels (row r is pixel (y0 + r / TW, x0 + r % TW)); taddr = TMEM address of the row's first accumulator column.
// The calling warp handles the 32-column chunks [c_begin, c_end) of the tile.  bias_s = the layer's bias in SHARED
// memory (zeros when the layer has none): the per-channel __ldg's this replaces queued behind the epilogue's own
// 32-line-per-instruction stores and made up 59 % of the epilogue warps' stall samples, which bounded every layer
// whose main loop is shorter than ~10k cycles per tile (profiles/r01/ncu_conv_epilogue_r01.txt).
//
// Stores are STAGED through `stage` (2 KB of shared memory per epilogue warp): with thread = pixel, a direct 16-byte store
// per thread hits 32 different 128-byte lines per instruction (pixels are Cout * 2 or 4 bytes apart), ~125 cycles each,
// which made the epilogue the bottleneck of the layers with a short main loop.  Each warp instead transposes its
// 32 rows x 64 bytes through a conflict-free XOR-swizzled buffer so that 4 consecutive lanes write one pixel's 64
// contiguous bytes (8 lines per instruction).  Chunks that are not 32 full, 16-byte-aligned channels take the direct path.
template <int BN>
__device__ __forceinline__ void epilogue_tile(const Params& p, const TileCoord& t, uint32_t taddr, int r, int cog,
                                              const float* __restrict__ bias_s, int c_begin, int c_end,
                                              uint8_t* __restrict__ stage, const int TW, const int TH) {
  const int lane = r & 31;
  // accumulator row R -> GEMM-grid pixel (gy, gx) -> output pixel (Y, X) of the out_H x out_W image
  auto map_row = [&](int R, int& Y, int& X) -> bool {
    const int ly = R / TW, gy = t.y0 + ly, gx = t.x0 + (R - ly * TW);
    Y = gy * p.ostep + t.oy;
    X = gx * p.ostep + t.ox;
    return ly < TH && gy < p.H && gx < p.W && Y < p.out_H && X < p.out_W;
  };
  int y, x;                                            // this thread's OUTPUT pixel
  const bool pix_ok = map_row(r, y, x);
  const size_t pix = static_cast<size_t>(t.n) * p.out_nstride + static_cast<size_t>(y) * p.out_W + x;        // fp32 out, residual
  const size_t opix = static_cast<size_t>(t.n) * p.osp_nstride + static_cast<size_t>(y) * p.out_pitch + p.out_lead + x;   // split outputs
  const size_t mpix = static_cast<size_t>(y) * p.out_W + x;                                         // bias map
  const int co_end = (t.g + 1) * cog;               // exclusive end of this group's output channels
  const bool vec_ok = (p.Cout & 3) == 0;            // 16-byte aligned channel groups
#pragma unroll 1
  for (int c = c_begin; c < c_end; ++c) {
    uint32_t v[32], v2[32];
    tmem_ld32(taddr + c * 32, v);
    tmem_ld32(taddr + BN + c * 32, v2);              // the Ah.Wl half of the accumulator
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) + __uint_as_float(v2[i]));
    const int co = t.co0 + c * 32;
    if (vec_ok && co + 32 <= co_end) {
      // ------------------------------------------------------------ staged, coalesced stores (warp-uniform branch)
      float f[32];
#pragma unroll
      for (int g4 = 0; g4 < 8; ++g4) {
        const float4 b = *reinterpret_cast<const float4*>(bias_s + co + g4 * 4);
        const float bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float a = __uint_as_float(v[g4 * 4 + i]) + bb[i];
          f[g4 * 4 + i] = a > 0.f ? a : a * p.slope;
        }
      }
      if (p.residual && pix_ok) {
        const float4* r4 = reinterpret_cast<const float4*>(p.residual + pix * p.Cout + co);
#pragma unroll
        for (int g4 = 0; g4 < 8; ++g4) {
          const float4 ra = __ldg(r4 + g4);
          f[g4 * 4] += ra.x; f[g4 * 4 + 1] += ra.y; f[g4 * 4 + 2] += ra.z; f[g4 * 4 + 3] += ra.w;
        }
      }
      if (p.bias_map && pix_ok) {
        const float4* m4 = reinterpret_cast<const float4*>(p.bias_map + mpix * p.Cout + co);
#pragma unroll
        for (int g4 = 0; g4 < 8; ++g4) {
          const float4 ra = __ldg(m4 + g4);
          f[g4 * 4] += ra.x; f[g4 * 4 + 1] += ra.y; f[g4 * 4 + 2] += ra.z; f[g4 * 4 + 3] += ra.w;
        }
      }
      // write side: row = lane, logical 16-byte chunk cc at physical chunk cc ^ ((lane >> 1) & 3); read side: lanes
      // 4k..4k+3 fetch the 4 chunks of row j*8 + k.  Both sides touch 8 distinct bank groups per quarter-warp.
      const uint32_t wbase = smem_u32(stage) + lane * 64, wsw = (lane >> 1) & 3;
      const int sub = lane & 3, prow = lane >> 2, rbase = r - lane;
      auto st_chunk = [&](int cc, uint32_t a, uint32_t b, uint32_t c2, uint32_t d) {
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(wbase + ((cc ^ wsw) << 4)), "r"(a), "r"(b), "r"(c2), "r"(d)
                     : "memory");
      };
      auto ld_chunk = [&](int rr) {
        uint4 u;
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w)
                     : "r"(smem_u32(stage) + rr * 64 + ((sub ^ ((rr >> 1) & 3)) << 4))
                     : "memory");
        return u;
      };
      if (p.out) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {                 // 16 fp32 channels = 64 bytes per row and pass
#pragma unroll
          for (int cc = 0; cc < 4; ++cc)
            st_chunk(cc, __float_as_uint(f[h * 16 + cc * 4]), __float_as_uint(f[h * 16 + cc * 4 + 1]),
                     __float_as_uint(f[h * 16 + cc * 4 + 2]), __float_as_uint(f[h * 16 + cc * 4 + 3]));
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int rr = j * 8 + prow, R = rbase + rr;
            int yy, xx;
            const bool ok = map_row(R, yy, xx);
            const uint4 u = ld_chunk(rr);
            if (ok)
              *reinterpret_cast<uint4*>(p.out + (static_cast<size_t>(t.n) * p.out_nstride + static_cast<size_t>(yy) * p.out_W + xx) * p.Cout +
                                        co + h * 16 + sub * 4) = u;
          }
          __syncwarp();
        }
      }
      if (p.out_hi) {
        uint32_t hp[16], lp[16];                      // packed bf16 pairs
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const __nv_bfloat162 hb = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
          const float2 hf = __bfloat1622float2(hb);
          const __nv_bfloat162 lb = __floats2bfloat162_rn(f[2 * i] - hf.x, f[2 * i + 1] - hf.y);
          hp[i] = *reinterpret_cast<const uint32_t*>(&hb);
          lp[i] = *reinterpret_cast<const uint32_t*>(&lb);
        }
#pragma unroll
        for (int part = 0; part < 2; ++part) {        // 32 bf16 channels = 64 bytes per row: hi, then lo
          const uint32_t* src = part ? lp : hp;
          __nv_bfloat16* dst = part ? p.out_lo : p.out_hi;
#pragma unroll
          for (int cc = 0; cc < 4; ++cc) st_chunk(cc, src[cc * 4], src[cc * 4 + 1], src[cc * 4 + 2], src[cc * 4 + 3]);
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int rr = j * 8 + prow, R = rbase + rr;
            int yy, xx;
            const bool ok = map_row(R, yy, xx);
            const uint4 u = ld_chunk(rr);
            if (ok)
              *reinterpret_cast<uint4*>(dst + (static_cast<size_t>(t.n) * p.osp_nstride + static_cast<size_t>(yy) * p.out_pitch + p.out_lead + xx) *
                                              p.Cout + co + sub * 8) = u;
          }
          __syncwarp();
        }
      }
    } else if (pix_ok && co < co_end) {
#pragma unroll
      for (int g8 = 0; g8 < 4; ++g8) {              // 8 output channels at a time
        const int cb = co + g8 * 8;
        if (vec_ok && cb + 8 <= co_end) {
          float f[8];
          const float4 b0 = *reinterpret_cast<const float4*>(bias_s + cb), b1 = *reinterpret_cast<const float4*>(bias_s + cb + 4);
          const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float a = __uint_as_float(v[g8 * 8 + i]) + bb[i];
            f[i] = a > 0.f ? a : a * p.slope;
          }
          if (p.residual) {
            const float4* r4 = reinterpret_cast<const float4*>(p.residual + pix * p.Cout + cb);
            const float4 ra = __ldg(r4), rb = __ldg(r4 + 1);
            f[0] += ra.x; f[1] += ra.y; f[2] += ra.z; f[3] += ra.w;
            f[4] += rb.x; f[5] += rb.y; f[6] += rb.z; f[7] += rb.w;
          }
          if (p.bias_map) {
            const float4* r4 = reinterpret_cast<const float4*>(p.bias_map + mpix * p.Cout + cb);
            const float4 ra = __ldg(r4), rb = __ldg(r4 + 1);
            f[0] += ra.x; f[1] += ra.y; f[2] += ra.z; f[3] += ra.w;
            f[4] += rb.x; f[5] += rb.y; f[6] += rb.z; f[7] += rb.w;
          }
          if (p.out) {
            float4* d4 = reinterpret_cast<float4*>(p.out + pix * p.Cout + cb);
            d4[0] = make_float4(f[0], f[1], f[2], f[3]);
            d4[1] = make_float4(f[4], f[5], f[6], f[7]);
          }
          if (p.out_hi) {
            uint32_t hp[4], lp[4];                  // packed bf16 pairs, kept in registers
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const __nv_bfloat162 hb = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
              const float2 hf = __bfloat1622float2(hb);
              const __nv_bfloat162 lb = __floats2bfloat162_rn(f[2 * i] - hf.x, f[2 * i + 1] - hf.y);
              hp[i] = *reinterpret_cast<const uint32_t*>(&hb);
              lp[i] = *reinterpret_cast<const uint32_t*>(&lb);
            }
            *reinterpret_cast<uint4*>(p.out_hi + opix * p.Cout + cb) = make_uint4(hp[0], hp[1], hp[2], hp[3]);
            *reinterpret_cast<uint4*>(p.out_lo + opix * p.Cout + cb) = make_uint4(lp[0], lp[1], lp[2], lp[3]);
          }
        } else if (cb < co_end) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            if (cb + i < co_end) {
              float a = __uint_as_float(v[g8 * 8 + i]) + bias_s[cb + i];
              a = a > 0.f ? a : a * p.slope;
              if (p.residual) a += __ldg(p.residual + pix * p.Cout + cb + i);
              if (p.bias_map) a += __ldg(p.bias_map + mpix * p.Cout + cb + i);
              if (p.epi_flags & EPI_TANH) a = tanhf(a);
              if (p.out) {
                if (p.epi_flags & EPI_NCHW)   // thread = pixel: consecutive lanes write consecutive x of one channel plane
                  p.out[((static_cast<size_t>(t.n) * p.Cout + cb + i) * p.out_H + y) * p.out_W + x] = a;
                else
                  p.out[pix * p.Cout + cb + i] = a;
              }
              if (p.out_hi) {
                const __nv_bfloat16 hb = __float2bfloat16_rn(a);
                p.out_hi[opix * p.Cout + cb + i] = hb;
                p.out_lo[opix * p.Cout + cb + i] = __float2bfloat16_rn(a - __bfloat162float(hb));
              }
            }
          }
        }
      }
    }
  }
  // row-gapped split output: the pixel at x == 0 also writes the zero gap in front of its row, the very last pixel
  // the zero tail (once per pixel: only the first N tile of group 0 does it; Cout % 8 == 0 is checked by the API)
  if (p.out_hi && p.out_lead && pix_ok && t.co0 == 0 && c_begin == 0) {
    const uint4 z = make_uint4(0, 0, 0, 0);
    if (x == 0) {
      uint4* zh = reinterpret_cast<uint4*>(p.out_hi + (opix - p.out_lead) * p.Cout);
      uint4* zl = reinterpret_cast<uint4*>(p.out_lo + (opix - p.out_lead) * p.Cout);
      for (int i = 0; i < p.out_lead * p.Cout / 8; ++i) zh[i] = zl[i] = z;
    }
    if (x == p.out_W - 1 && y == p.out_H - 1 && t.n == p.N - 1) {
      uint4* zh = reinterpret_cast<uint4*>(p.out_hi + (opix + 1) * p.Cout);
      uint4* zl = reinterpret_cast<uint4*>(p.out_lo + (opix + 1) * p.Cout);
      for (int i = 0; i < p.out_tail * p.Cout / 8; ++i) zh[i] = zl[i] = z;
    }
  }
}

template <int BN>
__global__ void __launch_bounds__(THREADS, 1) conv3x3_kernel(const __grid_constant__ Maps maps, const __grid_constant__ Params p) {
  constexpr int W_TILE = Cfg<BN>::W_TILE, STAGE = Cfg<BN>::STAGE, STAGES = Cfg<BN>::STAGES, TMEM_COLS = Cfg<BN>::TMEM_COLS;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE);
  uint64_t* empty = full + STAGES;
  uint64_t* acc_full = empty + STAGES;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  float* bias_s = reinterpret_cast<float*>(smem + STAGES * STAGE + 256);
  uint8_t* epi_stage = smem + STAGES * STAGE + 256 + MAX_COUT * 4;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int epi_warps = (blockDim.x >> 5) - 2;                // 8, or 4 when the launch has no room for 8 staging buffers
  for (int i = tid; i < MAX_COUT; i += blockDim.x) bias_s[i] = (p.bias && i < p.Cout) ? __ldg(p.bias + i) : 0.f;
  const int tiles_y = (p.H + p.tile_h - 1) / p.tile_h, tiles_x = (p.W + p.tile_w - 1) / p.tile_w;
  const int cog = p.Cout / p.groups;
  const int tiles_ng = (cog + BN - 1) / BN;
  const int num_tiles = p.N * tiles_y * tiles_x * p.nphase * p.groups * tiles_ng;

  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], epi_warps);
    }
    fence_barrier_init();
    tma_prefetch_desc(&maps.w_hi);
    tma_prefetch_desc(&maps.w_lo);
    for (int i = 0; i < p.nsrc; ++i) {
      tma_prefetch_desc(&maps.a_hi[i]);
      tma_prefetch_desc(&maps.a_lo[i]);
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tbase = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (im2col by coordinates)
    if (elect_one()) {   // one thread, chosen by elect.sync: ptxas then emits bare UTCHMMA / UTMALDG (no per-instruction ELECT loop)
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const TileCoord t = decode_tile<BN>(tile, p, tiles_y, tiles_x, tiles_ng);
        int kb = 0;
        // bytes a stage receives: two A boxes of tile_w * tile_h rows x 128 B (rows beyond stay idle) + [Wh | Wl]
        const uint32_t stage_tx = 2u * static_cast<uint32_t>(p.tile_w * p.tile_h) * 128u + 2u * W_TILE;
        if (p.rows_px) {
          // window-packed K: chunk (ky, g) = pixels [x*stride - pad + g*PX, +PX) x cin of input row y*stride - pad + ky
          for (int ky = 0; ky < p.ks; ++ky) {
            const int yy = t.y0 * p.stride - p.pad + ky;
            for (int g = 0; g < p.rows_g; ++g, ++kb, ++it) {
              const int stage = it % STAGES;
              mbar_wait(&empty[stage], ((it / STAGES) & 1) ^ 1);
              mbar_arrive_expect_tx(&full[stage], stage_tx);
              const uint32_t s0 = smem_u32(smem + stage * STAGE);
              const int xi = t.x0 + g * p.rows_px / p.stride;       // window-start index (row gap = left padding)
              tma_load_4d(s0, &maps.a_hi[0], &full[stage], 0, xi, yy, t.n);
              tma_load_4d(s0 + A_TILE, &maps.a_lo[0], &full[stage], 0, xi, yy, t.n);
              tma_load_2d(s0 + 2 * A_TILE, &maps.w_hi, &full[stage], kb * BK, t.co0);
              tma_load_2d(s0 + 2 * A_TILE + W_TILE, &maps.w_lo, &full[stage], kb * BK, t.co0);
            }
          }
          continue;
        }
        const int tap0 = p.ph_tap0[t.ph], tap1 = p.ph_tap0[t.ph + 1];
        kb = tap0 * p.chunks_total;                      // K blocks of the packed weight are ordered by tap
        for (int tap = tap0; tap < tap1; ++tap) {
          // input coordinate of the tile's first output pixel for this tap (TMA steps by `stride` inside the box)
          const int yy = t.y0 * p.stride + p.tap_dy[tap], xx = t.x0 * p.stride + p.tap_dx[tap];
          for (int s = 0; s < p.nsrc; ++s) {
            const int c_base = t.g * p.cig[s];
            for (int j = 0; j < p.chunks[s]; ++j, ++kb, ++it) {
              const int stage = it % STAGES;
              mbar_wait(&empty[stage], ((it / STAGES) & 1) ^ 1);
              mbar_arrive_expect_tx(&full[stage], stage_tx);
              const uint32_t s0 = smem_u32(smem + stage * STAGE);
              tma_load_4d(s0, &maps.a_hi[s], &full[stage], c_base + j * BK, xx, yy, t.n);
              tma_load_4d(s0 + A_TILE, &maps.a_lo[s], &full[stage], c_base + j * BK, xx, yy, t.n);
              tma_load_2d(s0 + 2 * A_TILE, &maps.w_hi, &full[stage], kb * BK, t.co0);
              tma_load_2d(s0 + 2 * A_TILE + W_TILE, &maps.w_lo, &full[stage], kb * BK, t.co0);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (elect_one()) {   // one thread, chosen by elect.sync: ptxas then emits bare UTCHMMA / UTMALDG (no per-instruction ELECT loop)
      // The Wh and Wl tiles are adjacent in a stage, so ONE N = 2*BN instruction multiplies Ah with both
      // (accumulator columns [0,BN) and [BN,2BN)); a second, N = BN wide, adds Al.Wh to the first half.  Ah is read from
      // shared memory once instead of twice: 3 -> 2 instructions per K step, ~20% less operand traffic on the
      // shared-memory port that the TMA writes share.
      const uint32_t idesc = idesc_bf16(BM, BN), idesc2 = idesc_bf16(BM, 2 * BN);
      const uint64_t d_ah0 = umma_desc_sw128(smem_u32(smem), 16, 1024);
      const uint64_t d_al0 = umma_desc_adv(d_ah0, A_TILE), d_wh0 = umma_desc_adv(d_ah0, 2 * A_TILE);
      uint32_t it = 0, local = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++local) {
        const int buf = local & 1;
        int num_kb = p.ks * p.rows_g;                     // window-packed K
        if (!p.rows_px) {
          const int ph = (tile / (tiles_ng * p.groups)) % p.nphase;
          num_kb = (p.ph_tap0[ph + 1] - p.ph_tap0[ph]) * p.chunks_total;
        }
        mbar_wait(&acc_empty[buf], ((local >> 1) & 1) ^ 1);
        tc_fence_after_sync();
        const uint32_t d = tbase + buf * Cfg<BN>::ACC_COLS;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int stage = it % STAGES;
          mbar_wait(&full[stage], (it / STAGES) & 1);
          tc_fence_after_sync();
          const uint32_t soff = (stage * STAGE) >> 4;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t dah = d_ah0 + soff + 2 * k, dal = d_al0 + soff + 2 * k;
            const uint64_t dwh = d_wh0 + soff + 2 * k;
            umma_f16(d, dah, dwh, idesc2, (kb | k) != 0);   // [Ah.Wh | Ah.Wl]
            umma_f16(d, dal, dwh, idesc, 1);                // + Al.Wh
          }
          umma_commit(&empty[stage]);
        }
        umma_commit(&acc_full[buf]);
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue
    const int q = warp & 3;                                   // TMEM lane quarter this warp may read
    constexpr int NCH = BN / 32;                              // 32-column chunks per tile, split between the two warps of a quarter
    const int c_split = epi_warps > 4 ? (NCH + 1) / 2 : NCH;
    const int c_begin = (warp - 2) < 4 ? 0 : c_split, c_end = (warp - 2) < 4 ? c_split : NCH;
    uint8_t* my_stage = epi_stage + (warp - 2) * EPI_STAGE;
    uint32_t local = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++local) {
      const int buf = local & 1;
      const TileCoord t = decode_tile<BN>(tile, p, tiles_y, tiles_x, tiles_ng);
      mbar_wait(&acc_full[buf], (local >> 1) & 1);
      tc_fence_after_sync();
      epilogue_tile<BN>(p, t, tbase + (static_cast<uint32_t>(q * 32) << 16) + buf * Cfg<BN>::ACC_COLS, q * 32 + lane, cog,
                        bias_s, c_begin, c_end, my_stage, p.tile_w, p.tile_h);
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tbase, TMEM_COLS);
}

// ---------------------------------------------------------------------------------------------------------------------
// HALO variant for the low-Cout, high-resolution 3x3 / stride 1 / pad 1 layers (encoder conv 2, decoder deconv-2 conv, the
// 3-channel output conv: Cout <= 64, groups 1).  There the generic kernel is bound by the L2 -> SM fabric (~6300 B/clk
// chip-wide): every 64-channel chunk of a tile is fetched nine times, once per tap, and the matching weight tile comes
// along every time — 432 KB per 128 x 64 output tile against 3456 MMA cycles.  Here
//  * the tile is 8 x 16 pixels and ONE TMA box {64 ch, 10 x, 18 y} per (chunk, hi | lo) brings its whole halo
//    (23 KB instead of 9 x 16 KB); tap (dy, dx) reads it in place through a UMMA descriptor whose start address is shifted
//    by (dy*10 + dx) pixels of 128 B and whose 8-row-group stride (SBO) is one halo row, 1280 B.  The 128B swizzle is a
//    function of absolute shared-memory address bits, so the shifted start needs no base offset (tools/halo_probe.cu,
//    profiles/r01/halo_probe.log);
//  * all weight tiles of the layer (9 taps x chunks x [Wh | Wl]) are loaded ONCE per persistent CTA and stay in shared
//    memory (144 KB for 64 -> 64), so steady-state L2 traffic is the 46 KB halo pair per chunk;
//  * the halo ring holds single (hi or lo) pieces: all Ah MMAs of a chunk (9 taps x 4, N = 2*BN against [Wh | Wl]) are
//    issued from one piece, then all Al MMAs (N = BN against Wh) from the next.
constexpr int HTILE_W = 8, HTILE_H = 16, HALO_W = HTILE_W + 2, HALO_H = HTILE_H + 2;
constexpr int HALO_BYTES = HALO_W * HALO_H * BK * 2;     // 23040
constexpr int HALO_SLOT = 23552;                        // rounded up to the 1024-byte swizzle atom
constexpr int HALO_MAX_SLOTS = 6;
constexpr int SMEM_LIMIT = 232448;                      // 227 KB opt-in maximum per CTA

__host__ __device__ constexpr int halo_w_bytes(int bn, int chunks_total) { return 9 * chunks_total * 2 * bn * BK * 2; }

template <int BN>
__global__ void __launch_bounds__(THREADS, 1) conv3x3_halo_kernel(const __grid_constant__ Maps maps, const __grid_constant__ Params p,
                                                                  const int nslots) {
  constexpr int W_TILE = Cfg<BN>::W_TILE, TMEM_COLS = Cfg<BN>::TMEM_COLS;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int w_bytes = halo_w_bytes(BN, p.chunks_total);
  uint8_t* ring = smem + w_bytes;                                   // w_bytes is a multiple of 1024
  uint64_t* a_full = reinterpret_cast<uint64_t*>(ring + nslots * HALO_SLOT);
  uint64_t* a_empty = a_full + HALO_MAX_SLOTS;
  uint64_t* acc_full = a_empty + HALO_MAX_SLOTS;
  uint64_t* acc_empty = acc_full + 2;
  uint64_t* w_full = acc_empty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_full + 1);
  float* bias_s = reinterpret_cast<float*>(ring + nslots * HALO_SLOT + 256);
  uint8_t* epi_stage = ring + nslots * HALO_SLOT + 256 + MAX_COUT * 4;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int epi_warps = (blockDim.x >> 5) - 2;                // 8, or 4 when the launch has no room for 8 staging buffers
  for (int i = tid; i < MAX_COUT; i += blockDim.x) bias_s[i] = (p.bias && i < p.Cout) ? __ldg(p.bias + i) : 0.f;
  const int tiles_y = (p.H + HTILE_H - 1) / HTILE_H, tiles_x = (p.W + HTILE_W - 1) / HTILE_W;
  const int num_tiles = p.N * tiles_y * tiles_x;
  const int pieces = 2 * p.chunks_total;                            // (chunk, hi | lo) halo pieces per tile

  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  if (tid == 0) {
    for (int s = 0; s < nslots; ++s) {
      mbar_init(&a_full[s], 1);
      mbar_init(&a_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], epi_warps);
    }
    mbar_init(w_full, 1);
    fence_barrier_init();
    tma_prefetch_desc(&maps.w_hi);
    tma_prefetch_desc(&maps.w_lo);
    for (int i = 0; i < p.nsrc; ++i) {
      tma_prefetch_desc(&maps.a_hi[i]);
      tma_prefetch_desc(&maps.a_lo[i]);
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tbase = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {   // one thread, chosen by elect.sync: ptxas then emits bare UTCHMMA / UTMALDG (no per-instruction ELECT loop)
      // resident weights: slot (tap, chunk) = [Wh (BN rows) | Wl (BN rows)], K column of the packed weight = (tap, chunk)
      mbar_arrive_expect_tx(w_full, static_cast<uint32_t>(w_bytes));
      for (int tap = 0; tap < 9; ++tap)
        for (int c = 0; c < p.chunks_total; ++c) {
          const int kb = tap * p.chunks_total + c;
          const uint32_t dst = smem_u32(smem) + kb * 2 * W_TILE;
          tma_load_2d(dst, &maps.w_hi, w_full, kb * BK, 0);
          tma_load_2d(dst + W_TILE, &maps.w_lo, w_full, kb * BK, 0);
        }
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int tx = tile % tiles_x, ty = (tile / tiles_x) % tiles_y, n = tile / (tiles_x * tiles_y);
        const int xx = tx * HTILE_W - 1, yy = ty * HTILE_H - 1;
        for (int s = 0; s < p.nsrc; ++s)
          for (int j = 0; j < p.chunks[s]; ++j)
            for (int part = 0; part < 2; ++part, ++it) {
              const int slot = it % nslots;
              mbar_wait(&a_empty[slot], ((it / nslots) & 1) ^ 1);
              mbar_arrive_expect_tx(&a_full[slot], HALO_BYTES);
              tma_load_4d(smem_u32(ring + slot * HALO_SLOT), part ? &maps.a_lo[s] : &maps.a_hi[s], &a_full[slot], j * BK,
                          xx, yy, n);
            }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (elect_one()) {   // one thread, chosen by elect.sync: ptxas then emits bare UTCHMMA / UTMALDG (no per-instruction ELECT loop)
      const uint32_t idesc = idesc_bf16(BM, BN), idesc2 = idesc_bf16(BM, 2 * BN);
      const uint64_t d_a0 = umma_desc_sw128(smem_u32(ring), 16, HALO_W * 128);
      const uint64_t d_w0 = umma_desc_sw128(smem_u32(smem), 16, 1024);
      mbar_wait(w_full, 0);
      uint32_t it = 0, local = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++local) {
        const int buf = local & 1;
        mbar_wait(&acc_empty[buf], ((local >> 1) & 1) ^ 1);
        tc_fence_after_sync();
        const uint32_t d = tbase + buf * Cfg<BN>::ACC_COLS;
        for (int pc = 0; pc < pieces; ++pc, ++it) {
          const int slot = it % nslots;
          const int chunk = pc >> 1, lo = pc & 1;
          mbar_wait(&a_full[slot], (it / nslots) & 1);
          tc_fence_after_sync();
          // Issue loop kept as tight as possible — the layers served here are bound by the issuing thread, not by the
          // tensor pipe (N <= 128: the pipe needs <= 64 cycles per MMA): taps fully unrolled so that the halo shift
          // (dy*10 + dx pixels) is an immediate, one running 64-bit add per weight slot, hi / lo pieces in separate loops.
          const uint64_t da = d_a0 + ((slot * HALO_SLOT) >> 4);
          uint64_t dw = d_w0 + ((chunk * 2 * W_TILE) >> 4);
          const uint32_t wstep = static_cast<uint32_t>(p.chunks_total * 2 * W_TILE) >> 4;
          if (lo) {
#pragma unroll
            for (int tap = 0; tap < 9; ++tap, dw += wstep) {
              const uint64_t dat = da + ((((tap / 3) * HALO_W + tap % 3) * 128) >> 4);
#pragma unroll
              for (int k = 0; k < BK / 16; ++k) umma_f16(d, dat + 2 * k, dw + 2 * k, idesc, 1);                 // + Al.Wh
            }
          } else {
#pragma unroll
            for (int tap = 0; tap < 9; ++tap, dw += wstep) {
              const uint64_t dat = da + ((((tap / 3) * HALO_W + tap % 3) * 128) >> 4);
#pragma unroll
              for (int k = 0; k < BK / 16; ++k) umma_f16(d, dat + 2 * k, dw + 2 * k, idesc2, (pc | tap | k) != 0);   // [Ah.Wh | Ah.Wl]
            }
          }
          umma_commit(&a_empty[slot]);
        }
        umma_commit(&acc_full[buf]);
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue
    const int q = warp & 3;                                   // TMEM lane quarter this warp may read
    constexpr int NCH = BN / 32;                              // 32-column chunks per tile, split between the two warps of a quarter
    const int c_split = epi_warps > 4 ? (NCH + 1) / 2 : NCH;
    const int c_begin = (warp - 2) < 4 ? 0 : c_split, c_end = (warp - 2) < 4 ? c_split : NCH;
    uint8_t* my_stage = epi_stage + (warp - 2) * EPI_STAGE;
    uint32_t local = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++local) {
      const int buf = local & 1;
      TileCoord t;
      t.x0 = (tile % tiles_x) * HTILE_W;
      t.y0 = ((tile / tiles_x) % tiles_y) * HTILE_H;
      t.n = tile / (tiles_x * tiles_y);
      t.g = 0;
      t.co0 = 0;
      t.ph = t.oy = t.ox = 0;
      mbar_wait(&acc_full[buf], (local >> 1) & 1);
      tc_fence_after_sync();
      epilogue_tile<BN>(p, t, tbase + (static_cast<uint32_t>(q * 32) << 16) + buf * Cfg<BN>::ACC_COLS, q * 32 + lane,
                        p.Cout, bias_s, c_begin, c_end, my_stage, HTILE_W, HTILE_H);
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tbase, TMEM_COLS);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess) p = nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

// NCHW fp32 (C <= cin channels) -> row-gapped NHWC bf16 (hi, lo) [N][H][lead + W][cin] + tail, zeros in the gaps, the
// tail and channels >= C: the operand layout of the window-packed conv.  One thread per (row pixel incl. gap, n*H+y).
__global__ void __launch_bounds__(256) pack_rows_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ hi,
                                                        __nv_bfloat16* __restrict__ lo, int N, int C, int H, int W, int cin,
                                                        int lead, int pitch, int tail) {
  const long long total = static_cast<long long>(N) * H * pitch + tail;
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const long long row = i / pitch;
  const int xp = static_cast<int>(i - row * pitch) - lead;
  const bool inside = row < static_cast<long long>(N) * H && xp >= 0 && xp < W;
  const long long n = row / H;
  const int y = static_cast<int>(row - n * H);
  for (int c0 = 0; c0 < cin; c0 += 4) {          // cin is a multiple of 4: 8-byte stores
    float f[4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
      f[j] = (inside && c0 + j < C) ? __ldg(x + ((n * C + c0 + j) * H + y) * static_cast<long long>(W) + xp) : 0.f;
    const __nv_bfloat162 h0 = __floats2bfloat162_rn(f[0], f[1]), h1 = __floats2bfloat162_rn(f[2], f[3]);
    const float2 g0 = __bfloat1622float2(h0), g1 = __bfloat1622float2(h1);
    const __nv_bfloat162 l0 = __floats2bfloat162_rn(f[0] - g0.x, f[1] - g0.y), l1 = __floats2bfloat162_rn(f[2] - g1.x, f[3] - g1.y);
    *reinterpret_cast<uint2*>(hi + i * cin + c0) = make_uint2(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
    *reinterpret_cast<uint2*>(lo + i * cin + c0) = make_uint2(*reinterpret_cast<const uint32_t*>(&l0), *reinterpret_cast<const uint32_t*>(&l1));
  }
}

}  // namespace conv

int conv_rows_tail(int lead, int channels) { return lead + (conv::BK + channels - 1) / channels; }

// pixels per row of the row-gapped layout: lead + W, rounded up so that a row is a multiple of 16 bytes (TMA stride
// rule; only matters for 4 channels); the extra pixel, if any, sits at the END of the row and is zero as well
int conv_rows_pitch(int w, int lead, int channels) {
  const int unit = channels >= 8 ? 1 : 8 / channels;
  return (w + lead + unit - 1) / unit * unit;
}

int launch_pack_rows(const float* x, void* hi, void* lo, int n, int c, int h, int w, int cin, int lead,
                     cudaStream_t stream) {
  const int pitch = conv_rows_pitch(w, lead, cin);
  const long long total = static_cast<long long>(n) * h * pitch + conv_rows_tail(lead, cin);
  const unsigned blocks = static_cast<unsigned>((total + 255) / 256);
  conv::pack_rows_kernel<<<blocks, 256, 0, stream>>>(x, static_cast<__nv_bfloat16*>(hi), static_cast<__nv_bfloat16*>(lo), n, c,
                                                     h, w, cin, lead, pitch, conv_rows_tail(lead, cin));
  count_launch();
  return static_cast<int>(cudaGetLastError());
}

int launch_conv3x3(int nsrc, const void* const* src_hi, const void* const* src_lo, const int* src_channels,
                   const void* w_hi, const void* w_lo, const float* bias, const float* residual, float* out,
                   void* out_hi, void* out_lo, int n, int h_in, int w_in, int cout, int groups, float slope, int ks,
                   int stride, int pad, int in_rows, int out_lead, cudaStream_t stream, const ConvGeom* geom, int epi_flags) {
  using namespace conv;
  if (epi_flags && ((cout & 3) == 0 || out_hi || residual || geom)) {
    set_error("conv: tanh / NCHW epilogues are implemented for the element-wise store path (Cout %% 4 != 0, fp32 output only)");
    return -2;
  }
  // GEMM grid = output size of the plain conv, or what the generalised geometry says
  const int h = geom ? geom->grid_h : (h_in + 2 * pad - ks) / stride + 1;
  const int w = geom ? geom->grid_w : (w_in + 2 * pad - ks) / stride + 1;
  // Tile shape.  Images smaller than a 16 x 8 tile (SPyNet's coarse pyramid levels: 2x4, 4x8 pixels) take a tile of their
  // own size: the A boxes shrink from 16 KB to 1-4 KB per K block, and these few-CTA launches are bound by the per-SM L2
  // port.  Otherwise the tile_w x tile_h <= 128 box that covers the image with the FEWEST tiles: at 60 x 108 (every
  // propagation / encoder conv of a 432x240 clip) 12 x 10 tiles the image exactly with 54 tiles where 16 x 8 needs 56 —
  // at 8 clips that is 432 instead of 448 tiles on 148 SMs, i.e. 3 waves instead of 4.
  int tile_w = TILE_W, tile_h = TILE_H;
  if (geom) {
    tile_w = geom->tile_w;
    tile_h = geom->tile_h;
  } else if (w < TILE_W || h < TILE_H) {
    tile_w = w < TILE_W ? w : TILE_W;
    tile_h = h < TILE_H ? h : TILE_H;
  } else if (!in_rows) {
    long long best = static_cast<long long>((h + TILE_H - 1) / TILE_H) * ((w + TILE_W - 1) / TILE_W);
    for (int tw = 32; tw >= 8; --tw) {                        // ties go to the wider tile (longer contiguous TMA rows)
      const int th = BM / tw;
      if (th < 4 || th > h || tw > w || tw * stride > 256 || th * stride > 256) continue;
      const long long cnt = static_cast<long long>((h + th - 1) / th) * ((w + tw - 1) / tw);
      if (cnt < best) {
        best = cnt;
        tile_w = tw;
        tile_h = th;
      }
    }
  }
  const int ntaps = geom ? geom->ntaps : ks * ks;
  if (ntaps < 1 || ntaps > 64 || tile_w < 1 || tile_h < 1 || tile_w * tile_h > BM || tile_w * stride > 256 ||
      tile_h * stride > 256 || (geom && (geom->nphase < 1 || geom->nphase > 9 || geom->ostep < 1))) {
    set_error("conv: unsupported geometry (taps=%d tile=%dx%d stride=%d)", ntaps, tile_w, tile_h, stride);
    return -2;
  }
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled is not available from the driver");
    return -4;
  }
  if (cout > MAX_COUT) {
    set_error("conv3x3: at most %d output channels (the bias is staged in shared memory), got %d", MAX_COUT, cout);
    return -2;
  }
  const int cog = cout / groups;
  int bn = cog <= 32 ? 32 : (cog <= 64 ? 64 : (cog == 96 ? 96 : 128));
  // few tiles (single-clip propagation steps: 51 pixel tiles on 148 SMs): halve the N tile so that twice as many SMs
  // work; a tile then costs (64 + 55) instead of (128 + 64) tensor-pipe cycles per K step (tools/mma_rate_probe.cu)
  if (bn == 128 && cog % 64 == 0 && !in_rows) {
    const long long t128 = static_cast<long long>(n) * ((h + tile_h - 1) / tile_h) * ((w + tile_w - 1) / tile_w) * groups *
                           ((cog + 127) / 128) * (geom ? geom->nphase : 1);
    if (2 * t128 <= num_sms()) bn = 64;
  }
  Maps maps;
  Params p;
  p.N = n; p.H = h; p.W = w; p.Cout = cout; p.groups = groups; p.nsrc = nsrc;
  p.ks = ks; p.stride = stride; p.pad = pad;
  p.slope = slope; p.bias = bias; p.residual = residual; p.out = out; p.epi_flags = epi_flags;
  p.out_hi = static_cast<__nv_bfloat16*>(out_hi); p.out_lo = static_cast<__nv_bfloat16*>(out_lo);
  p.chunks_total = 0;
  p.rows_px = p.rows_g = 0;
  p.out_lead = out_lead;
  p.tile_w = tile_w; p.tile_h = tile_h;
  p.bias_map = nullptr;
  if (geom) {
    p.nphase = geom->nphase; p.ostep = geom->ostep; p.out_H = geom->out_h; p.out_W = geom->out_w;
    p.bias_map = geom->bias_map;
    for (int i = 0; i < ntaps; ++i) { p.tap_dy[i] = geom->tap_dy[i]; p.tap_dx[i] = geom->tap_dx[i]; }
    for (int i = 0; i < geom->nphase; ++i) { p.ph_tap0[i] = geom->ph_tap0[i]; p.ph_oy[i] = geom->ph_oy[i]; p.ph_ox[i] = geom->ph_ox[i]; }
    p.ph_tap0[geom->nphase] = geom->ph_tap0[geom->nphase];
  } else {
    p.nphase = 1; p.ostep = 1; p.out_H = h; p.out_W = w;
    for (int i = 0; i < ntaps; ++i) { p.tap_dy[i] = static_cast<int8_t>(i / ks - pad); p.tap_dx[i] = static_cast<int8_t>(i % ks - pad); }
    p.ph_tap0[0] = 0; p.ph_tap0[1] = static_cast<uint8_t>(ntaps); p.ph_oy[0] = p.ph_ox[0] = 0;
  }
  p.out_pitch = conv_rows_pitch(p.out_W, out_lead, cout);   // == out_W + out_lead: split outputs have >= 8 channels
  p.out_nstride = static_cast<long long>(p.out_H) * p.out_W;
  p.osp_nstride = static_cast<long long>(p.out_H) * p.out_pitch;
  if (geom && geom->out_nstride > 0) {                       // batch-strided output (a frame slice of a (b, t, h, w, c) buffer)
    if (out_lead || geom->out_nstride < p.out_nstride) {
      set_error("conv: a batch-strided output needs a dense row layout and a stride >= out_h * out_w pixels");
      return -2;
    }
    p.out_nstride = p.osp_nstride = geom->out_nstride;
  }
  p.out_tail = out_lead ? conv_rows_tail(out_lead, cout) : 0;
  for (int i = 0; i < MAX_SRC; ++i) p.cig[i] = p.chunks[i] = 0;
  if (in_rows) {
    // ONE row-gapped source [N][H][w_in + pad][cin] (+ tail): dimension 1 steps by `stride` pixels, dimension 0 spans
    // 64 elements = 64/cin pixels (overlapping windows; validated by tools/tma_window_probe.cu)
    const int cin = src_channels[0];
    p.rows_px = BK / cin;
    p.rows_g = (ks + p.rows_px - 1) / p.rows_px;
    p.cig[0] = cin;
    p.chunks[0] = 1;
    p.chunks_total = 1;
    const int pitch = conv_rows_pitch(w_in, pad, cin);
    const cuuint64_t dims[4] = {BK, static_cast<cuuint64_t>((pitch + stride - 1) / stride), static_cast<cuuint64_t>(h_in),
                                static_cast<cuuint64_t>(n)};
    const cuuint64_t strides[3] = {static_cast<cuuint64_t>(stride) * cin * 2, static_cast<cuuint64_t>(pitch) * cin * 2,
                                   static_cast<cuuint64_t>(h_in) * pitch * cin * 2};
    const cuuint32_t box[4] = {BK, static_cast<cuuint32_t>(tile_w), static_cast<cuuint32_t>(tile_h * stride), 1};
    const cuuint32_t estr[4] = {1, 1, static_cast<cuuint32_t>(stride), 1};
    for (int part = 0; part < 2; ++part) {
      CUresult r = enc(part ? &maps.a_lo[0] : &maps.a_hi[0], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4,
                       const_cast<void*>(part ? src_lo[0] : src_hi[0]), dims, strides, box, estr,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) {
        set_error("conv2d: cuTensorMapEncodeTiled(row-gapped source) failed with CUresult %d (cin=%d w=%d h=%d n=%d)",
                  static_cast<int>(r), cin, w_in, h_in, n);
        return -4;
      }
    }
  }
  // HALO variant (see conv3x3_halo_kernel): dense 3x3 / s1 / p1 layers with <= 64 output channels whose weights fit
  // in shared memory next to >= 3 halo slots.  E2F_CONV_HALO=0 forces the generic kernel (A/B timing, debugging).
  int chunks_all = 0;
  for (int i = 0; i < nsrc; ++i) chunks_all += (src_channels[i] / groups + BK - 1) / BK;
  int halo_slots = 0, halo_epi = 8;
  if (!geom && !in_rows && ks == 3 && stride == 1 && pad == 1 && groups == 1 && cout <= 64) {
    static const bool enabled = [] {
      const char* e = getenv("E2F_CONV_HALO");
      return !(e && e[0] == '0');
    }();
    // 8 epilogue warps when their staging buffers fit next to the resident weights, else 4 (64 -> 64: 144 KB of weights)
    const int room8 = SMEM_LIMIT - 1024 - 256 - MAX_COUT * 4 - halo_w_bytes(bn, chunks_all) - 8 * EPI_STAGE;
    halo_epi = room8 >= 3 * HALO_SLOT ? 8 : 4;
    const int room = room8 + (8 - halo_epi) * EPI_STAGE;
    if (enabled && room >= 3 * HALO_SLOT) halo_slots = room / HALO_SLOT < HALO_MAX_SLOTS ? room / HALO_SLOT : HALO_MAX_SLOTS;
  }
  const cuuint32_t estr4[4] = {1, static_cast<cuuint32_t>(stride), static_cast<cuuint32_t>(stride), 1};
  for (int i = 0; i < (in_rows ? 0 : nsrc); ++i) {
    const int c = src_channels[i];
    p.cig[i] = c / groups;
    p.chunks[i] = (p.cig[i] + BK - 1) / BK;
    p.chunks_total += p.chunks[i];
    const cuuint64_t dims[4] = {static_cast<cuuint64_t>(c), static_cast<cuuint64_t>(w_in),
                                static_cast<cuuint64_t>(h_in), static_cast<cuuint64_t>(n)};
    long long nstride = static_cast<long long>(h_in) * w_in;                       // pixels between consecutive images
    if (geom && geom->src_nstride && geom->src_nstride[i] > 0) {
      if (geom->src_nstride[i] < nstride) {
        set_error("conv: source %d batch stride %lld is smaller than one image (%lld pixels)", i, geom->src_nstride[i], nstride);
        return -2;
      }
      nstride = geom->src_nstride[i];
    }
    const cuuint64_t strides[3] = {static_cast<cuuint64_t>(c) * 2, static_cast<cuuint64_t>(w_in) * c * 2,
                                   static_cast<cuuint64_t>(nstride) * c * 2};
    // the box spans TILE*stride input elements and is traversed with elementStrides = stride: TILE elements land
    const cuuint32_t box[4] = {BK, static_cast<cuuint32_t>(halo_slots ? HALO_W : tile_w * stride),
                               static_cast<cuuint32_t>(halo_slots ? HALO_H : tile_h * stride), 1};
    for (int part = 0; part < 2; ++part) {
      CUresult r = enc(part ? &maps.a_lo[i] : &maps.a_hi[i], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4,
                       const_cast<void*>(part ? src_lo[i] : src_hi[i]), dims, strides, box, estr4,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) {
        set_error("conv2d: cuTensorMapEncodeTiled(source %d) failed with CUresult %d (c=%d w=%d h=%d n=%d)", i,
                  static_cast<int>(r), c, w_in, h_in, n);
        return -4;
      }
    }
  }
  {
    const int kpad = (in_rows ? ks * p.rows_g : ntaps * p.chunks_total) * BK;
    const cuuint64_t dims[2] = {static_cast<cuuint64_t>(kpad), static_cast<cuuint64_t>(cout)};
    const cuuint64_t strides[1] = {static_cast<cuuint64_t>(kpad) * 2};
    const cuuint32_t box[2] = {BK, static_cast<cuuint32_t>(bn)};
    const cuuint32_t estr[2] = {1, 1};
    for (int part = 0; part < 2; ++part) {
      CUresult r = enc(part ? &maps.w_lo : &maps.w_hi, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                       const_cast<void*>(part ? w_lo : w_hi), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) {
        set_error("conv3x3: cuTensorMapEncodeTiled(weight) failed with CUresult %d", static_cast<int>(r));
        return -4;
      }
    }
  }
  static DeviceOnce configured, halo_configured;
  const int dev = current_device();
  if (!device_done(configured, dev)) {
    cudaError_t e = cudaFuncSetAttribute(conv3x3_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<128>::SMEM);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(conv3x3_kernel<96>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<96>::SMEM);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(conv3x3_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<64>::SMEM);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(conv3x3_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<32>::SMEM);
    if (e != cudaSuccess) return static_cast<int>(e);
    device_mark(configured, dev);
  }
  if (halo_slots) {
    if (!device_done(halo_configured, dev)) {
      cudaError_t e = cudaFuncSetAttribute(conv3x3_halo_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT);
      if (e == cudaSuccess)
        e = cudaFuncSetAttribute(conv3x3_halo_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT);
      if (e != cudaSuccess) return static_cast<int>(e);
      device_mark(halo_configured, dev);
    }
    const long long htiles = static_cast<long long>(n) * ((h + HTILE_H - 1) / HTILE_H) * ((w + HTILE_W - 1) / HTILE_W);
    if (htiles == 0) return 0;
    if (htiles > 0x7FFFFFFFLL) {
      set_error("conv3x3: too many tiles");
      return -2;
    }
    const int hgrid = htiles < num_sms() ? static_cast<int>(htiles) : num_sms();
    const int hsmem = 1024 + halo_w_bytes(bn, chunks_all) + halo_slots * HALO_SLOT + 256 + MAX_COUT * 4 + halo_epi * EPI_STAGE;
    const int hthreads = (2 + halo_epi) * 32;
    if (bn == 32)
      conv3x3_halo_kernel<32><<<hgrid, hthreads, hsmem, stream>>>(maps, p, halo_slots);
    else
      conv3x3_halo_kernel<64><<<hgrid, hthreads, hsmem, stream>>>(maps, p, halo_slots);
    count_launch();
    return static_cast<int>(cudaGetLastError());
  }
  const int tiles_y = (h + tile_h - 1) / tile_h, tiles_x = (w + tile_w - 1) / tile_w;
  const int tiles_ng = (cout / groups + bn - 1) / bn;
  const long long tiles = static_cast<long long>(n) * tiles_y * tiles_x * p.nphase * groups * tiles_ng;
  if (tiles == 0) return 0;
  if (tiles > 0x7FFFFFFFLL) {
    set_error("conv3x3: too many tiles");
    return -2;
  }
  const int grid = tiles < num_sms() ? static_cast<int>(tiles) : num_sms();
  if (bn == 32)
    conv3x3_kernel<32><<<grid, THREADS, Cfg<32>::SMEM, stream>>>(maps, p);
  else if (bn == 64)
    conv3x3_kernel<64><<<grid, THREADS, Cfg<64>::SMEM, stream>>>(maps, p);
  else if (bn == 96)
    conv3x3_kernel<96><<<grid, THREADS, Cfg<96>::SMEM, stream>>>(maps, p);
  else
    conv3x3_kernel<128><<<grid, THREADS, Cfg<128>::SMEM, stream>>>(maps, p);
  count_launch();
  return static_cast<int>(cudaGetLastError());
}

}  // namespace e2f
