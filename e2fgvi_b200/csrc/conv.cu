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
//  * epilogue: + bias, LeakyReLU(slope), optional residual add, fp32 NHWC store.
// Pipeline = gemm.cu: persistent CTAs, TMA warp / MMA warp / 4 epilogue warps, double-buffered TMEM accumulator.
#include <cuda.h>
#include <cuda_bf16.h>
#include "common.cuh"
#include "launch.h"

namespace e2f {
namespace conv {

constexpr int BM = 128, BK = 64;
constexpr int TILE_H = 8, TILE_W = 16;                  // 8 x 16 output pixels = 128 GEMM rows
constexpr int A_TILE = BM * BK * 2;
constexpr int EPI_WARPS = 4;
constexpr int THREADS = (2 + EPI_WARPS) * 32;
constexpr int MAX_SRC = 4;

// N tile: 128 output channels, or 64 for the layers with <= 64 output channels per group (decoder, encoder conv 1,
// the 3-channel output conv), which would otherwise waste half of every MMA.
template <int BN>
struct Cfg {
  static constexpr int W_TILE = BN * BK * 2;
  static constexpr int STAGE = 2 * A_TILE + 2 * W_TILE;   // 64 KB (BN=128) / 48 KB (BN=64)
  static constexpr int STAGES = (BN == 128) ? 3 : 4;
  static constexpr int TMEM_COLS = 2 * BN;
  static constexpr int SMEM = STAGES * STAGE + 256 + 1024;
};

struct Maps {
  CUtensorMap a_hi[MAX_SRC], a_lo[MAX_SRC], w_hi, w_lo;
};

struct Params {
  int N, H, W, Cout, groups;     // H, W: OUTPUT spatial size
  int ks, stride, pad;           // square kernel size (3 or 7), stride (1 or 2), zero padding
  int nsrc;
  int cig[MAX_SRC];        // channels per group of each source
  int chunks[MAX_SRC];     // ceil(cig / 64)
  int chunks_total;        // sum of chunks
  float slope;             // LeakyReLU negative slope (1 = identity)
  const float* bias;
  const float* residual;   // NHWC fp32 [N][H][W][Cout] or null
  float* out;              // NHWC fp32 or null
  __nv_bfloat16* out_hi;   // NHWC bf16 split of the result (operand format of the next conv) or null
  __nv_bfloat16* out_lo;
};

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
};

template <int BN>
__device__ __forceinline__ TileCoord decode_tile(int tile, const Params& p, int tiles_y, int tiles_x, int tiles_ng) {
  TileCoord t;
  const int nt = tile % tiles_ng;
  int r = tile / tiles_ng;
  t.g = r % p.groups;
  r /= p.groups;
  const int tx = r % tiles_x;
  r /= tiles_x;
  const int ty = r % tiles_y;
  t.n = r / tiles_y;
  t.y0 = ty * TILE_H;
  t.x0 = tx * TILE_W;
  t.co0 = t.g * (p.Cout / p.groups) + nt * BN;
  return t;
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

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tiles_y = (p.H + TILE_H - 1) / TILE_H, tiles_x = (p.W + TILE_W - 1) / TILE_W;
  const int cog = p.Cout / p.groups;
  const int tiles_ng = (cog + BN - 1) / BN;
  const int num_tiles = p.N * tiles_y * tiles_x * p.groups * tiles_ng;
  const int num_kb = p.ks * p.ks * p.chunks_total;

  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], EPI_WARPS);
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
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const TileCoord t = decode_tile<BN>(tile, p, tiles_y, tiles_x, tiles_ng);
        int kb = 0;
        const int taps = p.ks * p.ks;
        for (int tap = 0; tap < taps; ++tap) {
          // input coordinate of the tile's first output pixel for this tap (TMA steps by `stride` inside the box)
          const int yy = t.y0 * p.stride - p.pad + tap / p.ks, xx = t.x0 * p.stride - p.pad + tap % p.ks;
          for (int s = 0; s < p.nsrc; ++s) {
            const int c_base = t.g * p.cig[s];
            for (int j = 0; j < p.chunks[s]; ++j, ++kb, ++it) {
              const int stage = it % STAGES;
              mbar_wait(&empty[stage], ((it / STAGES) & 1) ^ 1);
              mbar_arrive_expect_tx(&full[stage], STAGE);
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
    if (lane == 0) {
      const uint32_t idesc = idesc_bf16(BM, BN);
      const uint64_t d_ah0 = umma_desc_sw128(smem_u32(smem), 16, 1024);
      const uint64_t d_al0 = umma_desc_adv(d_ah0, A_TILE), d_wh0 = umma_desc_adv(d_ah0, 2 * A_TILE);
      const uint64_t d_wl0 = umma_desc_adv(d_wh0, W_TILE);
      uint32_t it = 0, local = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++local) {
        const int buf = local & 1;
        mbar_wait(&acc_empty[buf], ((local >> 1) & 1) ^ 1);
        tc_fence_after_sync();
        const uint32_t d = tbase + buf * BN;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int stage = it % STAGES;
          mbar_wait(&full[stage], (it / STAGES) & 1);
          tc_fence_after_sync();
          const uint32_t soff = (stage * STAGE) >> 4;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t dah = d_ah0 + soff + 2 * k, dal = d_al0 + soff + 2 * k;
            const uint64_t dwh = d_wh0 + soff + 2 * k, dwl = d_wl0 + soff + 2 * k;
            umma_f16(d, dal, dwh, idesc, (kb | k) != 0);
            umma_f16(d, dah, dwl, idesc, 1);
            umma_f16(d, dah, dwh, idesc, 1);
          }
          umma_commit(&empty[stage]);
        }
        umma_commit(&acc_full[buf]);
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue
    const int q = warp & 3;
    uint32_t local = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++local) {
      const int buf = local & 1;
      const TileCoord t = decode_tile<BN>(tile, p, tiles_y, tiles_x, tiles_ng);
      mbar_wait(&acc_full[buf], (local >> 1) & 1);
      tc_fence_after_sync();
      const int r = q * 32 + lane;
      const int y = t.y0 + r / TILE_W, x = t.x0 + r % TILE_W;
      const bool pix_ok = (y < p.H) && (x < p.W);
      const size_t pix = (static_cast<size_t>(t.n) * p.H + y) * p.W + x;
      const int co_end = (t.g + 1) * cog;               // exclusive end of this group's output channels
      const uint32_t taddr = tbase + (static_cast<uint32_t>(q * 32) << 16) + buf * BN;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t v[32];
        tmem_ld32(taddr + c * 32, v);
        tmem_ld_wait();
        const int co = t.co0 + c * 32;
        if (pix_ok && co < co_end) {
          const size_t o = pix * p.Cout + co;
          if (co + 32 <= co_end) {
            float f[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              float a = __uint_as_float(v[i]) + (p.bias ? __ldg(p.bias + co + i) : 0.f);
              f[i] = a > 0.f ? a : a * p.slope;
            }
            if (p.residual) {
              const float4* r4 = reinterpret_cast<const float4*>(p.residual + o);
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float4 rr = __ldg(r4 + i);
                f[4 * i] += rr.x; f[4 * i + 1] += rr.y; f[4 * i + 2] += rr.z; f[4 * i + 3] += rr.w;
              }
            }
            if (p.out) {
              float4* d4 = reinterpret_cast<float4*>(p.out + o);
#pragma unroll
              for (int i = 0; i < 8; ++i) d4[i] = make_float4(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
            }
            if (p.out_hi) {
              uint32_t hp[16], lp[16];       // packed bf16 pairs, kept in registers
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const __nv_bfloat162 hb = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
                const float2 hf = __bfloat1622float2(hb);
                const __nv_bfloat162 lb = __floats2bfloat162_rn(f[2 * i] - hf.x, f[2 * i + 1] - hf.y);
                hp[i] = *reinterpret_cast<const uint32_t*>(&hb);
                lp[i] = *reinterpret_cast<const uint32_t*>(&lb);
              }
              uint4* dh = reinterpret_cast<uint4*>(p.out_hi + o);
              uint4* dl = reinterpret_cast<uint4*>(p.out_lo + o);
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                dh[i] = make_uint4(hp[4 * i], hp[4 * i + 1], hp[4 * i + 2], hp[4 * i + 3]);
                dl[i] = make_uint4(lp[4 * i], lp[4 * i + 1], lp[4 * i + 2], lp[4 * i + 3]);
              }
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              if (co + i < co_end) {
                float a = __uint_as_float(v[i]) + (p.bias ? __ldg(p.bias + co + i) : 0.f);
                a = a > 0.f ? a : a * p.slope;
                if (p.residual) a += __ldg(p.residual + o + i);
                if (p.out) p.out[o + i] = a;
                if (p.out_hi) {
                  const __nv_bfloat16 hb = __float2bfloat16_rn(a);
                  p.out_hi[o + i] = hb;
                  p.out_lo[o + i] = __float2bfloat16_rn(a - __bfloat162float(hb));
                }
              }
            }
          }
        }
      }
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

static int num_sms() {
  static int n = [] {
    int dev = 0, v = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    return v;
  }();
  return n;
}

}  // namespace conv

int launch_conv3x3(int nsrc, const void* const* src_hi, const void* const* src_lo, const int* src_channels,
                   const void* w_hi, const void* w_lo, const float* bias, const float* residual, float* out,
                   void* out_hi, void* out_lo, int n, int h_in, int w_in, int cout, int groups, float slope, int ks,
                   int stride, int pad, cudaStream_t stream) {
  using namespace conv;
  const int h = (h_in + 2 * pad - ks) / stride + 1, w = (w_in + 2 * pad - ks) / stride + 1;   // output size
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled is not available from the driver");
    return -4;
  }
  const int bn = (cout / groups <= 64) ? 64 : 128;
  Maps maps;
  Params p;
  p.N = n; p.H = h; p.W = w; p.Cout = cout; p.groups = groups; p.nsrc = nsrc;
  p.ks = ks; p.stride = stride; p.pad = pad;
  p.slope = slope; p.bias = bias; p.residual = residual; p.out = out;
  p.out_hi = static_cast<__nv_bfloat16*>(out_hi); p.out_lo = static_cast<__nv_bfloat16*>(out_lo);
  p.chunks_total = 0;
  for (int i = 0; i < MAX_SRC; ++i) p.cig[i] = p.chunks[i] = 0;
  const cuuint32_t estr4[4] = {1, static_cast<cuuint32_t>(stride), static_cast<cuuint32_t>(stride), 1};
  for (int i = 0; i < nsrc; ++i) {
    const int c = src_channels[i];
    p.cig[i] = c / groups;
    p.chunks[i] = (p.cig[i] + BK - 1) / BK;
    p.chunks_total += p.chunks[i];
    const cuuint64_t dims[4] = {static_cast<cuuint64_t>(c), static_cast<cuuint64_t>(w_in),
                                static_cast<cuuint64_t>(h_in), static_cast<cuuint64_t>(n)};
    const cuuint64_t strides[3] = {static_cast<cuuint64_t>(c) * 2, static_cast<cuuint64_t>(w_in) * c * 2,
                                   static_cast<cuuint64_t>(h_in) * w_in * c * 2};
    // the box spans TILE*stride input elements and is traversed with elementStrides = stride: TILE elements land
    const cuuint32_t box[4] = {BK, static_cast<cuuint32_t>(TILE_W * stride), static_cast<cuuint32_t>(TILE_H * stride), 1};
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
    const int kpad = ks * ks * p.chunks_total * BK;
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
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(conv3x3_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<128>::SMEM);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(conv3x3_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<64>::SMEM);
    if (e != cudaSuccess) return static_cast<int>(e);
    configured = true;
  }
  const int tiles_y = (h + TILE_H - 1) / TILE_H, tiles_x = (w + TILE_W - 1) / TILE_W;
  const int tiles_ng = (cout / groups + bn - 1) / bn;
  const long long tiles = static_cast<long long>(n) * tiles_y * tiles_x * groups * tiles_ng;
  if (tiles == 0) return 0;
  if (tiles > 0x7FFFFFFFLL) {
    set_error("conv3x3: too many tiles");
    return -2;
  }
  const int grid = tiles < num_sms() ? static_cast<int>(tiles) : num_sms();
  if (bn == 64)
    conv3x3_kernel<64><<<grid, THREADS, Cfg<64>::SMEM, stream>>>(maps, p);
  else
    conv3x3_kernel<128><<<grid, THREADS, Cfg<128>::SMEM, stream>>>(maps, p);
  count_launch();
  return static_cast<int>(cudaGetLastError());
}

}  // namespace e2f
