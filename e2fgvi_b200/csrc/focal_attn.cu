// Temporal focal window attention core — replaces model/modules/tfocal_transformer.py:226-396 (+ window_reverse
// :528): softmax(q k_all^T) v_all per (window, head) with k_all = own window | 4 circularly rolled ring sets |
// pooled-window neighbourhood, without materialising rolled copies, the key list or the logits.
//
// Key set per (window (wi,wj), frame t), equivalent to the reference's list (order is irrelevant to softmax.V):
//   * the expanded window (wh+2eh) x (ww+2ew) around the query window, coordinates wrapped modulo (H, W) exactly
//     like torch.roll; a token listed m times by the reference (own window + tl/tr/bl/br rolls after
//     valid_ind_rolled; m = 2 for the 12 duplicated ring tokens) gets log2(m) added to its log2-domain logit;
//   * the in-grid pooled windows of the (fh x fw) neighbourhood; zero-padded neighbours have k = v = 0 and a -100
//     logit in the reference (:301-316, :377-380): they only add n_masked * exp(-100) to the softmax denominator,
//     which is folded into the initial (running max, running sum) = (-100, n_masked).
//   Keys are ordered [multiplicity-1 ring keys | pooled keys | multiplicity>=2 ring keys | padding], so only the
//   last key tile(s) carry a non-zero logit bias and every other tile takes a bias-free fast path.
//
// One CTA = one (128-query tile, head, window).  Warp roles (416 threads):
//   warps 0-7   softmax: two warps per SM sub-partition; thread (q, lane, hh) owns HALF of query row q*32+lane
//               (64 of the 128 keys of a tile): S -> registers, row max exchanged through smem, online softmax with
//               lazy rescale, P -> fp16 128B-swizzled smem, final O / l -> global (un-partitioned layout).
//               (r01 v1 had one softmax warp per sub-partition and was issue-bound: tensor pipe 16 %.)
//   warps 8-11  loaders: per-key source addresses (wrap / pooled / padding) then coalesced 16-byte cp.async gathers
//               of K and V rows (256 B each) into swizzled smem, 2-stage ring
//   warp  12    tcgen05 issuer: S = Q K^T (K-major operands), O += P V (V as MN-major B operand), fp32 in TMEM
// Roofline: 4*B*nW*heads*(T*wh*ww)*(T*(wh*ww+ring+fh*fw))*128 FLOP on the tensor pipe.
#include <cstdlib>
#include "common.cuh"
#include "launch.h"

namespace e2f {
namespace attn {

constexpr int HD = 128;                    // head dim
constexpr int BM = 128, BN = 128;          // query tile, key tile
constexpr int ATOM = 16384;                // one [128 rows][64 halfs] swizzled sub-tile
constexpr int TILE = 2 * ATOM;             // [128][128] fp16
constexpr int KV_STAGES = 2;
constexpr int SOFTMAX_WARPS = 8, LOADER_WARPS = 4;
constexpr int SOFTMAX_THREADS = SOFTMAX_WARPS * 32;
constexpr int MMA_WARP = SOFTMAX_WARPS + LOADER_WARPS;
constexpr int THREADS = (MMA_WARP + 1) * 32;   // 416
constexpr int TMEM_COLS = 512;                 // S0 [0,128) S1 [128,256) O [256,384)
constexpr uint32_t COL_S = 0, COL_O = 256;
constexpr float LOG2E = 1.4426950408889634f;
constexpr float RESCALE_THRESHOLD = 8.0f;      // log2 domain: P stays <= 2^8
constexpr int MAX_RING = 256;                  // expanded-window positions (153 for the 5x9 window)

struct Smem {
  static constexpr int Q = 0;
  static constexpr int K = Q + TILE;
  static constexpr int V = K + KV_STAGES * TILE;
  static constexpr int P = V + KV_STAGES * TILE;
  static constexpr int KEYPTR = P + TILE;                       // [stages][128] uint64
  static constexpr int BIAS = KEYPTR + KV_STAGES * BN * 8;      // [stages][128] float
  static constexpr int XCHG = BIAS + KV_STAGES * BN * 4;        // [2 parity][2 halves][128 rows] float (row max)
  static constexpr int XSUM = XCHG + 2 * 2 * BM * 4;            // [2 halves][128 rows] float (row sum)
  static constexpr int BARS = XSUM + 2 * BM * 4;
  static constexpr int NUM_BARS = 1 + 3 * KV_STAGES + 2 + 2 + 2;
  static constexpr int TMEM_SLOT = BARS + NUM_BARS * 8;
  static constexpr int BYTES = TMEM_SLOT + 16;
};
constexpr int SMEM_BYTES = Smem::BYTES + 1024;

struct Params {
  const __half* qkv;
  const __half* pooled;
  void* out;
  int B, T, H, W, heads, C;       // C = heads*128
  int wh, ww, eh, ew, fh, fw;
  int nWh, nWw;
  int use_pooled;
  float scale_log2;               // scale * log2(e)
  int debug;                      // perf-experiment bits (E2F_ATTN_DEBUG): 1 skip softmax math, 2 skip gathers, 4 skip MMAs
  int n1, n2;                     // expanded-window positions listed once / more than once by the reference
  uint8_t ring_pos[MAX_RING];     // positions (er*EW + ec): the n1 single ones first, then the n2 multiple ones
  uint8_t ring_mult[MAX_RING];    // multiplicity of each entry
};

// how many times the reference lists expanded-window position (er, ec) as a key (tfocal_transformer.py:166-179,235-280)
__host__ __device__ inline int key_multiplicity(int er, int ec, int wh, int ww, int eh, int ew) {
  int m = 0;
  if (er >= eh && er < eh + wh && ec >= ew && ec < ew + ww) m += 1;                      // own window
  {  // tl: window pos (r,c) holds token (r+eh, c+ew) -> expanded (r+2eh, c+2ew); kept if r>=wh-eh or c>=ww-ew
    const int r = er - 2 * eh, c = ec - 2 * ew;
    if (r >= 0 && r < wh && c >= 0 && c < ww && (r >= wh - eh || c >= ww - ew)) m += 1;
  }
  {  // tr: expanded (r+2eh, c); kept if r>=wh-eh or c<ew
    const int r = er - 2 * eh, c = ec;
    if (r >= 0 && r < wh && c >= 0 && c < ww && (r >= wh - eh || c < ew)) m += 1;
  }
  {  // bl: expanded (r, c+2ew); kept if r<eh or c>=ww-ew
    const int r = er, c = ec - 2 * ew;
    if (r >= 0 && r < wh && c >= 0 && c < ww && (r < eh || c >= ww - ew)) m += 1;
  }
  {  // br: expanded (r, c); kept if r<eh or c<ew
    const int r = er, c = ec;
    if (r >= 0 && r < wh && c >= 0 && c < ww && (r < eh || c < ew)) m += 1;
  }
  return m;
}

__device__ __forceinline__ void loader_barrier() { asm volatile("bar.sync 1, 128;" ::: "memory"); }
__device__ __forceinline__ void softmax_barrier() { asm volatile("bar.sync 2, 256;" ::: "memory"); }

// coalesced gather of 128 rows x 256 B (two 64-half atoms) into a swizzled tile; 16 lanes cover one row.
__device__ __forceinline__ void gather_rows(uint32_t tile_smem, const uint64_t* row_ptr, int lwarp, int lane,
                                            int half_offset) {
  const int chunk = lane & 15;
  const uint32_t atom_off = (chunk >> 3) * ATOM;
#pragma unroll 4
  for (int it = 0; it < 16; ++it) {
    const int row = lwarp * 32 + it * 2 + (lane >> 4);
    const uint64_t p = row_ptr[row];
    const uint32_t dst = tile_smem + atom_off + sw128_offset(row, chunk & 7);
    const __half* src = reinterpret_cast<const __half*>(p) + half_offset + chunk * 8;
    cp_async16_zfill(dst, p ? static_cast<const void*>(src) : static_cast<const void*>(row_ptr), p ? 16u : 0u);
  }
}

template <typename OutT>
__global__ void __launch_bounds__(THREADS, 1) focal_attn_kernel(const __grid_constant__ Params prm) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* key_ptr = reinterpret_cast<uint64_t*>(smem + Smem::KEYPTR);
  float* key_bias = reinterpret_cast<float*>(smem + Smem::BIAS);
  float* xchg = reinterpret_cast<float*>(smem + Smem::XCHG);
  float* xsum = reinterpret_cast<float*>(smem + Smem::XSUM);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Smem::BARS);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;
  uint64_t* v_full = k_full + KV_STAGES;
  uint64_t* kv_empty = v_full + KV_STAGES;
  uint64_t* s_full = kv_empty + KV_STAGES;      // [2]
  uint64_t* s_free = s_full + 2;                // [2]
  uint64_t* p_full = s_free + 2;
  uint64_t* o_done = p_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + Smem::TMEM_SLOT);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // ---- problem geometry (uniform per CTA)
  const int area = prm.wh * prm.ww;
  const int nq = prm.T * area;
  const int qt = blockIdx.x, head = blockIdx.y;
  const int win = blockIdx.z % (prm.nWh * prm.nWw), b = blockIdx.z / (prm.nWh * prm.nWw);
  const int wi = win / prm.nWw, wj = win % prm.nWw;
  const int EW = prm.ww + 2 * prm.ew;
  int pi0 = 0, pj0 = 0, PH = 0, PW = 0;
  if (prm.use_pooled) {
    pi0 = max(0, wi - prm.fh / 2);
    pj0 = max(0, wj - prm.fw / 2);
    PH = min(prm.nWh - 1, wi + prm.fh / 2) - pi0 + 1;
    PW = min(prm.nWw - 1, wj + prm.fw / 2) - pj0 + 1;
  }
  const int npool = PH * PW;
  const int n_masked = prm.use_pooled ? prm.T * (prm.fh * prm.fw - npool) : 0;
  const int nA = prm.T * prm.n1;                 // single-listed ring keys (bias 0)
  const int nB = nA + prm.T * npool;             // + pooled keys (bias 0)
  const int NK = nB + prm.T * prm.n2;            // + multiply-listed ring keys (bias log2 m)
  const int num_kt = (NK + BN - 1) / BN;
  // first key tile that may contain a non-zero bias (multiplicity keys or -inf padding)
  const int bias_kt = (NK % BN) ? min(nB / BN, num_kt - 1) : ((prm.n2 > 0) ? nB / BN : num_kt);
  const size_t C3 = 3 * static_cast<size_t>(prm.C);

  if (warp == MMA_WARP) tmem_alloc(tmem_slot, TMEM_COLS);
  if (tid == 0) {
    mbar_init(q_full, LOADER_WARPS);
    for (int s = 0; s < KV_STAGES; ++s) {
      mbar_init(&k_full[s], LOADER_WARPS);
      mbar_init(&v_full[s], LOADER_WARPS);
      mbar_init(&kv_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&s_full[s], 1);
      mbar_init(&s_free[s], SOFTMAX_WARPS);
    }
    mbar_init(p_full, SOFTMAX_WARPS);
    mbar_init(o_done, 1);
    fence_barrier_init();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tbase = *tmem_slot;

  if (warp < SOFTMAX_WARPS) {
    // =================================================================== softmax + epilogue
    const int q = warp & 3, hh = warp >> 2;             // TMEM lane quarter, key/column half
    const int r = q * 32 + lane;                        // query row in the tile == TMEM lane
    const uint32_t lane_addr = tbase + (static_cast<uint32_t>(q * 32) << 16);
    float m_used = n_masked > 0 ? -100.0f * LOG2E : -INFINITY;
    float l = (hh == 0) ? static_cast<float>(n_masked) : 0.f;     // partial row sum of this thread's columns
    uint8_t* sP = smem + Smem::P + hh * ATOM;
    const float sc = prm.scale_log2;

    for (int kt = 0; kt < num_kt; ++kt) {
      const int sb = kt & 1, stage = kt % KV_STAGES;
      const bool biased = kt >= bias_kt;                // uniform over the CTA
      if (biased) mbar_wait(&k_full[stage], (kt / KV_STAGES) & 1);   // acquire the loaders' bias writes
      mbar_wait(&s_full[sb], (kt >> 1) & 1);
      tc_fence_after_sync();
      uint32_t sv[2][32];
      tmem_ld32(lane_addr + COL_S + sb * BN + hh * 64, sv[0]);
      tmem_ld32(lane_addr + COL_S + sb * BN + hh * 64 + 32, sv[1]);
      tmem_ld_wait();
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_free[sb]);

      float mx = -INFINITY;
      if (prm.debug & 1) {
        mx = 0.f;
      } else if (biased) {
        const float4* bias4 = reinterpret_cast<const float4*>(key_bias + stage * BN + hh * 64);
#pragma unroll
        for (int c = 0; c < 2; ++c) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 bb = bias4[c * 8 + i];
            const float s0 = fmaf(__uint_as_float(sv[c][4 * i + 0]), sc, bb.x);
            const float s1 = fmaf(__uint_as_float(sv[c][4 * i + 1]), sc, bb.y);
            const float s2 = fmaf(__uint_as_float(sv[c][4 * i + 2]), sc, bb.z);
            const float s3 = fmaf(__uint_as_float(sv[c][4 * i + 3]), sc, bb.w);
            sv[c][4 * i + 0] = __float_as_uint(s0);
            sv[c][4 * i + 1] = __float_as_uint(s1);
            sv[c][4 * i + 2] = __float_as_uint(s2);
            sv[c][4 * i + 3] = __float_as_uint(s3);
            mx = fmaxf(mx, fmaxf(fmaxf(s0, s1), fmaxf(s2, s3)));
          }
        }
      } else {
#pragma unroll
        for (int c = 0; c < 2; ++c) {
#pragma unroll
          for (int i = 0; i < 32; i += 4)
            mx = fmaxf(mx, fmaxf(fmaxf(__uint_as_float(sv[c][i]), __uint_as_float(sv[c][i + 1])),
                                 fmaxf(__uint_as_float(sv[c][i + 2]), __uint_as_float(sv[c][i + 3]))));
        }
        mx *= sc;                                       // scale > 0: max commutes with the scaling
      }
      // full-row max: exchange the two half-row maxima through smem (buffer alternates with the tile parity)
      float* xm = xchg + (kt & 1) * (2 * BM);
      xm[hh * BM + r] = mx;
      softmax_barrier();
      mx = fmaxf(mx, xm[(hh ^ 1) * BM + r]);

      // lazy rescale: only move the reference max when it grew by more than 2^8 (both half-row threads agree)
      float alpha = 1.0f;
      bool need = false;
      if (m_used == -INFINITY) {
        m_used = mx;                       // first tile without masked keys: nothing accumulated yet
      } else if (mx - m_used > RESCALE_THRESHOLD) {
        alpha = fast_exp2(m_used - mx);
        m_used = mx;
        need = true;
      }
      if (kt > 0) {                        // PV(kt-1) must be complete before P is overwritten / O is rescaled
        mbar_wait(o_done, (kt - 1) & 1);
        tc_fence_after_sync();
      }
      if (__any_sync(0xffffffffu, need)) {
        l *= alpha;
        if (kt > 0) {
#pragma unroll 1
          for (int c = 0; c < 2; ++c) {
            uint32_t ov[32];
            tmem_ld32(lane_addr + COL_O + hh * 64 + c * 32, ov);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * alpha);
            tmem_st32(lane_addr + COL_O + hh * 64 + c * 32, ov);
          }
          tmem_st_wait();
        }
      }
      float lsum = 0.f;
      const float neg_m = -m_used;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {       // 8 probabilities -> one 16-byte chunk of the P row
          float p[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float sval = __uint_as_float(sv[c][8 * i + e]);
            if (prm.debug & 1) p[e] = sval * 1e-3f;
            else p[e] = biased ? fast_exp2(sval + neg_m) : fast_exp2(fmaf(sval, sc, neg_m));
            lsum += p[e];
          }
          uint4 u;
          u.x = pack_half2(p[0], p[1]); u.y = pack_half2(p[2], p[3]);
          u.z = pack_half2(p[4], p[5]); u.w = pack_half2(p[6], p[7]);
          *reinterpret_cast<uint4*>(sP + sw128_offset(r, c * 4 + i)) = u;
        }
      }
      l += lsum;
      fence_proxy_async_smem();
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
    }

    // ---- epilogue: O / l -> out[b, t, y, x, head*128 + hh*64 ..]
    xsum[hh * BM + r] = l;
    softmax_barrier();
    const float inv_l = 1.0f / (l + xsum[(hh ^ 1) * BM + r]);
    mbar_wait(o_done, (num_kt - 1) & 1);
    tc_fence_after_sync();
    const int qi = qt * BM + r;
    OutT* dst = nullptr;
    if (qi < nq) {
      const int t = qi / area, p = qi - t * area;
      const int y = wi * prm.wh + p / prm.ww, x = wj * prm.ww + p % prm.ww;
      const size_t tok = ((static_cast<size_t>(b) * prm.T + t) * prm.H + y) * prm.W + x;
      dst = static_cast<OutT*>(prm.out) + tok * prm.C + head * HD + hh * 64;
    }
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
      uint32_t ov[32];
      tmem_ld32(lane_addr + COL_O + hh * 64 + c * 32, ov);
      tmem_ld_wait();
      if (dst) {
        if constexpr (sizeof(OutT) == 4) {
          float4* d4 = reinterpret_cast<float4*>(dst + c * 32);
#pragma unroll
          for (int i = 0; i < 8; ++i)
            d4[i] = make_float4(__uint_as_float(ov[4 * i]) * inv_l, __uint_as_float(ov[4 * i + 1]) * inv_l,
                                __uint_as_float(ov[4 * i + 2]) * inv_l, __uint_as_float(ov[4 * i + 3]) * inv_l);
        } else {
          uint4* d4 = reinterpret_cast<uint4*>(dst + c * 32);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            uint4 u;
            u.x = pack_half2(__uint_as_float(ov[8 * i]) * inv_l, __uint_as_float(ov[8 * i + 1]) * inv_l);
            u.y = pack_half2(__uint_as_float(ov[8 * i + 2]) * inv_l, __uint_as_float(ov[8 * i + 3]) * inv_l);
            u.z = pack_half2(__uint_as_float(ov[8 * i + 4]) * inv_l, __uint_as_float(ov[8 * i + 5]) * inv_l);
            u.w = pack_half2(__uint_as_float(ov[8 * i + 6]) * inv_l, __uint_as_float(ov[8 * i + 7]) * inv_l);
            d4[i] = u;
          }
        }
      }
    }
  } else if (warp < MMA_WARP) {
    // =================================================================== loaders
    const int lt = tid - SOFTMAX_THREADS;               // 0..127
    const int lwarp = lt >> 5;
    // Q tile: row lt's source address (key_ptr stage 0 is borrowed as scratch before the first K tile)
    {
      const int qi = qt * BM + lt;
      uint64_t p = 0;
      if (qi < nq) {
        const int t = qi / area, pp = qi - t * area;
        const int y = wi * prm.wh + pp / prm.ww, x = wj * prm.ww + pp % prm.ww;
        const size_t tok = ((static_cast<size_t>(b) * prm.T + t) * prm.H + y) * prm.W + x;
        p = reinterpret_cast<uint64_t>(prm.qkv + tok * C3 + head * HD);
      }
      key_ptr[lt] = p;
      loader_barrier();
      gather_rows(smem_u32(smem + Smem::Q), key_ptr, lwarp, lane, 0);
      cp_async_commit();
      cp_async_wait<0>();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(q_full);
      loader_barrier();                                 // everyone done reading key_ptr[0..127]
    }
    for (int kt = 0; kt < num_kt; ++kt) {
      const int stage = kt % KV_STAGES;
      mbar_wait(&kv_empty[stage], ((kt / KV_STAGES) & 1) ^ 1);
      // ---- one key per thread: where does key (kt*128 + lt) live, and what is its logit bias
      const int idx = kt * BN + lt;
      uint64_t p = 0;
      float bias = -INFINITY;
      if (idx < nB && idx >= nA) {                       // pooled window key
        const int j = idx - nA;
        const int t = j / npool, pp = j - t * npool;
        const int pi = pi0 + pp / PW, pj = pj0 + pp % PW;
        const size_t tok = ((static_cast<size_t>(b) * prm.T + t) * prm.nWh + pi) * prm.nWw + pj;
        p = reinterpret_cast<uint64_t>(prm.pooled + tok * C3 + prm.C + head * HD);
        bias = 0.f;
      } else if (idx < NK) {                             // ring key (single-listed first, multiply-listed last)
        int t, e;
        if (idx < nA) {
          t = idx / prm.n1;
          e = idx - t * prm.n1;
        } else {
          const int j = idx - nB;
          t = j / prm.n2;
          e = prm.n1 + (j - t * prm.n2);
        }
        const int pos = prm.ring_pos[e], mult = prm.ring_mult[e];
        const int er = pos / EW, ec = pos - er * EW;
        int y = (wi * prm.wh - prm.eh + er) % prm.H;
        int x = (wj * prm.ww - prm.ew + ec) % prm.W;
        y += (y < 0) ? prm.H : 0;
        x += (x < 0) ? prm.W : 0;
        const size_t tok = ((static_cast<size_t>(b) * prm.T + t) * prm.H + y) * prm.W + x;
        p = reinterpret_cast<uint64_t>(prm.qkv + tok * C3 + prm.C + head * HD);
        bias = (mult == 1) ? 0.f : log2f(static_cast<float>(mult));
      }
      key_ptr[stage * BN + lt] = p;
      key_bias[stage * BN + lt] = bias;
      loader_barrier();
      if (!(prm.debug & 2) || kt < KV_STAGES) {
        gather_rows(smem_u32(smem + Smem::K + stage * TILE), key_ptr + stage * BN, lwarp, lane, 0);
        cp_async_commit();
        gather_rows(smem_u32(smem + Smem::V + stage * TILE), key_ptr + stage * BN, lwarp, lane, prm.C);
        cp_async_commit();
      }
      cp_async_wait<1>();                               // K landed
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&k_full[stage]);
      cp_async_wait<0>();                               // V landed
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&v_full[stage]);
    }
  } else {
    // =================================================================== tcgen05 issuer
    if (lane == 0) {
      const uint32_t idesc_s = umma_idesc_f16(BM, BN, 0, 0);   // S = Q K^T, both K-major (d contiguous)
      const uint32_t idesc_o = umma_idesc_f16(BM, HD, 0, 1);   // O = P V, V is MN-major (d contiguous per key)
      const uint32_t sQ = smem_u32(smem + Smem::Q), sK = smem_u32(smem + Smem::K);
      const uint32_t sV = smem_u32(smem + Smem::V), sPa = smem_u32(smem + Smem::P);
      auto issue_s = [&](int kt) {
        const int stage = kt % KV_STAGES, sb = kt & 1;
        mbar_wait(&k_full[stage], (kt / KV_STAGES) & 1);
        mbar_wait(&s_free[sb], ((kt >> 1) & 1) ^ 1);
        tc_fence_after_sync();
        const uint32_t kb = sK + stage * TILE;
#pragma unroll
        for (int k = 0; k < HD / 16; ++k)
          if (!(prm.debug & 4) || kt < 2) umma_f16(tbase + COL_S + sb * BN, umma_desc_sw128(sQ + (k >> 2) * ATOM + (k & 3) * 32, 16, 1024),
                   umma_desc_sw128(kb + (k >> 2) * ATOM + (k & 3) * 32, 16, 1024), idesc_s, k != 0);
        umma_commit(&s_full[sb]);
      };
      mbar_wait(q_full, 0);
      issue_s(0);
      for (int kt = 0; kt < num_kt; ++kt) {
        const int stage = kt % KV_STAGES;
        if (kt + 1 < num_kt) issue_s(kt + 1);
        mbar_wait(p_full, kt & 1);
        mbar_wait(&v_full[stage], (kt / KV_STAGES) & 1);
        tc_fence_after_sync();
        const uint32_t vb = sV + stage * TILE;
#pragma unroll
        for (int k = 0; k < BN / 16; ++k)
          if (!(prm.debug & 4) || kt < 2) umma_f16(tbase + COL_O, umma_desc_sw128(sPa + (k >> 2) * ATOM + (k & 3) * 32, 16, 1024),
                   umma_desc_sw128(vb + k * 2048, ATOM, 1024), idesc_o, (kt | k) != 0);
        umma_commit(&kv_empty[stage]);
        umma_commit(o_done);
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == MMA_WARP) tmem_dealloc(tbase, TMEM_COLS);
}

}  // namespace attn

int launch_focal_attention(const void* qkv, const void* qkv_pooled, void* out, int b, int t, int h, int w, int heads,
                           int head_dim, int wh, int ww, int eh, int ew, int fh, int fw, int use_pooled, float scale,
                           int out_dtype, cudaStream_t stream) {
  using namespace attn;
  if (head_dim != HD) {
    set_error("focal attention: head_dim %d unsupported", head_dim);
    return -2;
  }
  if (b == 0) return 0;
  const int EH = wh + 2 * eh, EW = ww + 2 * ew;
  if (EH * EW > MAX_RING) {
    set_error("focal attention: expanded window %dx%d exceeds %d positions", EH, EW, MAX_RING);
    return -2;
  }
  if (!(scale > 0.f)) {
    set_error("focal attention: scale must be positive");
    return -1;
  }
  Params prm;
  prm.qkv = static_cast<const __half*>(qkv);
  prm.pooled = static_cast<const __half*>(qkv_pooled);
  prm.out = out;
  prm.B = b; prm.T = t; prm.H = h; prm.W = w; prm.heads = heads; prm.C = heads * HD;
  prm.wh = wh; prm.ww = ww; prm.eh = eh; prm.ew = ew; prm.fh = fh; prm.fw = fw;
  prm.nWh = h / wh; prm.nWw = w / ww;
  prm.use_pooled = use_pooled;
  prm.scale_log2 = scale * LOG2E;
  {
    const char* dbg = getenv("E2F_ATTN_DEBUG");
    prm.debug = dbg ? atoi(dbg) : 0;
  }
  // order the expanded-window positions: single-listed first, multiply-listed last (positions never listed are dropped)
  prm.n1 = prm.n2 = 0;
  for (int pass = 0; pass < 2; ++pass)
    for (int e = 0; e < EH * EW; ++e) {
      const int m = key_multiplicity(e / EW, e % EW, wh, ww, eh, ew);
      if ((pass == 0 && m == 1) || (pass == 1 && m > 1)) {
        const int slot = prm.n1 + prm.n2;
        prm.ring_pos[slot] = static_cast<uint8_t>(e);
        prm.ring_mult[slot] = static_cast<uint8_t>(m);
        (pass == 0 ? prm.n1 : prm.n2)++;
      }
    }
  const long long nwin = static_cast<long long>(b) * prm.nWh * prm.nWw;
  if (nwin > 65535 || heads > 65535) {
    set_error("focal attention: grid too large (B*nW=%lld)", nwin);
    return -2;
  }
  const dim3 grid((t * wh * ww + BM - 1) / BM, heads, static_cast<unsigned>(nwin));
  cudaError_t e;
  if (out_dtype == 1) {
    static bool cfg = false;
    if (!cfg) {
      e = cudaFuncSetAttribute(focal_attn_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
      if (e != cudaSuccess) return static_cast<int>(e);
      cfg = true;
    }
    focal_attn_kernel<__half><<<grid, THREADS, SMEM_BYTES, stream>>>(prm);
  } else {
    static bool cfg = false;
    if (!cfg) {
      e = cudaFuncSetAttribute(focal_attn_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
      if (e != cudaSuccess) return static_cast<int>(e);
      cfg = true;
    }
    focal_attn_kernel<float><<<grid, THREADS, SMEM_BYTES, stream>>>(prm);
  }
  count_launch();
  return static_cast<int>(cudaGetLastError());
}

}  // namespace e2f
