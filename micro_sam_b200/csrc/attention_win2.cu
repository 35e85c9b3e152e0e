// Windowed ViT attention (head_dim 64 / 80, 14x14 windows, 196 keys incl. the zero-pad tokens), second generation.
// Restates segment_anything's Attention.forward + add_decomposed_rel_pos (oracle/sam_ref.py:Attention) like attention.cu,
// but moves everything except max / exp off the CUDA cores (profiles/r1_ncu_attn_window_d80.txt: the first kernel spends
// ~10 instructions per logit and holds 28 bias values per thread; 168 registers, 30 % issue slots, 12.7 us per CTA):
//   * rel-pos bias inside the S MMA: T = Q R^T (as before) is shifted per query row, scaled by 1/scale and written as 28
//     extra bf16 K-columns next to Q (box 1, columns 16..43) plus their bf16 rounding residuals (a second tile over the dead
//     R table: hi + lo carries 16 mantissa bits, the bias stays at fp32-accumulator accuracy); the key tile gets the matching
//     one-hot columns E[k, kh(k)] = E[k, 14 + kw(k)] = 1, so S = Q K^T + (T/scale) E^T comes out of ONE accumulation chain
//     (9 k-steps instead of 5) and softmax(scale * S) needs no per-element bias arithmetic;
//   * row sums from the tensor core: V gets a ones column (box 1, column 16), O[:, 80] = sum_k P[:, k] of the bf16-rounded
//     probabilities (the same values the P V product sees);
//   * keys 192..195 (P covers 3 boxes = 192 keys) as a 13th k-step whose A tile lives in unused columns of the V tile
//     (box 1, columns 32..47), instead of 320 CUDA-core FMAs per row in the epilogue;
//   * row max with 3-input max (0.5 instructions / logit), exp pass = fma + ex2 + half a pack;
//   * TMEM loads double buffered (the next 32 columns are in flight while the current ones are processed).
// Layout: [box A0 16 KB | box A1 16 KB | R 16 KB | box B0 26 KB | box B1 26 KB]; P (keys 0..191, 3 boxes) goes over A0 | A1 | R
// once S is complete, V is loaded over K (B0 / B1).  head_dim 80: A0 | A1 = Q (80 of 128 columns), the bias columns sit in A1
// columns 16..47, the one-hot columns in B1 (= K box 1) columns 16..47, the ones column in V box 1 column 16.  head_dim 64:
// A0 = Q, A1 = the bias tile (columns 0..31), B0 = K / V, B1 = the one-hot tile (columns 0..31; column 0 becomes the ones
// column once S is complete): the same byte offsets, only the column offset inside box 1 differs.
// Warp roles: warps 0-3 softmax / epilogue (thread r <-> query row r <-> TMEM lane r), warp 4 TMA, warp 5 TMEM + MMA.
#include "kernels.h"
#include "ptx.cuh"
#include "tensormap.h"

namespace msam {

namespace {

constexpr int W8_THREADS = 192;
constexpr int W8_QBOX = 128 * 128;   // 128 rows x 64 bf16
constexpr int W8_NK = 208;           // keys padded to a multiple of 16
constexpr int W8_KBOX = W8_NK * 128;
constexpr int W8_RTBOX = 64 * 128;
constexpr int W8_OFF_RT = 2 * W8_QBOX;
constexpr int W8_OFF_K = W8_OFF_RT + 2 * W8_RTBOX;
constexpr int W8_OFF_BAR = W8_OFF_K + 2 * W8_KBOX;
constexpr int W8_SMEM = W8_OFF_BAR + 128 + 1024;
static_assert(W8_OFF_K % 1024 == 0 && W8_KBOX % 1024 == 0, "SW128 tiles must be 1024-byte aligned");

struct W8Params {
  __nv_bfloat16* out;
  int d_model, grid;
  float sl2;        // scale * log2(e)
  float inv_scale;  // 1 / scale
  unsigned long long* trace;
};

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// 16 columns into the low half of a 32-register buffer
__device__ __forceinline__ void tmem_ld16_lo(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
// x[i] <- x[i + sh] for i < 14, per-lane shift sh in [0, 14), static register indices only (x holds >= 29 valid values)
__device__ __forceinline__ void lane_shift14(float (&x)[32], int sh) {
#pragma unroll
  for (int i = 0; i < 21; ++i) x[i] = (sh & 8) ? x[i + 8] : x[i];
#pragma unroll
  for (int i = 0; i < 17; ++i) x[i] = (sh & 4) ? x[i + 4] : x[i];
#pragma unroll
  for (int i = 0; i < 15; ++i) x[i] = (sh & 2) ? x[i + 2] : x[i];
#pragma unroll
  for (int i = 0; i < 14; ++i) x[i] = (sh & 1) ? x[i + 1] : x[i];
}

#define W8_TRACE(slot) do { if (tr) tr[slot] = gtimer(); } while (0)

template <int D>
__global__ void __launch_bounds__(W8_THREADS, 2)
attn_window2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                     const __grid_constant__ CUtensorMap tmRT, const W8Params p) {
  constexpr int S = 14, G = 196;
  constexpr int NB = (D + 63) / 64;          // TMA boxes per operand
  constexpr int KS = D / 16;                 // k-steps of Q K^T
  constexpr int AUGC = (D == 80) ? 2 : 0;    // first 16-byte chunk of the bias / one-hot columns inside box 1
  constexpr int NO = D + 16;                 // P V columns: values + the ones column (+ 15 unused)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sRT = smem + W8_OFF_RT;
  uint8_t* sK = smem + W8_OFF_K;   // V later
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + W8_OFF_BAR);
  uint64_t *ld_full = bars, *v_full = bars + 1, *t_full = bars + 2, *t_done = bars + 3, *s_full = bars + 4,
           *p_full = bars + 5, *o_full = bars + 6, *q_full = bars + 7;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = blockIdx.x, head = blockIdx.y, group = blockIdx.z;
  unsigned long long* tr = nullptr;
  if (p.trace && threadIdx.x == 0 && qt == 0 && head == 0 && group < 64) tr = p.trace + group * 16;
  W8_TRACE(0);

  if (warp == 4 && lane == 0) {
    prefetch_tmap(&tmQ);
    prefetch_tmap(&tmKV);
    prefetch_tmap(&tmRT);
    mbar_init(ld_full, 1);
    mbar_init(q_full, 1);
    mbar_init(v_full, 1);
    mbar_init(t_full, 1);
    mbar_init(t_done, 128);
    mbar_init(s_full, 1);
    mbar_init(p_full, 128);
    mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == 5) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int row0 = group * G;
  pdl_wait();
  pdl_trigger();
  W8_TRACE(1);

  if (warp == 4) {
    if (lane == 0) {
      const int qcol = head * D, kcol = p.d_model + head * D, vcol = 2 * p.d_model + head * D;
      mbar_expect_tx(q_full, NB * (W8_QBOX + W8_RTBOX));   // T = Q R^T can start before K has landed
      for (int b = 0; b < NB; ++b) {
        tma_load_2d(sQ + b * W8_QBOX, &tmQ, q_full, qcol + b * 64, row0 + qt * 128);
        tma_load_2d(sRT + b * W8_RTBOX, &tmRT, q_full, b * 64, 0);
      }
      mbar_expect_tx(ld_full, NB * W8_KBOX);
      for (int b = 0; b < NB; ++b) tma_load_2d(sK + b * W8_KBOX, &tmKV, ld_full, kcol + b * 64, row0);
      mbar_wait(s_full, 0, 44);  // S has been computed: V goes over the dead K tile
      mbar_expect_tx(v_full, NB * W8_KBOX);
      for (int b = 0; b < NB; ++b) tma_load_2d(sK + b * W8_KBOX, &tmKV, v_full, vcol + b * 64, row0);
    }
  } else if (warp == 5) {
    constexpr uint32_t idescT = make_idesc_bf16(128, 64);
    constexpr uint32_t idescS = make_idesc_bf16(128, W8_NK);
    constexpr uint32_t idescO = make_idesc_bf16(128, NO, 1);   // D value columns + the ones column (+ 15 unused)
    const uint32_t aQ = smem_u32(sQ), aK = smem_u32(sK), aRT = smem_u32(sRT);
    auto kdesc = [](uint32_t base, uint32_t box_bytes, int ks) {
      return make_desc_sw128(base + (uint32_t)(ks >> 2) * box_bytes + (uint32_t)(ks & 3) * 32u, 0, 1024);
    };
    mbar_wait(q_full, 0, 40);
    tc_fence_after();
    if (elect_one()) {
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) umma_bf16(tmem, kdesc(aQ, W8_QBOX, ks), kdesc(aRT, W8_RTBOX, ks), idescT, ks > 0);
      umma_commit(t_full);
    }
    __syncwarp();
    mbar_wait(t_done, 0, 41);  // T is in registers, the bias (hi next to Q, lo over R) / one-hot columns are in shared memory
    mbar_wait(ld_full, 0, 45);
    tc_fence_after();
    if (elect_one()) {
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) umma_bf16(tmem, kdesc(aQ, W8_QBOX, ks), kdesc(aK, W8_KBOX, ks), idescS, ks > 0);
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {   // bias (hi, then its rounding residuals) against the one-hot columns of box 1
        const uint64_t de = make_desc_sw128(aK + W8_KBOX + (uint32_t)(AUGC * 16 + ks * 32), 0, 1024);
        umma_bf16(tmem, make_desc_sw128(aQ + W8_QBOX + (uint32_t)(AUGC * 16 + ks * 32), 0, 1024), de, idescS, 1);
        umma_bf16(tmem, make_desc_sw128(aRT + (uint32_t)ks * 32u, 0, 1024), de, idescS, 1);
      }
      umma_commit(s_full);
    }
    __syncwarp();
    mbar_wait(p_full, 0, 42);  // P written (over Q | R), ones column in V, S fully consumed
    mbar_wait(v_full, 0, 43);
    tc_fence_after();
    if (elect_one()) {
#pragma unroll
      for (int ks = 0; ks < 12; ++ks) {
        const uint64_t da = make_desc_sw128(aQ + (uint32_t)(ks >> 2) * W8_QBOX + (uint32_t)(ks & 3) * 32u, 0, 1024);
        const uint64_t db = make_desc_sw128(aK + (uint32_t)ks * 2048u, W8_KBOX, 1024);
        umma_bf16(tmem, da, db, idescO, ks > 0);   // O over the dead S columns [0, NO)
      }
      // keys 192..207: A = the tail probabilities parked in V box 1 columns 32..47, B = V rows 192..207
      umma_bf16(tmem, make_desc_sw128(aK + W8_KBOX + 64u, 0, 1024), make_desc_sw128(aK + 12u * 2048u, W8_KBOX, 1024), idescO, 1);
      umma_commit(o_full);
    }
    __syncwarp();
  } else {
    const int r = threadIdx.x;
    const uint32_t tlane = tmem + ((uint32_t)(warp * 32) << 16);
    const int qi = qt * 128 + r;
    const bool live = (qt * 128 + warp * 32) < G;   // warp-uniform: this warp holds at least one real query row
    const uint32_t aQ = smem_u32(sQ), aK = smem_u32(sK);
    mbar_wait(t_full, 0, 50);
    tc_fence_after();
    W8_TRACE(2);
    if (live) {
      int qh = qi / S;
      const int qw = qi - qh * S;
      if (qh > S - 1) qh = S - 1;
      uint32_t w[16], wl[16];
      uint32_t v[32];
      float x[32];
      // hi = bf16(y), lo = bf16(y - hi): packs a pair of each
      auto split = [](float a, float b, uint32_t& hi, uint32_t& lo) {
        hi = pack_bf16(a, b);
        lo = pack_bf16(a - __uint_as_float(hi << 16), b - __uint_as_float(hi & 0xffff0000u));
      };
      tmem_ld32(tlane + 0, v);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) x[i] = __uint_as_float(v[i]) * p.inv_scale;
      tmem_ld32(tlane + 32, v);       // in flight during the shift
      lane_shift14(x, qh);   // x[i] = T[q, i + qh]
#pragma unroll
      for (int j = 0; j < 7; ++j) split(x[S - 1 - 2 * j], x[S - 2 - 2 * j], w[j], wl[j]);   // yh[kh] = x[13 - kh]
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) x[i] = __uint_as_float(v[i]) * p.inv_scale;
      lane_shift14(x, qw);
#pragma unroll
      for (int j = 0; j < 7; ++j) split(x[S - 1 - 2 * j], x[S - 2 - 2 * j], w[7 + j], wl[7 + j]);
      w[14] = 0u; w[15] = 0u; wl[14] = 0u; wl[15] = 0u;
      const uint32_t qrow = aQ + W8_QBOX + (uint32_t)r * 128u;          // hi: Q box 1, columns 16..47
      const uint32_t lrow = smem_u32(sRT) + (uint32_t)r * 128u;         // lo: its own tile over the dead R table, columns 0..31
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        st_shared_v4(qrow + ((uint32_t)((c + AUGC) ^ (r & 7)) << 4), make_uint4(w[4 * c], w[4 * c + 1], w[4 * c + 2], w[4 * c + 3]));
        st_shared_v4(lrow + ((uint32_t)(c ^ (r & 7)) << 4), make_uint4(wl[4 * c], wl[4 * c + 1], wl[4 * c + 2], wl[4 * c + 3]));
      }
    }
    // one-hot key columns: rows r and r + 128 of box B1.  head_dim 80: B1 is K box 1 -- wait until the K tile has landed (its
    // TMA box covers these columns with the next head's data); head_dim 64: B1 is a tile of its own
    if constexpr (D == 80) mbar_wait(ld_full, 0, 54);
    for (int k = r; k < W8_NK; k += 128) {
      unsigned long long bits = 0ull;
      if (k < G) {
        const int kh = k / S, kw = k - kh * S;
        bits = (1ull << kh) | (1ull << (S + kw));
      }
      const uint32_t krow = aK + W8_KBOX + (uint32_t)k * 128u;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t e[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int col = 2 * (4 * c + j);
          e[j] = (((bits >> col) & 1ull) ? 0x3F80u : 0u) | (((bits >> (col + 1)) & 1ull) ? 0x3F800000u : 0u);
        }
        st_shared_v4(krow + ((uint32_t)((c + AUGC) ^ (k & 7)) << 4), make_uint4(e[0], e[1], e[2], e[3]));
      }
    }
    fence_proxy_async_smem();
    tc_fence_before();
    mbar_arrive(t_done);
    W8_TRACE(3);

    mbar_wait(s_full, 0, 51);
    tc_fence_after();
    W8_TRACE(4);
    uint32_t ptail[2] = {0u, 0u};   // probabilities of keys 192..195 (bf16 pairs), A operand of the 13th P V k-step
    if (live) {
      uint32_t va[32], vb[32];
      float m = -INFINITY, nmsl2 = 0.f;
      const float sl2 = p.sl2;
      tmem_ld32(tlane, va);
      // 14 steps: chunks 0..6 for the row max, then chunks 0..6 again for exp; the next chunk is always in flight
#pragma unroll
      for (int s = 0; s < 14; ++s) {
        const int c = s % 7, cn = (s + 1) % 7;
        uint32_t(&cur)[32] = (s & 1) ? vb : va;
        uint32_t(&nxt)[32] = (s & 1) ? va : vb;
        tmem_ld_wait();
        if (s + 1 < 14) {
          if (cn < 6) tmem_ld32(tlane + cn * 32, nxt);
          else tmem_ld16_lo(tlane + 192, nxt);
        }
        if (s < 7) {
          if (c < 6) {
#pragma unroll
            for (int i = 0; i < 32; i += 2) m = fmax3(m, __uint_as_float(cur[i]), __uint_as_float(cur[i + 1]));
          } else {
            m = fmax3(m, __uint_as_float(cur[0]), __uint_as_float(cur[1]));
            m = fmax3(m, __uint_as_float(cur[2]), __uint_as_float(cur[3]));
            nmsl2 = -m * sl2;
            W8_TRACE(5);
          }
        } else if (c < 6) {
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 32; i += 2)
            pk[i >> 1] = pack_bf16(ex2_approx(fmaf(__uint_as_float(cur[i]), sl2, nmsl2)),
                                   ex2_approx(fmaf(__uint_as_float(cur[i + 1]), sl2, nmsl2)));
          const uint32_t prow = aQ + (uint32_t)(c >> 1) * W8_QBOX + (uint32_t)r * 128u;
#pragma unroll
          for (int q = 0; q < 4; ++q)
            st_shared_v4(prow + ((uint32_t)(((c & 1) * 4 + q) ^ (r & 7)) << 4),
                         make_uint4(pk[q * 4], pk[q * 4 + 1], pk[q * 4 + 2], pk[q * 4 + 3]));
        } else {
          ptail[0] = pack_bf16(ex2_approx(fmaf(__uint_as_float(cur[0]), sl2, nmsl2)), ex2_approx(fmaf(__uint_as_float(cur[1]), sl2, nmsl2)));
          ptail[1] = pack_bf16(ex2_approx(fmaf(__uint_as_float(cur[2]), sl2, nmsl2)), ex2_approx(fmaf(__uint_as_float(cur[3]), sl2, nmsl2)));
        }
      }
    }
    W8_TRACE(6);
    // V has landed: ones column (box 1, column 16) for all key rows, and this row's tail probabilities (keys 192..207, zero
    // beyond 195) into box 1 columns 32..47 = the K-major A tile of the 13th k-step
    mbar_wait(v_full, 0, 53);
    for (int k = r; k < W8_NK; k += 128) {
      const uint32_t a = aK + W8_KBOX + (uint32_t)k * 128u + ((uint32_t)(AUGC ^ (k & 7)) << 4);
      asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "h"((unsigned short)0x3F80) : "memory");
    }
    {
      const uint32_t a = aK + W8_KBOX + (uint32_t)r * 128u;
      st_shared_v4(a + ((uint32_t)(4 ^ (r & 7)) << 4), make_uint4(ptail[0], ptail[1], 0u, 0u));
      st_shared_v4(a + ((uint32_t)(5 ^ (r & 7)) << 4), make_uint4(0u, 0u, 0u, 0u));
    }
    tc_fence_before();
    fence_proxy_async_smem();
    mbar_arrive(p_full);
    W8_TRACE(7);

    mbar_wait(o_full, 0, 52);
    tc_fence_after();
    W8_TRACE(8);
    long out_row = -1;
    {
      const int wpr = (p.grid + S - 1) / S;
      const int b = group / (wpr * wpr), wy = (group / wpr) % wpr, wx = group % wpr;
      const int y = wy * S + qi / S, x = wx * S + qi % S;
      if (qi < G && y < p.grid && x < p.grid) out_row = (long)b * p.grid * p.grid + y * p.grid + x;
    }
    if (live) {
      __nv_bfloat16* orow = p.out + (out_row < 0 ? 0 : out_row) * p.d_model + head * D;
      uint32_t va[32], vb[32];
      tmem_ld32(tlane + 64, va);   // head_dim 80: 16 value columns + the row sum at column 80; head_dim 64: the row sum at column 64
      tmem_ld_wait();
      tmem_ld32(tlane + 0, vb);
      const float inv = 1.0f / __uint_as_float(va[D - 64]);
      auto store = [&](const uint32_t(&v)[32], int c0, int n) {
        if (out_row < 0) return;
#pragma unroll
        for (int cc = 0; cc < 32; cc += 8) {
          if (cc >= n) break;
          uint4 u;
          u.x = pack_bf16(__uint_as_float(v[cc + 0]) * inv, __uint_as_float(v[cc + 1]) * inv);
          u.y = pack_bf16(__uint_as_float(v[cc + 2]) * inv, __uint_as_float(v[cc + 3]) * inv);
          u.z = pack_bf16(__uint_as_float(v[cc + 4]) * inv, __uint_as_float(v[cc + 5]) * inv);
          u.w = pack_bf16(__uint_as_float(v[cc + 6]) * inv, __uint_as_float(v[cc + 7]) * inv);
          *reinterpret_cast<uint4*>(orow + c0 + cc) = u;
        }
      };
      if constexpr (D == 80) store(va, 64, 16);
      tmem_ld_wait();
      tmem_ld32(tlane + 32, va);
      store(vb, 0, 32);
      tmem_ld_wait();
      store(va, 32, 32);
    }
    W8_TRACE(9);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc(tmem, 256);
  }
}

unsigned long long* g_attn_trace = nullptr;

}  // namespace

void set_attn_trace(unsigned long long* dev_buf) { g_attn_trace = dev_buf; }
unsigned long long* get_attn_trace() { return g_attn_trace; }

template <int D>
static int launch_attn_window2_t(const AttnArgs& a, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attn_window2_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, W8_SMEM);
    if (e != cudaSuccess) return set_error("attention: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  constexpr int NB = (D + 63) / 64;
  const int d_model = a.heads * D, S = 14;
  const int wpr = (a.grid + S - 1) / S;
  const int groups = a.batch * wpr * wpr;
  const long rows = (long)groups * 196;
  CUtensorMap tmQ, tmKV, tmRT;
  if (make_tmap_bf16_2d(&tmQ, a.qkv, rows, 3 * d_model, 3 * d_model, 128)) return -1;
  if (make_tmap_bf16_2d(&tmKV, a.qkv, rows, 3 * d_model, 3 * d_model, W8_NK)) return -1;
  if (make_tmap_bf16_2d(&tmRT, a.rel_table, 64, NB * 64, NB * 64, 64)) return -1;
  W8Params p;
  p.out = a.out; p.d_model = d_model; p.grid = a.grid; p.sl2 = a.scale * 1.4426950408889634f; p.inv_scale = 1.0f / a.scale;
  p.trace = g_attn_trace;
  prof_begin(stream, D == 64 ? "attn_window<64>" : "attn_window<80>",
             (double)groups * a.heads * (4.0 * 196 * 196 * D + 4.0 * 196 * S * D), (double)groups * 196 * a.heads * D * 2 * 4);
  launch_pdl(attn_window2_kernel<D>, dim3(2, a.heads, groups), dim3(W8_THREADS), W8_SMEM, stream, tmQ, tmKV, tmRT, p);
  prof_end(stream);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error("window attention launch failed: %s", cudaGetErrorString(e));
  count_launch();
  return 0;
}

int launch_attn_window2(const AttnArgs& a, cudaStream_t stream) {
  return a.head_dim == 64 ? launch_attn_window2_t<64>(a, stream) : launch_attn_window2_t<80>(a, stream);
}

}  // namespace msam
