// Persistent warp-specialised bf16 GEMM for sm_100a:  C[M,N] = epilogue(A[M,K] * W[N,K]^T)
//   * operands: TMA (SWIZZLE_128B boxes of 64 K-elements) -> shared memory ring (mbarrier full/empty)
//   * math:     tcgen05.mma cta_group::1 kind::f16, M=128 x N=BN x K=16, fp32 accumulators in TMEM (2 stages)
//   * epilogue: NG column groups x 4 warps (warp%4 = TMEM lane quadrant), tcgen05.ld 32x32b -> registers ->
//               bias / GELU(erf) / ReLU / residual (fp32 or bf16, row modulus) -> bf16 or fp32, or one of the fused
//               row epilogues: LayerNorm over N=256 (EPI_LN256), LayerNorm over 64-column groups + GELU (EPI_LN64_GELU),
//               GELU + hyper-network mask product (EPI_HYPER).  Residual / bias loads are issued BEFORE the wait on the
//               accumulator barrier so their latency hides behind the MMA of the same tile.
// One CTA per SM; tiles are distributed round-robin, N-block fastest so that the CTAs that run concurrently share the
// same A row-block through L2.
//
// Replaces (on the B200 path) the nn.Linear / 1x1-conv / conv-transpose / patch-embed conv calls inside
// segment_anything's ImageEncoderViT / MaskDecoder that micro_sam reaches through util.py:674 (image_encoder) and
// SamPredictor.predict_torch (inference.py:248, instance_segmentation.py:361).
#include "kernels.h"
#include "ptx.cuh"
#include "tensormap.h"
#include <cstdlib>

namespace msam {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;

enum { EPI_PLAIN = 0, EPI_LN256 = 1, EPI_LN64_GELU = 2, EPI_HYPER = 3 };

// MT = 2: the CTA computes a 256 x BN tile as two M=128 MMAs per K step that share the weight tile in shared memory.
// The 1-CTA 128x256 tile pulls 48 KB per 128x256x64 MMA block; at the measured chip-wide L2 throughput (~6300 B/clk,
// 42.6 B/clk/SM) that caps the tensor pipe at ~45 % -- exactly what the MT = 1 kernel reached (1.0 PFLOP/s; a 2-CTA
// cluster with TMA weight multicast did not move it).  MT = 2 needs 32 KB per block-equivalent (cap ~68 %), at the
// price of a single (not double-buffered) accumulator stage.
template <int BN, int MT = 1>
struct GemmCfg {
  static constexpr int NG = (BN >= 128) ? 4 : 2;        // epilogue column groups
  static constexpr int CPW = BN / NG;                   // columns per epilogue thread (64 / 32 / 32)
  static constexpr int THREADS = 128 + NG * 128;        // warp0 TMA, warp1 MMA, warp2 TMEM alloc, warp3 idle, epilogue
  static constexpr int STAGES = (MT == 2) ? 3 : ((BN == 256) ? 4 : 6);
  static constexpr int A_BYTES = MT * GEMM_BM * GEMM_BK * 2;
  static constexpr int B_BYTES = BN * GEMM_BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int OFF_BAR = STAGES * STAGE_BYTES;
  static constexpr int SMEM_BYTES = OFF_BAR + 256 + 1024 /*align slack*/;   // + 11 KB static (exch, rowp)
  static constexpr int TMEM_COLS = 2 * BN;              // 2 accumulator stages (MT=1) or 2 row halves (MT=2)
  static constexpr int ACC_STAGES = (MT == 2) ? 1 : 2;
  static_assert(MT == 1 || BN == 256, "MT = 2 is only instantiated for BN = 256");
};

struct GemmParams {
  int M, N, K;
  const float* bias;      // [N] or null
  const void* residual;   // fp32 (or bf16 if res_bf16) [res_rows, ldr] or null; row index = row % res_rows
  int res_rows, ldr, res_bf16;
  void* out;              // bf16 or fp32 [M, ldc]
  int ldc;
  int out_fp32;
  int act;                // 0 none, 1 GELU(erf), 2 ReLU
  const float* ln_gamma;  // fused LayerNorm epilogues
  const float* ln_beta;
  float ln_eps;
  const float* hyper;     // EPI_HYPER: [P, 4, 32] hyper-network outputs
  int hyper_m0, hyper_nm;
};

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
// GELU(x) = 0.5 x (1 + erf(x / sqrt 2)) with erf from Abramowitz-Stegun 7.1.26 (|abs err| < 1.5e-7, far below the bf16
// rounding of the value that is stored): 1 MUFU.RCP + 1 MUFU.EX2 + 7 FMA instead of ~25 instructions of erff().
__device__ __forceinline__ float gelu_fast(float x) {
  const float z = fabsf(x) * 0.70710678118654752440f;
  const float t = __frcp_rn(fmaf(0.3275911f, z, 1.0f));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  const float e = 1.0f - poly * t * exp2f(-z * z * 1.4426950408889634f);  // erf(|x|/sqrt2)
  return 0.5f * x + 0.5f * fabsf(x) * e;                                   // 0.5 x (1 + sign(x) e)
}
__device__ __forceinline__ float apply_act(float x, int act) {
  if (act == 1) return gelu_erf(x);
  if (act == 2) return fmaxf(x, 0.0f);
  return x;
}
__device__ __forceinline__ void bf16x8_to_f32(const uint4& r, float* f) {
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    f[2 * q] = __uint_as_float(w[q] << 16);
    f[2 * q + 1] = __uint_as_float(w[q] & 0xffff0000u);
  }
}

template <int BN, int EPI, int MT>
__global__ void __launch_bounds__(GemmCfg<BN, MT>::THREADS, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
  using Cfg = GemmCfg<BN, MT>;
  constexpr int TILE_M = MT * GEMM_BM;
  constexpr int NG = Cfg::NG, CPW = Cfg::CPW, NCH = CPW / 32;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
  uint64_t* empty_bar = full_bar + Cfg::STAGES;
  uint64_t* tfull_bar = empty_bar + Cfg::STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  // statically declared so that the compiler emits LDS/STS (not generic LD/ST through the LG path)
  __shared__ __align__(16) float2 exch[2 * 4 * 128];  // EPI_LN256 statistics exchange, double buffered by tile parity
  __shared__ __align__(16) float rowp[768];           // [0,256) bias, [256,512) gamma, [512,768) beta (fused epilogues)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m_blocks = (p.M + TILE_M - 1) / TILE_M;
  const int n_blocks = (p.N + BN - 1) / BN;
  const int k_blocks = (p.K + GEMM_BK - 1) / GEMM_BK;
  const int num_tiles = m_blocks * n_blocks;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < Cfg::STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], NG * 4);  // one arrive per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  if constexpr (EPI != EPI_PLAIN) {  // N == BN: per-column parameters are tile-invariant -> shared memory
    for (int i = threadIdx.x; i < BN; i += Cfg::THREADS) {
      rowp[i] = p.bias ? p.bias[i] : 0.f;
      if constexpr (EPI == EPI_LN256) { rowp[256 + i] = p.ln_gamma[i]; rowp[512 + i] = p.ln_beta[i]; }
      if constexpr (EPI == EPI_LN64_GELU) { rowp[256 + i] = p.ln_gamma[i & 63]; rowp[512 + i] = p.ln_beta[i & 63]; }
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_blk = tile / n_blocks, n_blk = tile % n_blocks;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1, 1);
          uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
          uint8_t* sb = sa + Cfg::A_BYTES;
          mbar_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
          tma_load_2d(sa, &tmA, &full_bar[stage], kb * GEMM_BK, m_blk * TILE_M);  // one box of TILE_M rows
          tma_load_2d(sb, &tmB, &full_bar[stage], kb * GEMM_BK, n_blk * BN);
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (single thread)
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(GEMM_BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int as = (Cfg::ACC_STAGES == 2) ? (it & 1) : 0;
        const uint32_t aphase = (Cfg::ACC_STAGES == 2) ? ((it >> 1) & 1) : (it & 1);
        mbar_wait(&tempty_bar[as], aphase ^ 1, 2);  // epilogue drained this accumulator stage
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * BN;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase, 3);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
          const uint32_t sb = sa + Cfg::A_BYTES;
          const uint64_t da = make_desc_sw128(sa, 0, 1024);
          const uint64_t db = make_desc_sw128(sb, 0, 1024);
#pragma unroll
          for (int k = 0; k < GEMM_BK / 16; ++k) {
            // +32 B per K=16 step inside the 128-B swizzle atom -> +2 on the encoded (>>4) start address
            umma_bf16(tmem_d, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb | k) != 0);
            if constexpr (MT == 2)  // rows 128..255 of the tile: A box + 16 KB (encoded +1024), accumulator columns + BN
              umma_bf16(tmem_d + BN, da + (uint64_t)(1024 + 2 * k), db + (uint64_t)(2 * k), idesc, (kb | k) != 0);
          }
          umma_commit(&empty_bar[stage]);  // smem slot free once these MMAs have read it
          if (kb == k_blocks - 1) umma_commit(&tfull_bar[as]);
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------ epilogue: TMEM -> regs -> (fused op) -> global
    const int quad = warp & 3;           // TMEM lane quadrant this warp may access
    const int grp = (warp - 4) >> 2;     // column group of the tile
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int m_blk = tile / n_blocks, n_blk = tile % n_blocks;
      const int as = (Cfg::ACC_STAGES == 2) ? (it & 1) : 0;
      const uint32_t aphase = (Cfg::ACC_STAGES == 2) ? ((it >> 1) & 1) : (it & 1);
      const int colbase = n_blk * BN + grp * CPW;
#pragma unroll 1
      for (int mh = 0; mh < MT; ++mh) {
      const int row = m_blk * TILE_M + mh * GEMM_BM + quad * 32 + lane;
      const bool row_ok = row < p.M;
      const uint32_t tcol = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)((MT == 2 ? mh : as) * BN + grp * CPW);
      const bool last_half = (mh == MT - 1);

      if constexpr (EPI == EPI_PLAIN) {
        const bool active = row_ok && colbase < p.N;  // N % 32 == 0 and CPW % 32 == 0: whole chunks are in or out
        const float* res_f = nullptr;
        const __nv_bfloat16* res_b = nullptr;
        if (p.residual && active) {
          const size_t off = (size_t)(row % p.res_rows) * p.ldr + colbase;
          if (p.res_bf16) res_b = reinterpret_cast<const __nv_bfloat16*>(p.residual) + off;
          else res_f = reinterpret_cast<const float*>(p.residual) + off;
        }
        // pull this thread's residual row segment towards L1 while the MMA of the tile is still running
        if (res_f) {
#pragma unroll
          for (int b = 0; b < CPW * 4; b += 128) prefetch_l1(reinterpret_cast<const char*>(res_f) + b);
        } else if (res_b) {
#pragma unroll
          for (int b = 0; b < CPW * 2; b += 128) prefetch_l1(reinterpret_cast<const char*>(res_b) + b);
        }
        if (mh == 0) {
          mbar_wait(&tfull_bar[as], aphase, 4);
          tc_fence_after();
        }
#pragma unroll 1
        for (int c = 0; c < NCH; ++c) {
          const int col0 = colbase + c * 32;
          const bool on = active && col0 < p.N;
          uint32_t v[32];
          tmem_ld32(tcol + c * 32, v);
          tmem_ld_wait();
          if (c == NCH - 1 && last_half) {  // this warp's reads of the accumulator are complete -> back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[as]);
          }
          if (on) {
            float f[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
            if (p.bias) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + j));
                f[j] += b.x; f[j + 1] += b.y; f[j + 2] += b.z; f[j + 3] += b.w;
              }
            }
            if (p.act) {
#pragma unroll
              for (int j = 0; j < 32; ++j) f[j] = apply_act(f[j], p.act);
            }
            if (res_f) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const float4 r = *reinterpret_cast<const float4*>(res_f + c * 32 + j);
                f[j] += r.x; f[j + 1] += r.y; f[j + 2] += r.z; f[j + 3] += r.w;
              }
            } else if (res_b) {
#pragma unroll
              for (int j = 0; j < 32; j += 8) {
                float r8[8];
                bf16x8_to_f32(*reinterpret_cast<const uint4*>(res_b + c * 32 + j), r8);
#pragma unroll
                for (int q = 0; q < 8; ++q) f[j + q] += r8[q];
              }
            }
            if (p.out_fp32) {
              float* o = reinterpret_cast<float*>(p.out) + (size_t)row * p.ldc + col0;
#pragma unroll
              for (int j = 0; j < 32; j += 4)
                *reinterpret_cast<float4*>(o + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
            } else {
              __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + (size_t)row * p.ldc + col0;
#pragma unroll
              for (int j = 0; j < 32; j += 8)
                *reinterpret_cast<uint4*>(o + j) = make_uint4(pack_bf16(f[j], f[j + 1]), pack_bf16(f[j + 2], f[j + 3]),
                                                              pack_bf16(f[j + 4], f[j + 5]), pack_bf16(f[j + 6], f[j + 7]));
            }
          }
          __syncwarp();  // reconverge before the next warp-collective tcgen05.ld
        }
      } else {
        // ---- fused row epilogues (N == BN): this thread's CPW columns of the row live in registers
        float f[CPW];
        // residual (bf16 only in these modes): pulled towards L1 before the accumulator is ready, read after it
        const bool has_res = (EPI == EPI_LN256) && p.residual != nullptr && row_ok;
        const __nv_bfloat16* rp = nullptr;
        if (has_res) {
          rp = reinterpret_cast<const __nv_bfloat16*>(p.residual) + (size_t)(row % p.res_rows) * p.ldr + colbase;
          prefetch_l1(rp);
        }
        mbar_wait(&tfull_bar[as], aphase, 4);
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          uint32_t v[32];
          tmem_ld32(tcol + c * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) f[c * 32 + j] = __uint_as_float(v[j]) + rowp[grp * CPW + c * 32 + j];
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty_bar[as]);

        if constexpr (EPI == EPI_LN256) {
          if (has_res) {
#pragma unroll
            for (int j = 0; j < CPW / 8; ++j) {
              float r8[8];
              bf16x8_to_f32(*reinterpret_cast<const uint4*>(rp + 8 * j), r8);
#pragma unroll
              for (int q = 0; q < 8; ++q) f[8 * j + q] += r8[q];
            }
          }
          // LayerNorm over the full 256-wide row: exact two-pass statistics per 64-column group, combined across the four
          // groups with Chan's formula through shared memory.
          float s4[4] = {0.f, 0.f, 0.f, 0.f};  // 4 independent chains
#pragma unroll
          for (int j = 0; j < CPW; ++j) s4[j & 3] += f[j];
          const float mean_g = ((s4[0] + s4[1]) + (s4[2] + s4[3])) * (1.0f / CPW);
          float q4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int j = 0; j < CPW; ++j) { const float d = f[j] - mean_g; q4[j & 3] = fmaf(d, d, q4[j & 3]); }
          const float m2_g = (q4[0] + q4[1]) + (q4[2] + q4[3]);
          float2* ex = exch + (it & 1) * (NG * 128);
          ex[grp * 128 + quad * 32 + lane] = make_float2(mean_g, m2_g);
          asm volatile("bar.sync 1, %0;" ::"n"(NG * 128) : "memory");
          float mean = 0.f;
          float2 st[NG];
#pragma unroll
          for (int g = 0; g < NG; ++g) { st[g] = ex[g * 128 + quad * 32 + lane]; mean += st[g].x; }
          mean *= (1.0f / NG);
          float m2 = 0.f;
#pragma unroll
          for (int g = 0; g < NG; ++g) { const float d = st[g].x - mean; m2 += st[g].y + d * d * CPW; }
          const float rstd = rsqrtf(m2 * (1.0f / (NG * CPW)) + p.ln_eps);
          if (row_ok) {
            __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + (size_t)row * p.ldc + colbase;
#pragma unroll
            for (int j = 0; j < CPW; j += 8) {
              float y[8];
#pragma unroll
              for (int q = 0; q < 8; ++q)
                y[q] = (f[j + q] - mean) * rstd * rowp[256 + grp * CPW + j + q] + rowp[512 + grp * CPW + j + q];
              *reinterpret_cast<uint4*>(o + j) =
                  make_uint4(pack_bf16(y[0], y[1]), pack_bf16(y[2], y[3]), pack_bf16(y[4], y[5]), pack_bf16(y[6], y[7]));
            }
          }
        } else if constexpr (EPI == EPI_LN64_GELU) {
          // LayerNorm2d over one 64-channel group (= one conv-transpose sub-pixel) + exact GELU; CPW == 64
          static_assert(EPI != EPI_LN64_GELU || CPW == 64, "one group per thread");
          if (row_ok) {
            float s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int j = 0; j < CPW; ++j) s4[j & 3] += f[j];
            const float mean = ((s4[0] + s4[1]) + (s4[2] + s4[3])) * (1.0f / CPW);
            float q4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int j = 0; j < CPW; ++j) { const float d = f[j] - mean; q4[j & 3] = fmaf(d, d, q4[j & 3]); }
            const float m2 = (q4[0] + q4[1]) + (q4[2] + q4[3]);
            const float rstd = rsqrtf(m2 * (1.0f / CPW) + p.ln_eps);
            __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + (size_t)row * p.ldc + colbase;
#pragma unroll
            for (int j = 0; j < CPW; j += 8) {
              float y[8];
#pragma unroll
              for (int q = 0; q < 8; ++q)
                y[q] = gelu_fast((f[j + q] - mean) * rstd * rowp[256 + j + q] + rowp[512 + j + q]);
              *reinterpret_cast<uint4*>(o + j) =
                  make_uint4(pack_bf16(y[0], y[1]), pack_bf16(y[2], y[3]), pack_bf16(y[4], y[5]), pack_bf16(y[6], y[7]));
            }
          }
        } else if constexpr (EPI == EPI_HYPER) {
          // second conv-transpose (N = 128 = 4 sub-sub-pixels x 32 channels; bias added above) + GELU + hyper product:
          // masks[p, mi, Y, X] = sum_ch hyper[p, m0+mi, ch] * gelu(up[row, ss*32 + ch]); CPW == 32: thread = one (row, ss).
          // GEMM row = (prompt p, token (y,x), sub-pixel (dy,dx)); ss = grp = ey*2 + ex.
          static_assert(EPI != EPI_HYPER || CPW == 32, "one sub-sub-pixel per thread");
          if (row_ok) {
#pragma unroll
            for (int j = 0; j < CPW; ++j) f[j] = gelu_fast(f[j]);
            const int sub = row & 3, tok = (row >> 2) & 4095, pp = row >> 14;
            const int Y = 4 * (tok >> 6) + 2 * (sub >> 1) + (grp >> 1), X = 4 * (tok & 63) + 2 * (sub & 1) + (grp & 1);
            for (int mi = 0; mi < p.hyper_nm; ++mi) {
              const float* hw = p.hyper + ((size_t)pp * 4 + p.hyper_m0 + mi) * 32;
              float a = 0.f;
#pragma unroll
              for (int c = 0; c < 32; c += 4) {
                const float4 h4 = __ldg(reinterpret_cast<const float4*>(hw + c));
                a += h4.x * f[c] + h4.y * f[c + 1] + h4.z * f[c + 2] + h4.w * f[c + 3];
              }
              reinterpret_cast<float*>(p.out)[(((size_t)pp * p.hyper_nm + mi) * 256 + Y) * 256 + X] = a;
            }
          }
        }
        __syncwarp();
      }
      }  // mh
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

template <int BN, int EPI, int MT = 1>
static int launch_gemm_bn(const GemmArgs& a, int num_sms, cudaStream_t stream) {
  using Cfg = GemmCfg<BN, MT>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e =
        cudaFuncSetAttribute(gemm_bf16_kernel<BN, EPI, MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return set_error("gemm: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  CUtensorMap tmA, tmB;
  if (make_tmap_bf16_2d(&tmA, a.A, a.M, a.K, a.lda, MT * GEMM_BM)) return -1;
  if (make_tmap_bf16_2d(&tmB, a.W, a.N, a.K, a.ldw, BN)) return -1;
  GemmParams p;
  p.M = a.M; p.N = a.N; p.K = a.K;
  p.bias = a.bias;
  p.residual = a.residual;
  p.res_bf16 = a.res_bf16;
  p.res_rows = a.res_rows > 0 ? a.res_rows : a.M;
  p.ldr = a.ldr > 0 ? a.ldr : a.N;
  p.out = a.out;
  p.ldc = a.ldc > 0 ? a.ldc : a.N;
  p.out_fp32 = a.out_fp32;
  p.act = a.act;
  p.ln_gamma = a.ln_gamma; p.ln_beta = a.ln_beta; p.ln_eps = a.ln_eps;
  p.hyper = a.hyper; p.hyper_m0 = a.hyper_m0; p.hyper_nm = a.hyper_nm;
  const int tiles = ((a.M + MT * GEMM_BM - 1) / (MT * GEMM_BM)) * ((a.N + BN - 1) / BN);
  const int grid = tiles < num_sms ? tiles : num_sms;
  prof_begin(stream, PROF_GEMM, 2.0 * a.M * a.N * a.K);
  gemm_bf16_kernel<BN, EPI, MT><<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, stream>>>(tmA, tmB, p);
  prof_end(stream);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error("gemm launch failed: %s", cudaGetErrorString(e));
  count_launch();
  return 0;
}

int launch_gemm(const GemmArgs& a, int num_sms, cudaStream_t stream) {
  if (a.M <= 0 || a.N <= 0 || a.K <= 0) return set_error("gemm: empty problem M=%d N=%d K=%d", a.M, a.N, a.K);
  if (a.N % 32 != 0) return set_error("gemm: N=%d must be a multiple of 32", a.N);
  if (a.K % 8 != 0 || a.lda % 8 != 0 || a.ldw % 8 != 0)
    return set_error("gemm: K/lda/ldw must be multiples of 8 (16-byte TMA strides)");
  if (a.epi == EPI_LN256 || a.epi == EPI_LN64_GELU) {
    if (a.N != 256 || a.out_fp32 || !a.ln_gamma || !a.ln_beta || a.act)
      return set_error("gemm: fused LN needs N=256, bf16 out, gamma/beta, no act");
    if (a.residual && !a.res_bf16) return set_error("gemm: fused LN takes a bf16 residual");
    return a.epi == EPI_LN256 ? launch_gemm_bn<256, EPI_LN256>(a, num_sms, stream)
                              : launch_gemm_bn<256, EPI_LN64_GELU>(a, num_sms, stream);
  }
  if (a.epi == EPI_HYPER) {
    if (a.N != 128 || !a.hyper || a.hyper_nm < 1 || a.hyper_nm > 4 || a.residual)
      return set_error("gemm: fused hyper product needs N=128");
    return launch_gemm_bn<128, EPI_HYPER>(a, num_sms, stream);
  }
  // BN=256 keeps the tensor pipe at its 1-CTA rate with the fewest smem bytes per flop; fall back to 128 / 64 when N is
  // not a multiple (or is small), to avoid wasted columns.
  if (a.N % 256 == 0) {
    // long-K GEMMs with enough 256-row tiles to fill the machine: 256x256 tiles (weight tile shared by two MMAs)
    static const bool mt2_off = getenv("MSAM_GEMM_MT1") != nullptr;
    const long tiles256 = (long)((a.M + 255) / 256) * (a.N / 256);
    if (!mt2_off && a.K >= 512 && tiles256 >= 2L * num_sms) return launch_gemm_bn<256, EPI_PLAIN, 2>(a, num_sms, stream);
    return launch_gemm_bn<256, EPI_PLAIN>(a, num_sms, stream);
  }
  if (a.N % 128 == 0) return launch_gemm_bn<128, EPI_PLAIN>(a, num_sms, stream);
  return launch_gemm_bn<64, EPI_PLAIN>(a, num_sms, stream);
}

}  // namespace msam
