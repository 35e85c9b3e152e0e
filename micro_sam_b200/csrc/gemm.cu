// Persistent warp-specialised bf16 GEMM for sm_100a:  C[M,N] = epilogue(A[M,K] * W[N,K]^T)
//   * operands: TMA (SWIZZLE_128B boxes of 64 K-elements) -> shared memory ring (mbarrier full/empty)
//   * math:     tcgen05.mma cta_group::1 kind::f16, M=128 x N=BN x K=16, fp32 accumulators in TMEM (2 stages)
//   * epilogue: 8 warps, tcgen05.ld 32x32b -> bias / GELU(erf) / ReLU / fp32 residual (row modulus) -> bf16 or fp32
// One CTA per SM; tiles are distributed round-robin, N-block fastest so that the CTAs that run concurrently share the
// same A row-block through L2.
//
// Replaces (on the B200 path) the nn.Linear / 1x1-conv / patch-embed conv calls inside segment_anything's
// ImageEncoderViT / MaskDecoder that micro_sam reaches through util.py:674 (image_encoder) and
// SamPredictor.predict_torch (inference.py:248, instance_segmentation.py:361).
#include "kernels.h"
#include "ptx.cuh"
#include "tensormap.h"

namespace msam {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;
constexpr int GEMM_THREADS = 384;  // warp0 TMA, warp1 MMA, warp2 TMEM alloc, warp3 idle, warps 4..11 epilogue

template <int BN>
struct GemmCfg {
  static constexpr int STAGES = (BN == 256) ? 4 : 6;
  static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;
  static constexpr int B_BYTES = BN * GEMM_BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/ + 4096 /*LN exchange*/;
  static constexpr int TMEM_COLS = 2 * BN;  // power of two >= 32 for BN in {64,128,256}
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

__device__ __forceinline__ float apply_act(float x, int act) {
  if (act == 1) return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
  if (act == 2) return fmaxf(x, 0.0f);
  return x;
}

enum { EPI_PLAIN = 0, EPI_LN256 = 1, EPI_LN64_GELU = 2, EPI_HYPER = 3 };

template <int BN, int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
  using Cfg = GemmCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + Cfg::STAGES;
  uint64_t* tfull_bar = empty_bar + Cfg::STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  float* exch = reinterpret_cast<float*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES + 256);  // [2][256] float2 (EPI_LN256)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m_blocks = (p.M + GEMM_BM - 1) / GEMM_BM;
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
      mbar_init(&tempty_bar[i], 8);  // one arrive per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
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
          tma_load_2d(sa, &tmA, &full_bar[stage], kb * GEMM_BK, m_blk * GEMM_BM);
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
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
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
          }
          umma_commit(&empty_bar[stage]);  // smem slot free once these MMAs have read it
          if (kb == k_blocks - 1) umma_commit(&tfull_bar[as]);
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------ epilogue: TMEM -> regs -> (fused op) -> global
    const int ew = warp - 4;         // 0..7
    const int quad = warp & 3;       // TMEM lane quadrant this warp may access
    const int half = ew >> 2;        // column half of the tile
    constexpr int CPW = BN / 2;      // columns per warp
    constexpr int NCH = CPW / 32;    // 32-column chunks per warp
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int m_blk = tile / n_blocks, n_blk = tile % n_blocks;
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      mbar_wait(&tfull_bar[as], aphase, 4);
      tc_fence_after();
      const int row = m_blk * GEMM_BM + quad * 32 + lane;
      const bool row_ok = row < p.M;
      const float* res_row = nullptr;
      const __nv_bfloat16* res_row_bf = nullptr;
      if (p.residual && row_ok) {
        const size_t off = (size_t)(row % p.res_rows) * p.ldr;
        if (p.res_bf16) res_row_bf = reinterpret_cast<const __nv_bfloat16*>(p.residual) + off;
        else res_row = reinterpret_cast<const float*>(p.residual) + off;
      }
      // accumulator chunk c (+ bias, activation, residual) -> f[0..32)
      auto load_chunk = [&](int c, float* f, int col0) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(as * BN + half * CPW + c * 32), v);
        tmem_ld_wait();
        if (c == NCH - 1) {
          // all TMEM reads of this accumulator stage by this warp are complete -> hand it back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty_bar[as]);
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
        if (!(row_ok && col0 < p.N)) return;
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
        if (res_row) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 r = *reinterpret_cast<const float4*>(res_row + col0 + j);
            f[j] += r.x; f[j + 1] += r.y; f[j + 2] += r.z; f[j + 3] += r.w;
          }
        } else if (res_row_bf) {
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            const uint4 r = *reinterpret_cast<const uint4*>(res_row_bf + col0 + j);
            const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              f[j + 2 * q] += __uint_as_float(w[q] << 16);
              f[j + 2 * q + 1] += __uint_as_float(w[q] & 0xffff0000u);
            }
          }
        }
      };
      auto store_bf16 = [&](const float* f, int col0) {
        __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + (size_t)row * p.ldc + col0;
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          uint4 u;
          u.x = pack_bf16(f[j], f[j + 1]);
          u.y = pack_bf16(f[j + 2], f[j + 3]);
          u.z = pack_bf16(f[j + 4], f[j + 5]);
          u.w = pack_bf16(f[j + 6], f[j + 7]);
          *reinterpret_cast<uint4*>(o + j) = u;
        }
      };

      if constexpr (EPI == EPI_PLAIN) {
#pragma unroll 1
        for (int c = 0; c < NCH; ++c) {
          const int col0 = n_blk * BN + half * CPW + c * 32;
          float f[32];
          load_chunk(c, f, col0);
          if (row_ok && col0 < p.N) {
            if (p.out_fp32) {
              float* o = reinterpret_cast<float*>(p.out) + (size_t)row * p.ldc + col0;
#pragma unroll
              for (int j = 0; j < 32; j += 4)
                *reinterpret_cast<float4*>(o + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
            } else {
              store_bf16(f, col0);
            }
          }
          __syncwarp();  // reconverge before the next warp-collective tcgen05.ld
        }
      } else {
        // fused row epilogues: the whole half-row (CPW columns) of this thread lives in registers
        float f[CPW];
        const int colbase = n_blk * BN + half * CPW;
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          load_chunk(c, f + c * 32, colbase + c * 32);
          __syncwarp();
        }
        if constexpr (EPI == EPI_LN256) {
          // LayerNorm over the full 256-wide row (BN == N == 256): exact two-pass statistics per half, combined
          // across the two column halves with Chan's formula through shared memory.
          float mean_h = 0.f;
#pragma unroll
          for (int j = 0; j < CPW; ++j) mean_h += f[j];
          mean_h *= (1.0f / CPW);
          float m2_h = 0.f;
#pragma unroll
          for (int j = 0; j < CPW; ++j) { const float d = f[j] - mean_h; m2_h += d * d; }
          float2* ex = reinterpret_cast<float2*>(exch) + (it & 1) * 256;
          ex[half * 128 + quad * 32 + lane] = make_float2(mean_h, m2_h);
          asm volatile("bar.sync 1, 256;" ::: "memory");
          const float2 o2 = ex[(half ^ 1) * 128 + quad * 32 + lane];
          const float mean = 0.5f * (mean_h + o2.x);
          const float dm = mean_h - o2.x;
          const float var = (m2_h + o2.y + dm * dm * (0.5f * CPW)) * (1.0f / (2 * CPW));
          const float rstd = rsqrtf(var + p.ln_eps);
          if (row_ok) {
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const float4 g = __ldg(reinterpret_cast<const float4*>(p.ln_gamma + colbase + c * 32 + j));
                const float4 b = __ldg(reinterpret_cast<const float4*>(p.ln_beta + colbase + c * 32 + j));
                float* x = f + c * 32 + j;
                x[0] = (x[0] - mean) * rstd * g.x + b.x; x[1] = (x[1] - mean) * rstd * g.y + b.y;
                x[2] = (x[2] - mean) * rstd * g.z + b.z; x[3] = (x[3] - mean) * rstd * g.w + b.w;
              }
              store_bf16(f + c * 32, colbase + c * 32);
            }
          }
        } else if constexpr (EPI == EPI_LN64_GELU) {
          // LayerNorm2d over 64-channel groups (conv-transpose sub-pixels) + exact GELU; groups never straddle threads
          if (row_ok) {
#pragma unroll
            for (int g0 = 0; g0 < CPW; g0 += 64) {
              float mean = 0.f;
#pragma unroll
              for (int j = 0; j < 64; ++j) mean += f[g0 + j];
              mean *= (1.0f / 64);
              float m2 = 0.f;
#pragma unroll
              for (int j = 0; j < 64; ++j) { const float d = f[g0 + j] - mean; m2 += d * d; }
              const float rstd = rsqrtf(m2 * (1.0f / 64) + p.ln_eps);
#pragma unroll
              for (int j = 0; j < 64; ++j) {
                const float y = (f[g0 + j] - mean) * rstd * __ldg(p.ln_gamma + j) + __ldg(p.ln_beta + j);
                f[g0 + j] = 0.5f * y * (1.0f + erff(y * 0.70710678118654752440f));
              }
            }
#pragma unroll
            for (int c = 0; c < NCH; ++c) store_bf16(f + c * 32, colbase + c * 32);
          }
        } else if constexpr (EPI == EPI_HYPER) {
          // second conv-transpose (BN == N == 128: 4 sub-sub-pixels x 32 channels, bias + GELU already applied) fused
          // with the hyper-network product: masks[p, mi, Y, X] = sum_ch hyper[p, m0+mi, ch] * up[.., ss*32 + ch].
          // GEMM row = (prompt p, token (y,x), sub-pixel (dy,dx)); this thread holds sub-sub-pixels ey = half, ex = 0,1.
          if (row_ok) {
            const int sub = row & 3, tok = (row >> 2) & 4095, pp = row >> 14;
            const int y = tok >> 6, x = tok & 63, dy = sub >> 1, dx = sub & 1;
            const int Y = 4 * y + 2 * dy + half, X = 4 * x + 2 * dx;
            for (int mi = 0; mi < p.hyper_nm; ++mi) {
              const float* hw = p.hyper + ((size_t)pp * 4 + p.hyper_m0 + mi) * 32;
              float a0 = 0.f, a1 = 0.f;
#pragma unroll
              for (int c = 0; c < 32; c += 4) {
                const float4 h4 = __ldg(reinterpret_cast<const float4*>(hw + c));
                a0 += h4.x * f[c] + h4.y * f[c + 1] + h4.z * f[c + 2] + h4.w * f[c + 3];
                a1 += h4.x * f[32 + c] + h4.y * f[33 + c] + h4.z * f[34 + c] + h4.w * f[35 + c];
              }
              float* o = reinterpret_cast<float*>(p.out) + (((size_t)pp * p.hyper_nm + mi) * 256 + Y) * 256 + X;
              *reinterpret_cast<float2*>(o) = make_float2(a0, a1);
            }
          }
        }
        __syncwarp();
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

template <int BN, int EPI>
static int launch_gemm_bn(const GemmArgs& a, int num_sms, cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e =
        cudaFuncSetAttribute(gemm_bf16_kernel<BN, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return set_error("gemm: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  CUtensorMap tmA, tmB;
  if (make_tmap_bf16_2d(&tmA, a.A, a.M, a.K, a.lda, GEMM_BM)) return -1;
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
  const int tiles = ((a.M + GEMM_BM - 1) / GEMM_BM) * ((a.N + BN - 1) / BN);
  const int grid = tiles < num_sms ? tiles : num_sms;
  prof_begin(stream, PROF_GEMM, 2.0 * a.M * a.N * a.K);
  gemm_bf16_kernel<BN, EPI><<<grid, GEMM_THREADS, Cfg::SMEM_BYTES, stream>>>(tmA, tmB, p);
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
    if (a.N != 256 || a.out_fp32 || !a.ln_gamma || !a.ln_beta) return set_error("gemm: fused LN needs N=256, bf16 out, gamma/beta");
    return a.epi == EPI_LN256 ? launch_gemm_bn<256, EPI_LN256>(a, num_sms, stream)
                              : launch_gemm_bn<256, EPI_LN64_GELU>(a, num_sms, stream);
  }
  if (a.epi == EPI_HYPER) {
    if (a.N != 128 || !a.hyper || a.hyper_nm < 1 || a.hyper_nm > 4) return set_error("gemm: fused hyper product needs N=128");
    return launch_gemm_bn<128, EPI_HYPER>(a, num_sms, stream);
  }
  // BN=256 keeps the tensor pipe at its 1-CTA rate with the fewest smem bytes per flop; fall back to 128 / 64 when N is
  // not a multiple (or is small), to avoid wasted columns.
  if (a.N % 256 == 0) return launch_gemm_bn<256, EPI_PLAIN>(a, num_sms, stream);
  if (a.N % 128 == 0) return launch_gemm_bn<128, EPI_PLAIN>(a, num_sms, stream);
  return launch_gemm_bn<64, EPI_PLAIN>(a, num_sms, stream);
}

}  // namespace msam
