// Fused output up-scaling + hyper-network mask product of the mask decoder (MaskDecoder.predict_masks after the
// transformer, restated in oracle/sam_ref.py:  upscaled = GELU(convT2(GELU(LN2d(convT1(src)))));  masks = hyper_in @ upscaled):
//
//   keys tile [128 tokens x 256]  --MMA1-->  D1 [128 x (4 sub-pixels x 64 ch)]        conv-transpose 1 (k2 s2) as a GEMM
//        E1: + bias, LayerNorm2d over the 64 channels of a (token, sub-pixel), GELU  -> A2_s [128 x 64] fp16 in shared memory
//   A2_s  --MMA2-->  D2 [128 x (4 sub-sub-pixels x 32 ch)]  (s = 0..3)               conv-transpose 2 (k2 s2) as a GEMM
//        E2: + bias, GELU, dot product with the prompt's hyper-network vectors (<= 4 masks x 32 ch) -> low-res logits
//
// so neither the 64-channel up-scaled embedding (2 MB per prompt written + read by the two-kernel version) nor the
// 32-channel one ever leaves the SM: per prompt the kernel reads 2 MB of `keys` and writes 0.25 MB per mask.
// The element-wise epilogues -- 3.2 G GELUs per 32x32-grid tile, what bounded the previous kernels at 62 % issue-slot
// utilisation with the tensor pipe 5 % active (profiles/r1_ncu_hyper_final.txt) -- run in PACKED fp16x2 arithmetic:
//   GELU(x) = 0.5 x (1 + erf(x / sqrt 2)),  erf(x / sqrt 2) ~ tanh(x (a + b x^2))   (minimax a, b: |err| <= 2.7e-4; with the
//   fp16 rounding of the 6-instruction chain the N(0,1)-weighted rms error is 2.9e-4 -- a quarter of the bf16 rounding the
//   two-kernel version applied when it stored the intermediate), one MUFU (tanh.approx.f16x2) per TWO elements,
// the intermediate operand A2 and the conv-transpose-2 weights are fp16 (11-bit mantissa instead of bf16's 8), the hyper
// product accumulates 2 x 8 fp16x2 FMAs per mask and finishes in fp32.
//
// CTA = TMA warp + MMA warp + 16 epilogue warps (4 TMEM lane quadrants x 4 column groups), persistent over a contiguous
// range of (prompt, 128-token tile) items; TMEM: D1 = columns [0,256), D2 = 2 x 128 columns (double buffered over s).
#include <cuda_fp16.h>

#include "kernels.h"
#include "ptx.cuh"
#include "tensormap.h"

namespace msam {

namespace up {
constexpr int STAGES = 2;
constexpr int SUBA = 128 * 128;                 // [128 rows x 64 x 16-bit] SWIZZLE_128B sub-tile (16 KB)
constexpr int SUBW = 256 * 128;                 // conv-transpose-1 weight K-slice [256 x 64] bf16 (32 KB)
constexpr int STAGE_BYTES = SUBA + SUBW;
constexpr int OFF_A2 = STAGES * STAGE_BYTES;    // 4 x [128 x 64] fp16
constexpr int OFF_W2 = OFF_A2 + 4 * SUBA;       // [128 x 64] fp16
constexpr int OFF_BAR = OFF_W2 + SUBA;
constexpr int SMEM_BYTES = OFF_BAR + 256 + 1024;
constexpr int THREADS = 128 + 512;
constexpr uint32_t TM_D1 = 0, TM_D2 = 256, TMEM_COLS = 512;
constexpr int TILES = 32;                       // 4096 image tokens / 128
}  // namespace up

struct UpParams {
  int P, nm, m0;
  const float* b1;      // [256] conv-transpose-1 bias per (sub-pixel, channel)
  const float* gamma;   // [64] LayerNorm2d
  const float* beta;
  float eps;
  const float* b2;      // [128] conv-transpose-2 bias per (sub-sub-pixel, channel)
  const float* hyper;   // [P, 4, 32]
  float* out;           // [P, nm, 256, 256]
};

__host__ __device__ constexpr uint32_t make_idesc_f16(uint32_t M, uint32_t N) {  // kind::f16, A = B = F16 (format 0), D = F32
  return (1u << 4) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

__device__ __forceinline__ __half2 tanh_h2(__half2 x) {
  uint32_t r, a = *reinterpret_cast<uint32_t*>(&x);
  asm("tanh.approx.f16x2 %0, %1;" : "=r"(r) : "r"(a));
  return *reinterpret_cast<__half2*>(&r);
}
__device__ __forceinline__ __half2 gelu_h2(__half2 v) {
  const __half2 a = __floats2half2_rn(0.80015708f, 0.80015708f), b = __floats2half2_rn(0.03470089f, 0.03470089f);
  const __half2 hlf = __floats2half2_rn(0.5f, 0.5f);
  const __half2 t = __hfma2(b, __hmul2(v, v), a);
  const __half2 th = tanh_h2(__hmul2(v, t));
  const __half2 h = __hmul2(v, hlf);
  return __hfma2(h, th, h);
}
__device__ __forceinline__ uint32_t h2u(__half2 v) { return *reinterpret_cast<uint32_t*>(&v); }

__global__ void __launch_bounds__(up::THREADS, 1)
upscale_fused_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW1,
                     const __grid_constant__ CUtensorMap tmW2, const UpParams p) {
  using namespace up;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* w2_full = empty_bar + STAGES;
  uint64_t* d1_full = w2_full + 1;
  uint64_t* d1_empty = d1_full + 1;
  uint64_t* a2_full = d1_empty + 1;    // [4]
  uint64_t* a2_empty = a2_full + 4;    // [4]
  uint64_t* d2_full = a2_empty + 4;    // [2]
  uint64_t* d2_empty = d2_full + 2;    // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(d2_empty + 2);
  __shared__ __align__(16) float b1_s[256];
  __shared__ __align__(16) __half2 gb_s[64];          // [0,32) gamma pairs, [32,64) beta pairs
  __shared__ __align__(16) __half2 b2_s[64];          // conv-transpose-2 bias pairs, index ss*16 + i
  __shared__ __align__(16) __half2 hyp_s[16 * 2 * 64];  // per epilogue warp and item parity: [mask < 4][16 channel pairs]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long total = (long)p.P * TILES;
  const int it_begin = (int)(total * blockIdx.x / gridDim.x), it_end = (int)(total * (blockIdx.x + 1) / gridDim.x);
  const int n_items = it_end - it_begin;

  if (warp == 0 && lane == 0) { prefetch_tmap(&tmX); prefetch_tmap(&tmW1); prefetch_tmap(&tmW2); }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    mbar_init(w2_full, 1); mbar_init(d1_full, 1); mbar_init(d1_empty, 16);
    for (int i = 0; i < 4; ++i) { mbar_init(&a2_full[i], 4); mbar_init(&a2_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&d2_full[i], 1); mbar_init(&d2_empty[i], 16); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, TMEM_COLS);
  for (int i = threadIdx.x; i < 256; i += THREADS) b1_s[i] = p.b1[i];
  for (int i = threadIdx.x; i < 32; i += THREADS) {
    gb_s[i] = __floats2half2_rn(p.gamma[2 * i], p.gamma[2 * i + 1]);
    gb_s[32 + i] = __floats2half2_rn(p.beta[2 * i], p.beta[2 * i + 1]);
  }
  for (int i = threadIdx.x; i < 64; i += THREADS) b2_s[i] = __floats2half2_rn(p.b2[2 * i], p.b2[2 * i + 1]);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0 && n_items > 0) {
      mbar_expect_tx(w2_full, SUBA);
      tma_load_2d(smem + OFF_W2, &tmW2, w2_full, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      for (int item = it_begin; item < it_end; ++item) {
        const int row0 = (item / TILES) * 4096 + (item % TILES) * 128;
        if (item + 2 < it_end) {  // keys tile of a later item -> L2
          const int pr = ((item + 2) / TILES) * 4096 + ((item + 2) % TILES) * 128;
#pragma unroll
          for (int j = 0; j < 4; ++j) tma_prefetch_2d(&tmX, 64 * j, pr);
        }
        for (int j = 0; j < 4; ++j) {
          mbar_wait(&empty_bar[stage], phase ^ 1, 40);
          uint8_t* sa = smem + stage * STAGE_BYTES;
          mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
          tma_load_2d(sa, &tmX, &full_bar[stage], 64 * j, row0);
          tma_load_2d(sa + SUBA, &tmW1, &full_bar[stage], 64 * j, 0);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (warp-uniform control flow, elected lane issues)
    constexpr uint32_t idesc1 = make_idesc_bf16(128, 256);
    constexpr uint32_t idesc2 = make_idesc_f16(128, 128);
    const uint64_t dw2 = make_desc_sw128(smem_u32(smem + OFF_W2), 0, 1024);
    int stage = 0;
    uint32_t phase = 0;
    auto mma1 = [&](int it) {
      if (it > 0) mbar_wait(d1_empty, (it - 1) & 1, 41);  // every epilogue warp has pulled the previous D1 out of TMEM
      for (int j = 0; j < 4; ++j) {
        mbar_wait(&full_bar[stage], phase, 42);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
        const uint64_t da = make_desc_sw128(sa, 0, 1024), db = make_desc_sw128(sa + SUBA, 0, 1024);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(tmem_base + TM_D1, da + 2 * k, db + 2 * k, idesc1, (j | k) != 0);
          umma_commit(&empty_bar[stage]);
          if (j == 3) umma_commit(d1_full);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    };
    auto mma2 = [&](int it, int s) {
      const int b = s & 1;
      const uint32_t n = 2u * it + (s >> 1);  // use counter of D2 buffer b
      mbar_wait(&a2_full[s], it & 1, 43);
      if (n > 0) mbar_wait(&d2_empty[b], (n - 1) & 1, 44);
      tc_fence_after();
      const uint64_t da = make_desc_sw128(smem_u32(smem + OFF_A2 + s * SUBA), 0, 1024);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tmem_base + TM_D2 + b * 128, da + 2 * k, dw2 + 2 * k, idesc2, k != 0);
        umma_commit(&d2_full[b]);
        umma_commit(&a2_empty[s]);
      }
      __syncwarp();
    };
    if (n_items > 0) {
      mbar_wait(w2_full, 0, 45);
      mma1(0);
      for (int it = 0; it < n_items; ++it) {
        mma2(it, 0);
        mma2(it, 1);
        if (it + 1 < n_items) mma1(it + 1);  // overlaps E2 of this item
        mma2(it, 2);
        mma2(it, 3);
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------ epilogue warps
    const int ew = warp - 4, quad = warp & 3, grp = ew >> 2, r = quad * 32 + lane;
    const uint32_t tlane = tmem_base + ((uint32_t)(quad * 32) << 16);
    const uint32_t a2row = smem_u32(smem + OFF_A2 + grp * SUBA) + r * 128;
    const float inv64 = 1.0f / 64.0f;
    for (int it = 0; it < n_items; ++it) {
      const int item = it_begin + it, pp = item / TILES, rt = item % TILES;
      // hyper-network vectors of this prompt -> private fp16x2 copy (read back as shared-memory broadcasts in E2)
      __half2* hw = hyp_s + (ew * 2 + (it & 1)) * 64;
      for (int i = lane; i < p.nm * 16; i += 32) {
        const float2 v = __ldg(reinterpret_cast<const float2*>(p.hyper + ((size_t)pp * 4 + p.m0) * 32) + i);
        hw[i] = __floats2half2_rn(v.x, v.y);
      }
      __syncwarp();

      // ---- E1: (row r, sub-pixel grp): + bias, LayerNorm over 64 channels, GELU -> A2[grp] row r (fp16, K-major SW128)
      mbar_wait(d1_full, it & 1, 46);
      tc_fence_after();
      float s4[4] = {0.f, 0.f, 0.f, 0.f}, q4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t v[32];
        tmem_ld32(tlane + TM_D1 + 64 * grp + 32 * c, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float x = __uint_as_float(v[j]) + b1_s[64 * grp + 32 * c + j];
          s4[j & 3] += x;
          q4[j & 3] = fmaf(x, x, q4[j & 3]);
        }
      }
      const float sum = (s4[0] + s4[1]) + (s4[2] + s4[3]);
      const float mean = sum * inv64;
      const float var = fmaxf(((q4[0] + q4[1]) + (q4[2] + q4[3])) * inv64 - mean * mean, 0.f);
      const float rstd = rsqrtf(var + p.eps);
      const float shift = -mean * rstd;
      if (it > 0) mbar_wait(&a2_empty[grp], (it - 1) & 1, 47);  // MMA2 of the previous item has read A2[grp]
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t v[32];
        tmem_ld32(tlane + TM_D1 + 64 * grp + 32 * c, v);
        tmem_ld_wait();
        uint32_t o[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float x0 = __uint_as_float(v[2 * i]) + b1_s[64 * grp + 32 * c + 2 * i];
          const float x1 = __uint_as_float(v[2 * i + 1]) + b1_s[64 * grp + 32 * c + 2 * i + 1];
          const __half2 n2 = __floats2half2_rn(fmaf(x0, rstd, shift), fmaf(x1, rstd, shift));
          o[i] = h2u(gelu_h2(__hfma2(n2, gb_s[16 * c + i], gb_s[32 + 16 * c + i])));
        }
#pragma unroll
        for (int q = 0; q < 4; ++q)
          st_shared_v4(a2row + (((4 * c + q) ^ (r & 7)) << 4), make_uint4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]));
      }
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) { mbar_arrive(&a2_full[grp]); mbar_arrive(d1_empty); }

      // ---- E2: (row r, sub-pixel s, sub-sub-pixel grp): + bias, GELU, hyper product over the 32 channels -> low-res logits
      const int tok = rt * 128 + r, ty = tok >> 6, tx = tok & 63;
      float* obase = p.out + (size_t)pp * p.nm * 65536 + (size_t)(4 * ty + (grp >> 1)) * 256 + 4 * tx + (grp & 1);
#pragma unroll 1
      for (int s = 0; s < 4; ++s) {
        const int b = s & 1;
        const uint32_t n = 2u * it + (s >> 1);
        mbar_wait(&d2_full[b], n & 1, 48);
        tc_fence_after();
        uint32_t v[32];
        tmem_ld32(tlane + TM_D2 + b * 128 + 32 * grp, v);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&d2_empty[b]);
        __half2 g[16];
#pragma unroll
        for (int i = 0; i < 16; ++i)
          g[i] = gelu_h2(__hadd2(__floats2half2_rn(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1])), b2_s[grp * 16 + i]));
        float* o = obase + (size_t)(2 * (s >> 1)) * 256 + 2 * (s & 1);
        for (int mi = 0; mi < p.nm; ++mi) {
          const __half2* hm = hw + mi * 16;
          __half2 a0 = __floats2half2_rn(0.f, 0.f), a1 = a0;
#pragma unroll
          for (int i = 0; i < 16; i += 2) {
            a0 = __hfma2(g[i], hm[i], a0);
            a1 = __hfma2(g[i + 1], hm[i + 1], a1);
          }
          const float2 f0 = __half22float2(a0), f1 = __half22float2(a1);
          o[(size_t)mi * 65536] = (f0.x + f0.y) + (f1.x + f1.y);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

int launch_upscale_fused(const UpscaleFusedArgs& a, int num_sms, cudaStream_t stream) {
  using namespace up;
  if (a.P <= 0 || a.nm < 1 || a.nm > 4 || a.m0 < 0 || a.m0 + a.nm > 4) return set_error("upscale_fused: bad arguments (P=%d nm=%d m0=%d)", a.P, a.nm, a.m0);
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(upscale_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) return set_error("upscale_fused: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  CUtensorMap tmX, tmW1, tmW2;
  if (make_tmap_bf16_2d(&tmX, a.keys, (uint64_t)a.P * 4096, 256, 256, 128)) return -1;
  if (make_tmap_bf16_2d(&tmW1, a.w1, 256, 256, 256, 256)) return -1;
  if (make_tmap_f16_2d(&tmW2, a.w2_f16, 128, 64, 64, 128)) return -1;
  UpParams p;
  p.P = a.P; p.nm = a.nm; p.m0 = a.m0; p.b1 = a.b1; p.gamma = a.gamma; p.beta = a.beta; p.eps = a.eps; p.b2 = a.b2;
  p.hyper = a.hyper; p.out = a.out;
  const long total = (long)a.P * TILES;
  const int grid = total < num_sms ? (int)total : num_sms;
  prof_begin(stream, "upscale_fused (convT1+LN2d+GELU+convT2+GELU+hyper)",
             (double)a.P * 4096 * (2.0 * 256 * 256 + 4 * 2.0 * 128 * 64 + 2.0 * 512 * a.nm),
             (double)a.P * (4096.0 * 256 * 2 + 65536.0 * 4 * a.nm));
  upscale_fused_kernel<<<grid, THREADS, SMEM_BYTES, stream>>>(tmX, tmW1, tmW2, p);
  prof_end(stream);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error("upscale_fused launch failed: %s", cudaGetErrorString(e));
  count_launch();
  return 0;
}

}  // namespace msam
