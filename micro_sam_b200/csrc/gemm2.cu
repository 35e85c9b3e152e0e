// 2-SM (CTA pair) variant of the persistent bf16 GEMM for the large encoder products:  C[M,N] = epi(A[M,K] W[N,K]^T)
//   * cluster of 2 CTAs (one TPC), tcgen05.mma.cta_group::2: one 256 x 256 x 16 MMA per K step for the pair -- CTA r holds
//     rows [128 r, 128 r + 128) of the A tile and rows [128 r, ...) of the W tile (its half of N), so each SM stages 32 KB
//     per 64-wide K block instead of 48 KB and reads 8 KB instead of 12 KB of operands per MMA.  With one CTA per tile the
//     shared-memory port (TMA fill + operand reads ~ 190 B/clk against ~128 B/clk) capped the 128 x 256 kernel at ~80 % of the
//     tensor peak (profiles/r1_gemm_shapes_after_elect.log).
//   * protocol (as in CUTLASS' sm100 2-SM pipelines): both CTAs' TMA loads complete on the LEADER's full barrier (peer bit
//     cleared in the barrier address), the leader's elected lane issues the MMAs and multicasts its commits to the empty /
//     accumulator-full barriers of both CTAs, both CTAs' epilogue warps arrive on the leader's accumulator-empty barrier.
//   * epilogue per CTA (its 128 rows x 256 columns, TMEM lanes = rows): bias / GELU / ReLU, fp32 residual fetched by TMA
//     into the staging tile, swizzled staging -> TMA store.  Same numerics as gemm_bf16_kernel<256, EPI_PLAIN>.
#include "kernels.h"
#include "ptx.cuh"
#include "tensormap.h"

namespace msam {

namespace g2 {
constexpr int BM = 128, BN = 256, BK = 64, STAGES = 5, NG = 4;
constexpr int A_BYTES = BM * BK * 2, B_BYTES = (BN / 2) * BK * 2, STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int STG_BYTES = 128 * 128;
constexpr int OFF_STG = STAGES * STAGE_BYTES;
constexpr int OFF_BAR = OFF_STG + NG * STG_BYTES;
constexpr int SMEM_BYTES = OFF_BAR + 256 + 1024;
constexpr int THREADS = 128 + NG * 128;
constexpr uint32_t PEER_MASK = 0xFEFFFFFFu;  // clears the CTA-rank bit of a shared::cluster address -> the even (leader) CTA
}  // namespace g2

struct Gemm2Params {
  int M, N, K;
  const float* bias;
  int res_rows, has_res;
  int out_fp32, act;
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA load issued by either CTA of the pair; the transaction bytes are credited to the leader CTA's barrier
__device__ __forceinline__ void tma_load_2d_2sm(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & g2::PEER_MASK), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma2_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive (once the MMAs issued so far have completed) on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma2_commit_mc(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {  // arrive on the leader CTA's copy of `bar`
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & g2::PEER_MASK) : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ float g2_gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float g2_gelu_fast(float x) {  // see gemm.cu
  const float z = fabsf(x);
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.47047f * 0.70710678118654752440f, z, 1.0f)));
  float poly = fmaf(0.5f * 0.7478556f, t, 0.5f * -0.0958798f);
  poly = fmaf(poly, t, 0.5f * 0.3480242f);
  poly *= t;
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * (x * -0.72134752044448170368f)));
  return fmaf(-z * poly, e, fmaxf(x, 0.0f));
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(g2::THREADS, 1)
gemm2_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmR, const Gemm2Params p) {
  using namespace g2;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* res_bar = tempty_bar + 2;  // [NG]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_bar + NG);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
  const int m2_blocks = (p.M + 2 * BM - 1) / (2 * BM), n_blocks = p.N / BN, k_blocks = (p.K + BK - 1) / BK;
  const int num_tiles = m2_blocks * n_blocks;

  if (warp == 0 && lane == 0) { prefetch_tmap(&tmA); prefetch_tmap(&tmB); prefetch_tmap(&tmC); prefetch_tmap(&tmR); }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], 2 * NG * 4); }
    for (int i = 0; i < NG; ++i) mbar_init(&res_bar[i], 1);
    fence_barrier_init();
  }
  cluster_sync_all();                    // barriers of both CTAs are initialised before any remote arrive / TMA credit
  if (warp == 2) tmem_alloc2(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();      // the previous kernel's outputs (A, residual) are complete and visible from here on
  pdl_trigger();

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (one per CTA: its A rows, its half of W)
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += n_clusters) {
        const int m_blk = 2 * (tile / n_blocks) + (int)rank, n_blk = tile % n_blocks;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1, 1);
          uint8_t* sa = smem + stage * STAGE_BYTES;
          if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * STAGE_BYTES);  // bytes of both CTAs land on the leader's barrier
          tma_load_2d_2sm(sa, &tmA, &full_bar[stage], kb * BK, m_blk * BM);
          tma_load_2d_2sm(sa + A_BYTES, &tmB, &full_bar[stage], kb * BK, n_blk * BN + (int)rank * (BN / 2));
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer: leader CTA only
    if (rank == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(2 * BM, BN);
      int stage = 0, it = 0;
      uint32_t phase = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += n_clusters, ++it) {
        const int as = it & 1;
        mbar_wait(&tempty_bar[as], ((it >> 1) & 1) ^ 1, 2);  // both CTAs' epilogues have drained this accumulator stage
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * BN;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase, 3);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
          const uint64_t da = make_desc_sw128(sa, 0, 1024), db = make_desc_sw128(sa + A_BYTES, 0, 1024);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) umma2_bf16(tmem_d, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb | k) != 0);
            umma2_commit_mc(&empty_bar[stage]);
            if (kb == k_blocks - 1) umma2_commit_mc(&tfull_bar[as]);
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------ epilogue (both CTAs, own 128 rows)
    const int quad = warp & 3, grp = (warp - 4) >> 2, r = quad * 32 + lane;
    uint8_t* stg = smem + OFF_STG + grp * STG_BYTES;
    const uint32_t stg_a = smem_u32(stg) + r * 128;
    const bool issuer = (quad == 0 && lane == 0);
    const int bar_id = 2 + grp;
    int it = 0;
    uint32_t res_cnt = 0;
    for (int tile = cluster_id; tile < num_tiles; tile += n_clusters, ++it) {
      const int m_blk = 2 * (tile / n_blocks) + (int)rank, n_blk = tile % n_blocks;
      const int as = it & 1;
      const int colbase = n_blk * BN + grp * 64;
      const uint32_t tcol = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(as * BN + grp * 64);
      auto release_acc = [&]() {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_leader(&tempty_bar[as]);
      };
      auto publish = [&](int col) {
        fence_proxy_async_smem();
        named_bar_sync(bar_id, 128);
        if (issuer) { tma_store_2d(&tmC, stg, col, m_blk * BM); tma_store_commit(); }
      };
      if (p.has_res && issuer) {  // residual of the first 32-column chunk: prefetched while the tile is accumulated
        tma_store_wait_read();
        mbar_expect_tx(&res_bar[grp], STG_BYTES);
        tma_load_2d(stg, &tmR, &res_bar[grp], colbase, (m_blk * BM) % p.res_rows);
      }
      mbar_wait(&tfull_bar[as], (it >> 1) & 1, 4);
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        const int col0 = colbase + c * 32;
        uint32_t v[32];
        tmem_ld32(tcol + c * 32, v);
        tmem_ld_wait();
        if (c == 1) release_acc();
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
        if (p.act == 1) {
          if (p.out_fp32) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = g2_gelu_erf(f[j]);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = g2_gelu_fast(f[j]);
          }
        } else if (p.act == 2) {
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
        }
        if (p.out_fp32) {  // one [128 x 32] fp32 store per chunk
          if (p.has_res) {
            mbar_wait(&res_bar[grp], res_cnt & 1, 5);
            ++res_cnt;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const uint32_t a = stg_a + ((q ^ (r & 7)) << 4);
              const uint4 rr = ld_shared_v4(a);
              st_shared_v4(a, make_uint4(__float_as_uint(f[4 * q] + __uint_as_float(rr.x)), __float_as_uint(f[4 * q + 1] + __uint_as_float(rr.y)),
                                         __float_as_uint(f[4 * q + 2] + __uint_as_float(rr.z)), __float_as_uint(f[4 * q + 3] + __uint_as_float(rr.w))));
            }
          } else {
            if (issuer) tma_store_wait_read();
            named_bar_sync(bar_id, 128);
#pragma unroll
            for (int q = 0; q < 8; ++q)
              st_shared_v4(stg_a + ((q ^ (r & 7)) << 4), make_uint4(__float_as_uint(f[4 * q]), __float_as_uint(f[4 * q + 1]),
                                                                   __float_as_uint(f[4 * q + 2]), __float_as_uint(f[4 * q + 3])));
          }
          publish(col0);
          if (p.has_res && c == 0 && issuer) {
            tma_store_wait_read();
            mbar_expect_tx(&res_bar[grp], STG_BYTES);
            tma_load_2d(stg, &tmR, &res_bar[grp], col0 + 32, (m_blk * BM) % p.res_rows);
          }
        } else {           // 64 bf16 columns per staging row: one store per tile
          if (c == 0) {
            if (issuer) tma_store_wait_read();
            named_bar_sync(bar_id, 128);
          }
#pragma unroll
          for (int q = 0; q < 4; ++q)
            st_shared_v4(stg_a + (((c * 4 + q) ^ (r & 7)) << 4),
                         make_uint4(pack_bf16(f[8 * q], f[8 * q + 1]), pack_bf16(f[8 * q + 2], f[8 * q + 3]),
                                    pack_bf16(f[8 * q + 4], f[8 * q + 5]), pack_bf16(f[8 * q + 6], f[8 * q + 7])));
          if (c == 1) publish(colbase);
        }
      }
      __syncwarp();
    }
    if (issuer) tma_store_wait_all();
  }

  tc_fence_before();
  cluster_sync_all();  // the partner may still be reading this CTA's shared memory / signalling its barriers
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, 512);
  }
}

// Returns 1 if the problem was launched on the 2-SM kernel, 0 if it does not qualify (caller falls back), -1 on error.
int launch_gemm_2sm(const GemmArgs& a, int num_sms, cudaStream_t stream) {
  using namespace g2;
  const int res_rows = a.res_rows > 0 ? a.res_rows : a.M, ldr = a.ldr > 0 ? a.ldr : a.N;
  if (a.epi != 0 || a.act_after_res || a.N % BN != 0 || a.K % BK != 0 || a.K < 512 || a.M < 4096) return 0;
  if (a.residual && (a.res_bf16 || !a.out_fp32 || res_rows % BM != 0 || ldr % 4 != 0)) return 0;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm2_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) return set_error("gemm2: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  const int ldc = a.ldc > 0 ? a.ldc : a.N;
  CUtensorMap tmA, tmB, tmC, tmR;
  if (make_tmap_bf16_2d(&tmA, a.A, a.M, a.K, a.lda, BM)) return -1;
  if (make_tmap_bf16_2d(&tmB, a.W, a.N, a.K, a.ldw, BN / 2)) return -1;
  if (make_tmap_2d(&tmC, a.out, a.out_fp32 ? 4 : 2, a.M, a.N, ldc, BM)) return -1;
  if (a.residual) {
    if (make_tmap_2d(&tmR, a.residual, 4, res_rows, a.N, ldr, BM)) return -1;
  } else {
    tmR = tmA;
  }
  Gemm2Params p;
  p.M = a.M; p.N = a.N; p.K = a.K; p.bias = a.bias; p.res_rows = res_rows; p.has_res = a.residual ? 1 : 0;
  p.out_fp32 = a.out_fp32; p.act = a.act;
  const int tiles = ((a.M + 2 * BM - 1) / (2 * BM)) * (a.N / BN);
  const int max_clusters = num_sms / 2;
  const int clusters = tiles < max_clusters ? tiles : max_clusters;
  prof_begin(stream, "gemm2_bf16 (2-SM, encoder)", 2.0 * a.M * a.N * a.K,
             (double)a.M * a.K * 2 + (double)a.N * a.K * 2 + (double)a.M * a.N * (a.out_fp32 ? 4 : 2) * (a.residual ? 2 : 1));
  launch_pdl(gemm2_bf16_kernel, dim3(2 * clusters), dim3(THREADS), SMEM_BYTES, stream, tmA, tmB, tmC, tmR, p);
  prof_end(stream);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error("gemm2 launch failed: %s", cudaGetErrorString(e));
  count_launch();
  return 1;
}

}  // namespace msam
