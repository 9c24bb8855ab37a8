// tcgen05 / TMEM / TMA bf16 GEMM, CTA-PAIR version (cta_group::2):  D[M,N] = epilogue(A[M,K] * Bw[N,K]^T)
//
// Two CTAs of a cluster (one TPC) compute a 256 x BN output tile together: each CTA loads ITS 128 rows of A and ITS
// half of the B (weight) tile, the leader CTA issues tcgen05.mma.cta_group::2 (M=256) which reads both halves, and each
// CTA's tensor memory receives its own 128 accumulator rows.  Per SM that halves the weight bytes pulled from L2 per
// FLOP.  Why: the 1-CTA kernel (gemm_tc.cu) is latency-bound on its operand ring — 3 x 48 KB in flight is less than the
// bandwidth-delay product at the MMA rate (ncu r1: tensor pipe 41 %, L2->L1 fabric only 23 % of peak); here a stage is
// 32 KB for the same MMA work and 4 stages fit.
//   warp 0 (both CTAs)  TMA producer: cp.async.bulk.tensor .cta_group::2, completion bytes land on the LEADER's barrier
//   warp 1 (leader)     MMA issuer; tcgen05.commit ...multicast::cluster releases the smem stage in BOTH CTAs and
//                       publishes the accumulator to BOTH epilogues
//   warps 2..17 (both)  epilogue as in gemm_tc.cu on the CTA's own 128 rows; one thread per CTA hands the accumulator
//                       stage back to the leader's MMA warp (remote mbarrier arrive from the peer)
// Every mbarrier wait is bounded (2 s) and traps instead of hanging.
#include "tc_common.cuh"

namespace nfdpm {

constexpr int T2_BM = 128;                 // rows per CTA (256 per pair)
constexpr int T2_BK = 64;
constexpr int T2_STAGES = 4;
constexpr int T2_A_BYTES = T2_BM * T2_BK * 2;           // 16 KB
constexpr int T2_B_BYTES_MAX = 128 * T2_BK * 2;         // 16 KB: this CTA's half of the B tile (BN <= 256)
constexpr int T2_STAGE_BYTES = T2_A_BYTES + T2_B_BYTES_MAX;
constexpr int T2_EPI_WARPS = 16;           // four per TMEM lane quadrant (as gemm_tc.cu)
constexpr int T2_EPQ = T2_EPI_WARPS / 4;
constexpr int T2_THREADS = 64 + 32 * T2_EPI_WARPS;
__device__ __forceinline__ void epi_barrier2() { asm volatile("bar.sync 1, %0;" ::"n"(32 * T2_EPI_WARPS) : "memory"); }
constexpr int T2_ACC_COLS = 256;
constexpr int T2_CSTAGE_BYTES = 64 * 1024;

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the mbarrier at the same shared-memory offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar, uint32_t cta) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(bar), "r"(cta));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
// 2-CTA TMA load: the data lands in THIS CTA's shared memory, the transaction bytes on the LEADER's barrier
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  uint32_t lead_bar;                                 // the barrier at this offset in CTA 0 of the pair
  asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(lead_bar) : "r"(bar));
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(lead_bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive (once the MMAs issued so far have completed) on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_2sm(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((uint16_t)3)
               : "memory");
}

template <int EPI, typename OutT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(T2_THREADS, 1)
    gemm_nt_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                       const __grid_constant__ CUtensorMap tmD, int M, int N, int K, int BN,
                       const float* __restrict__ ep_scale, const float* __restrict__ ep_bias) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * T2_STAGES + 4];
  __shared__ uint32_t s_tmem_base;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const uint32_t ring = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t cstage = ring + T2_STAGES * T2_STAGE_BYTES;
  float* s_ep = reinterpret_cast<float*>(smem_raw + (ring - smem_u32(smem_raw)) + T2_STAGES * T2_STAGE_BYTES + T2_CSTAGE_BYTES);
  const uint32_t bar_full = smem_u32(&bars[0]), bar_empty = smem_u32(&bars[T2_STAGES]);
  const uint32_t bar_tfull = smem_u32(&bars[2 * T2_STAGES]), bar_tempty = smem_u32(&bars[2 * T2_STAGES + 2]);

  const int num_n = (N + BN - 1) / BN;
  const int num_m2 = (M + 2 * T2_BM - 1) / (2 * T2_BM);       // 256-row pair tiles
  const int num_tiles = num_m2 * num_n;
  const int num_kb = K / T2_BK;
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
  const int BNh = BN >> 1;                                     // B rows loaded by this CTA

  const int n_pad = num_n * BN;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmD);
    for (int s = 0; s < T2_STAGES; ++s) {
      mbar_init(bar_full + 8 * s, 1);      // leader only: its producer's arrive.expect_tx (bytes of BOTH CTAs)
      mbar_init(bar_empty + 8 * s, 1);     // each CTA: one multicast commit from the leader's MMA warp
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar_tfull + 8 * a, 1);     // each CTA: one multicast commit
      mbar_init(bar_tempty + 8 * a, 2);    // leader only: one arrive per CTA of the pair
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_2sm(smem_u32(&s_tmem_base), 512);
  pdl_trigger();
  pdl_wait();
  if (EPI == NFDPM_EPI_ACTNORM_RELU) {
    for (int i = threadIdx.x; i < n_pad; i += T2_THREADS) {
      const float e = (i < N) ? expf(ep_scale[i]) : 0.f;
      s_ep[i] = e;
      s_ep[n_pad + i] = (i < N) ? e * ep_bias[i] : 0.f;
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                       // barriers of both CTAs initialised, TMEM allocated in both
  tc_fence_after();
  const uint32_t tmem_base = s_tmem_base;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    int stage = 0;
    uint32_t phase = 0;
    const uint32_t tx_pair = 2u * (uint32_t)(T2_A_BYTES + BNh * T2_BK * 2);
    for (int tile = pair; tile < num_tiles; tile += num_pairs) {
      const int m_blk = tile / num_n, n_blk = tile - m_blk * num_n;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(bar_empty + 8 * stage, phase ^ 1);          // this CTA's copy of the stage is free
        if (lane == 0) {
          const uint32_t sa = ring + stage * T2_STAGE_BYTES, sb = sa + T2_A_BYTES;
          if (leader) mbar_arrive_expect_tx(bar_full + 8 * stage, tx_pair);
          tma_load_2d_2sm(sa, &tmA, kb * T2_BK, m_blk * 2 * T2_BM + (int)rank * T2_BM, bar_full + 8 * stage);
          tma_load_2d_2sm(sb, &tmB, kb * T2_BK, n_blk * BN + (int)rank * BNh, bar_full + 8 * stage);
        }
        __syncwarp();
        if (++stage == T2_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (leader) {
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      const uint32_t idesc = make_idesc(2 * T2_BM, BN);
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        mbar_wait(bar_tempty + 8 * acc, acc_phase ^ 1);       // both epilogues have drained this accumulator stage
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * T2_ACC_COLS;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(bar_full + 8 * stage, phase);             // the bytes of BOTH CTAs have landed
          tc_fence_after();
          if (lane == 0) {
            const uint32_t sa = ring + stage * T2_STAGE_BYTES, sb = sa + T2_A_BYTES;
            const uint64_t adesc = make_smem_desc(sa), bdesc = make_smem_desc(sb);
#pragma unroll
            for (int k = 0; k < T2_BK / 16; ++k)
              umma_bf16_2sm(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
            umma_commit_2sm(bar_empty + 8 * stage);           // frees the stage in both CTAs
            if (kb == num_kb - 1) umma_commit_2sm(bar_tfull + 8 * acc);
          }
          __syncwarp();
          if (++stage == T2_STAGES) { stage = 0; phase ^= 1; }
        }
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else {
    // ===================== epilogue (warps 2..17, both CTAs, own 128 rows) =====================
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int n_chunks = BN >> 4;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = pair; tile < num_tiles; tile += num_pairs) {
      const int m_blk = tile / num_n, n_blk = tile - m_blk * num_n;
      mbar_wait(bar_tfull + 8 * acc, acc_phase);
      tc_fence_after();
      const int trow = q * 32 + lane;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * T2_ACC_COLS;
      if (warp == 2 && lane == 0) tma_store_wait_read();
      epi_barrier2();
      auto process = [&](const uint32_t (&r)[16], int c0) {
        const int n0 = n_blk * BN + c0;
        float v[16];
        if (EPI == NFDPM_EPI_ACTNORM_RELU) {
          const float4* pe = reinterpret_cast<const float4*>(s_ep + n0);
          const float4* pb = reinterpret_cast<const float4*>(s_ep + n_pad + n0);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 e = pe[j], b = pb[j];
            v[4 * j + 0] = fmaxf(0.f, fmaf(__uint_as_float(r[4 * j + 0]), e.x, b.x));
            v[4 * j + 1] = fmaxf(0.f, fmaf(__uint_as_float(r[4 * j + 1]), e.y, b.y));
            v[4 * j + 2] = fmaxf(0.f, fmaf(__uint_as_float(r[4 * j + 2]), e.z, b.z));
            v[4 * j + 3] = fmaxf(0.f, fmaf(__uint_as_float(r[4 * j + 3]), e.w, b.w));
          }
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
        }
        TcStage<OutT>::put16(cstage, trow, c0, v);
      };
      uint32_t ra[16], rb[16];
      int ch = half;
      if (ch < n_chunks) tmem_ld16(taddr + ch * 16, ra);
      while (ch < n_chunks) {
        tmem_ld_wait();
        if (ch + T2_EPQ < n_chunks) tmem_ld16(taddr + (ch + T2_EPQ) * 16, rb);
        process(ra, ch * 16);
        ch += T2_EPQ;
        if (ch >= n_chunks) break;
        tmem_ld_wait();
        if (ch + T2_EPQ < n_chunks) tmem_ld16(taddr + (ch + T2_EPQ) * 16, ra);
        process(rb, ch * 16);
        ch += T2_EPQ;
      }
      tc_fence_before();
      fence_proxy_async();
      epi_barrier2();                                         // every TMEM read and staging write of this CTA is done
      if (warp == 2 && lane == 0) {
        mbar_arrive_cluster(bar_tempty + 8 * acc, 0);         // hand the accumulator stage back to the leader's MMA warp
        constexpr int CPB = TcStage<OutT>::kColsPerBox;
        const int n_boxes = (BN + CPB - 1) / CPB;
        for (int j = 0; j < n_boxes; ++j)
          tma_store_2d(&tmD, cstage + (uint32_t)j * 16384u, n_blk * BN + j * CPB, m_blk * 2 * T2_BM + (int)rank * T2_BM);
        tma_store_commit();
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (warp == 2 && lane == 0) tma_store_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                       // no CTA frees tensor memory / exits while its peer may still address it
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, 512);
  }
}

// ---------------------------------------------------------------- host side
template <int EPI, typename OutT>
static int launch_tc2(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmD, int M, int N, int K, int BN,
                      const float* es, const float* eb, int grid, size_t smem, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    NFDPM_CUDA(cudaFuncSetAttribute(gemm_nt_tc2_kernel<EPI, OutT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    attr_set = true;
  }
  NFDPM_CUDA(launch_pdl(gemm_nt_tc2_kernel<EPI, OutT>, dim3(grid), dim3(T2_THREADS), smem, st, tmA, tmB, tmD, M, N, K, BN, es,
                        eb));
  return 0;
}

// returns -1 when the shape is not handled by the pair kernel (the caller takes gemm_nt_tc)
int gemm_nt_tc2(const void* A, int64_t lda, const void* Bw, int64_t ldb, void* D, int64_t ldd, int M, int N, int K,
                int out_dtype, int epilogue, const float* ep_scale, const float* ep_bias, cudaStream_t st) {
  if (M < 2 * T2_BM || N % 32 != 0) return -1;
  const int bn_max = (out_dtype == NFDPM_F32) ? 128 : 256;
  const int cpb = (out_dtype == NFDPM_F32) ? 32 : 64;
  const int nblk = (N + bn_max - 1) / bn_max;
  const int BN = (nblk == 1) ? N : ((N + nblk - 1) / nblk + cpb - 1) / cpb * cpb;
  if (BN % 32 != 0 || BN > 256 || N % BN != 0) return -1;      // each CTA loads BN/2 rows (multiple of 16: UMMA N % 16)
  CUtensorMap tmA, tmB, tmD;
  if (make_map(&tmA, A, M, K, lda, T2_BM)) return 1;
  if (make_map(&tmB, Bw, N, K, ldb, BN / 2)) return 1;
  if (make_map(&tmD, D, M, N, ldd, T2_BM, out_dtype == NFDPM_F32)) return 1;
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    NFDPM_CUDA(cudaGetDevice(&dev));
    NFDPM_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  }
  const int tiles = ((M + 2 * T2_BM - 1) / (2 * T2_BM)) * (N / BN);
  const int pairs = tiles < sms / 2 ? tiles : sms / 2;
  const int grid = 2 * pairs;
  const size_t n_pad = (size_t)N;
  const size_t smem = 1024 + (size_t)T2_STAGES * T2_STAGE_BYTES + T2_CSTAGE_BYTES +
                      (epilogue == NFDPM_EPI_ACTNORM_RELU ? 2 * n_pad * 4 : 0);
#define GO(EPI, T) return launch_tc2<EPI, T>(tmA, tmB, tmD, M, N, K, BN, ep_scale, ep_bias, grid, smem, st)
  if (out_dtype == NFDPM_F32) {
    if (epilogue == NFDPM_EPI_RAW) GO(NFDPM_EPI_RAW, float); else GO(NFDPM_EPI_ACTNORM_RELU, float);
  } else {
    if (epilogue == NFDPM_EPI_RAW) GO(NFDPM_EPI_RAW, __nv_bfloat16); else GO(NFDPM_EPI_ACTNORM_RELU, __nv_bfloat16);
  }
#undef GO
}

}  // namespace nfdpm
