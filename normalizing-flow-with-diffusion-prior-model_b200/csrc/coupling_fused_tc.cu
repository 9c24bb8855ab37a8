// Fused coupling network on tcgen05 (bf16): the three convolutions of normalizing_flow/utils.py:83-89 as ONE kernel
//   h1 = relu(actnorm1(A1 * W1^T))      A1 [M,K1p]  im2col rows (bf16), W1 [512,K1p]
//   h2 = relu(actnorm2(h1 * W2^T))      W2 [512,512]
//   pm = h2 * W3^T                      W3 [ldp,512] "taps as N" rows -> pm [M,ldp] fp32 (consumed by nfdpm_flow_boundary)
// per 128-row tile, persistent CTAs.  h1 and h2 (128 x 512 bf16 = 128 KB) never leave the SM: the epilogue writes them
// into shared memory in the 128B-swizzled K-major layout that tcgen05.mma reads as its A operand; only the weights are
// streamed (TMA, 3 x 32 KB ring, 256 weight rows x 64 K per stage) and only A1 / pm touch global memory.
//
//   warp 0      TMA producer: A1 tile (into the activation buffer, free during GEMM1), then W1 / W2 / W3 stages
//   warp 1      MMA issuer: GEMM1 (N=2x256) -> wait h1 -> GEMM2 (N=2x256) -> wait h2 -> GEMM3 (N=n3 x BN3)
//   warps 2..9  epilogues: TMEM -> ActNorm+ReLU -> bf16 -> swizzled smem (E1, E2); TMEM -> fp32 global rows (E3)
// TMEM: one 512-column accumulator reused by the three GEMMs (they are serially dependent).
// Every mbarrier wait is bounded (traps instead of hanging).
#include "tc_common.cuh"

namespace nfdpm {

constexpr int CF_BM = 128;
constexpr int CF_F = 512;                       // hidden width (coupling_net_n_features, reference default)
constexpr int CF_NST = 3;                       // weight ring stages
constexpr int CF_WSTAGE = 256 * 64 * 2;         // 32 KB: 256 weight rows x 64 K (bf16)
constexpr int CF_ACT_BYTES = CF_BM * CF_F * 2;  // 128 KB: 8 K-blocks of [128 rows][64 cols] bf16, 128B swizzle
constexpr int CF_THREADS = 64 + 32 * TC_EPI_WARPS;

// barrier indices
enum { B_FULL = 0, B_EMPTY = CF_NST, B_AFULL = 2 * CF_NST, B_ACTFREE, B_D1, B_D2, B_D3, B_H1, B_H2, B_D3EMPTY, B_COUNT };

__device__ __forceinline__ void put16_bf16(uint32_t base, int row, int c0, const float (&v)[16]) {
  uint32_t w[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    __nv_bfloat162 t = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    w[i] = *reinterpret_cast<uint32_t*>(&t);
  }
  const uint32_t blk = base + (uint32_t)(c0 >> 6) * 16384u + (uint32_t)row * 128u;
  const int u0 = (c0 & 63) >> 3;
  st_shared_v4(blk + (uint32_t)(((u0 + 0) ^ (row & 7)) << 4), w[0], w[1], w[2], w[3]);
  st_shared_v4(blk + (uint32_t)(((u0 + 1) ^ (row & 7)) << 4), w[4], w[5], w[6], w[7]);
}

__global__ void __launch_bounds__(CF_THREADS, 1)
coupling_fused_tc_kernel(const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmW1,
                         const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmW3,
                         float* __restrict__ pm, int64_t ldp, int M, int nk1, int n3, int BN3,
                         const float* __restrict__ ep /* [4][512]: e1, e1*b1, e2, e2*b2 (nfdpm_fold_actnorm) */) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[B_COUNT];
  __shared__ uint32_t s_tmem_base;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t ring = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t act = ring + CF_NST * CF_WSTAGE;
  auto bar = [&](int i) -> uint32_t { return smem_u32(&bars[i]); };
  const int num_tiles = (M + CF_BM - 1) / CF_BM;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA1);
    tma_prefetch_desc(&tmW1);
    tma_prefetch_desc(&tmW2);
    tma_prefetch_desc(&tmW3);
    for (int s = 0; s < CF_NST; ++s) {
      mbar_init(bar(B_FULL + s), 1);
      mbar_init(bar(B_EMPTY + s), 1);
    }
    mbar_init(bar(B_AFULL), 1);
    mbar_init(bar(B_ACTFREE), 1);
    mbar_init(bar(B_D1), 1);
    mbar_init(bar(B_D2), 1);
    mbar_init(bar(B_D3), 1);
    mbar_init(bar(B_H1), 32 * TC_EPI_WARPS);
    mbar_init(bar(B_H2), 32 * TC_EPI_WARPS);
    mbar_init(bar(B_D3EMPTY), 32 * TC_EPI_WARPS);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(&s_tmem_base), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = s_tmem_base;

  if (warp == 0) {
    // ===================== TMA producer =====================
    int stage = 0;
    uint32_t phase = 0;
    auto load_w = [&](const CUtensorMap* map, int k0, int r0, uint32_t bytes) {
      mbar_wait(bar(B_EMPTY + stage), phase ^ 1);
      if (lane == 0) {
        mbar_arrive_expect_tx(bar(B_FULL + stage), bytes);
        tma_load_2d(ring + stage * CF_WSTAGE, map, k0, r0, bar(B_FULL + stage));
      }
      __syncwarp();
      if (++stage == CF_NST) { stage = 0; phase ^= 1; }
    };
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      // the activation buffer is free once GEMM3 of the previous tile has finished reading h2
      mbar_wait(bar(B_ACTFREE), (it & 1) ^ 1);
      if (lane == 0) {
        mbar_arrive_expect_tx(bar(B_AFULL), (uint32_t)nk1 * 16384u);
        for (int kb = 0; kb < nk1; ++kb) tma_load_2d(act + kb * 16384u, &tmA1, kb * 64, tile * CF_BM, bar(B_AFULL));
      }
      __syncwarp();
      for (int kb = 0; kb < nk1; ++kb)
        for (int h = 0; h < 2; ++h) load_w(&tmW1, kb * 64, h * 256, CF_WSTAGE);
      for (int kb = 0; kb < 8; ++kb)
        for (int h = 0; h < 2; ++h) load_w(&tmW2, kb * 64, h * 256, CF_WSTAGE);
      for (int c = 0; c < n3; ++c)
        for (int kb = 0; kb < 8; ++kb) load_w(&tmW3, kb * 64, c * BN3, (uint32_t)BN3 * 128u);
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    int stage = 0;
    uint32_t phase = 0;
    const uint32_t idesc256 = make_idesc(CF_BM, 256), idesc3 = make_idesc(CF_BM, BN3);
    // one weight stage: 4 x (K=16) MMAs, A = activation K-block kb, B = ring stage
    auto mma_stage = [&](int kb, uint32_t tmem_d, uint32_t idesc, bool first_k) {
      mbar_wait(bar(B_FULL + stage), phase);
      tc_fence_after();
      if (lane == 0) {
        const uint64_t adesc = make_smem_desc(act + kb * 16384u), bdesc = make_smem_desc(ring + stage * CF_WSTAGE);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (first_k && k == 0) ? 0u : 1u);
        umma_commit(bar(B_EMPTY + stage));
      }
      __syncwarp();
      if (++stage == CF_NST) { stage = 0; phase ^= 1; }
    };
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const uint32_t par = it & 1;
      mbar_wait(bar(B_D3EMPTY), par ^ 1);     // previous tile's E3 has drained the accumulator
      mbar_wait(bar(B_AFULL), par);           // A1 tile landed
      tc_fence_after();
      for (int kb = 0; kb < nk1; ++kb)
        for (int h = 0; h < 2; ++h) mma_stage(kb, tmem_base + h * 256, idesc256, kb == 0);
      if (lane == 0) umma_commit(bar(B_D1));
      __syncwarp();
      mbar_wait(bar(B_H1), par);              // h1 written to smem, accumulator drained
      tc_fence_after();
      for (int kb = 0; kb < 8; ++kb)
        for (int h = 0; h < 2; ++h) mma_stage(kb, tmem_base + h * 256, idesc256, kb == 0);
      if (lane == 0) umma_commit(bar(B_D2));
      __syncwarp();
      mbar_wait(bar(B_H2), par);
      tc_fence_after();
      for (int c = 0; c < n3; ++c)
        for (int kb = 0; kb < 8; ++kb) mma_stage(kb, tmem_base + c * BN3, idesc3, kb == 0);
      if (lane == 0) {
        umma_commit(bar(B_D3));
        umma_commit(bar(B_ACTFREE));
      }
      __syncwarp();
    }
  } else {
    // ===================== epilogue warps =====================
    const int q = warp & 3, half = (warp - 2) >> 2;
    const int trow = q * 32 + lane;
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    // hidden-layer epilogue: 32 chunks of 16 columns, alternating between the two warps of the quadrant
    auto hidden = [&](const float* pe, const float* pb) {
      auto process = [&](const uint32_t (&r)[16], int c0) {
        float v[16];
        const float4* e4 = reinterpret_cast<const float4*>(pe + c0);
        const float4* b4 = reinterpret_cast<const float4*>(pb + c0);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 e = __ldg(e4 + j), b = __ldg(b4 + j);     // 8 KB of parameters: L1 resident
          v[4 * j + 0] = fmaxf(0.f, fmaf(__uint_as_float(r[4 * j + 0]), e.x, b.x));
          v[4 * j + 1] = fmaxf(0.f, fmaf(__uint_as_float(r[4 * j + 1]), e.y, b.y));
          v[4 * j + 2] = fmaxf(0.f, fmaf(__uint_as_float(r[4 * j + 2]), e.z, b.z));
          v[4 * j + 3] = fmaxf(0.f, fmaf(__uint_as_float(r[4 * j + 3]), e.w, b.w));
        }
        put16_bf16(act, trow, c0, v);
      };
      uint32_t ra[16], rb[16];
      constexpr int n_chunks = CF_F / 16;
      int ch = half;
      tmem_ld16(taddr + ch * 16, ra);
      while (ch < n_chunks) {
        tmem_ld_wait();
        if (ch + 2 < n_chunks) tmem_ld16(taddr + (ch + 2) * 16, rb);
        process(ra, ch * 16);
        ch += 2;
        if (ch >= n_chunks) break;
        tmem_ld_wait();
        if (ch + 2 < n_chunks) tmem_ld16(taddr + (ch + 2) * 16, ra);
        process(rb, ch * 16);
        ch += 2;
      }
      tc_fence_before();        // TMEM reads done before the MMA warp overwrites the accumulator
      fence_proxy_async();      // smem writes visible to the tensor core (async proxy)
    };
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const uint32_t par = it & 1;
      mbar_wait(bar(B_D1), par);
      tc_fence_after();
      hidden(ep, ep + CF_F);
      mbar_arrive(bar(B_H1));
      mbar_wait(bar(B_D2), par);
      tc_fence_after();
      hidden(ep + 2 * CF_F, ep + 3 * CF_F);
      mbar_arrive(bar(B_H2));
      mbar_wait(bar(B_D3), par);
      tc_fence_after();
      {
        const int row = tile * CF_BM + trow;
        const int n_chunks = (n3 * BN3) >> 4;
        float* prow = pm + (int64_t)row * ldp;
        uint32_t r[16];
        for (int ch = half; ch < n_chunks; ch += 2) {
          tmem_ld16(taddr + ch * 16, r);
          tmem_ld_wait();
          const int n0 = ch * 16;
          if (row < M && n0 + 15 < ldp) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
              *reinterpret_cast<float4*>(prow + n0 + 4 * j) =
                  make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]),
                              __uint_as_float(r[4 * j + 3]));
          } else if (row < M) {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (n0 + j < ldp) prow[n0 + j] = __uint_as_float(r[j]);
          }
        }
        tc_fence_before();
      }
      mbar_arrive(bar(B_D3EMPTY));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace nfdpm

using namespace nfdpm;

extern "C" int nfdpm_coupling_fused(const void* a1, int64_t lda1, const void* w1, const void* w2, const void* w3,
                                    float* pm, int64_t ldp, int M, int K1p, const float* ep,
                                    nfdpm_stream_t stream) {
  NFDPM_REQUIRE(a1 && w1 && w2 && w3 && pm && ep, "nfdpm_coupling_fused: null pointer");
  NFDPM_REQUIRE((uintptr_t)ep % 16 == 0, "nfdpm_coupling_fused: ep must be 16-byte aligned");
  NFDPM_REQUIRE(M > 0 && K1p > 0 && K1p % 64 == 0 && K1p <= 512, "nfdpm_coupling_fused: K1p=%d must be a multiple of 64, <= 512", K1p);
  NFDPM_REQUIRE(lda1 >= K1p && lda1 % 8 == 0, "nfdpm_coupling_fused: bad lda1");
  NFDPM_REQUIRE(ldp > 0 && ldp % 16 == 0 && ldp <= 512, "nfdpm_coupling_fused: ldp=%lld must be a multiple of 16, <= 512 "
                "(wider coupling layers use the unfused GEMMs)", (long long)ldp);
  NFDPM_REQUIRE(((uintptr_t)a1 % 16 == 0) && ((uintptr_t)w1 % 16 == 0) && ((uintptr_t)w2 % 16 == 0) &&
                ((uintptr_t)w3 % 16 == 0) && ((uintptr_t)pm % 16 == 0), "nfdpm_coupling_fused: operands must be 16-byte aligned");
  const int n3 = (int)((ldp + 255) / 256);
  const int BN3 = (int)(((ldp + n3 - 1) / n3 + 15) / 16 * 16);
  CUtensorMap tmA1, tmW1, tmW2, tmW3;
  if (make_map(&tmA1, a1, M, K1p, lda1, CF_BM)) return 1;
  if (make_map(&tmW1, w1, CF_F, K1p, K1p, 256)) return 1;
  if (make_map(&tmW2, w2, CF_F, CF_F, CF_F, 256)) return 1;
  if (make_map(&tmW3, w3, ldp, CF_F, CF_F, BN3)) return 1;
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    NFDPM_CUDA(cudaGetDevice(&dev));
    NFDPM_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  }
  const int tiles = (M + CF_BM - 1) / CF_BM;
  const int grid = tiles < sms ? tiles : sms;
  const size_t smem = 1024 + (size_t)CF_NST * CF_WSTAGE + CF_ACT_BYTES;
  static bool attr_set = false;
  if (!attr_set) {
    NFDPM_CUDA(cudaFuncSetAttribute(coupling_fused_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  coupling_fused_tc_kernel<<<grid, CF_THREADS, smem, as_stream(stream)>>>(tmA1, tmW1, tmW2, tmW3, pm, ldp, M, K1p / 64, n3,
                                                                         BN3, ep);
  NFDPM_CHECK_LAUNCH("coupling_fused_tc_kernel");
  return 0;
}

namespace nfdpm {
__global__ void fold_actnorm_kernel(const float* __restrict__ scale, const float* __restrict__ bias,
                                    float* __restrict__ e_out, float* __restrict__ eb_out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const float e = expf(scale[i]);
    e_out[i] = e;
    eb_out[i] = e * bias[i];
  }
}
}  // namespace nfdpm

extern "C" int nfdpm_fold_actnorm(const float* scale, const float* bias, float* e_out, float* eb_out, int n,
                                  nfdpm_stream_t stream) {
  NFDPM_REQUIRE(scale && bias && e_out && eb_out && n > 0, "nfdpm_fold_actnorm: bad arguments");
  fold_actnorm_kernel<<<(n + 255) / 256, 256, 0, as_stream(stream)>>>(scale, bias, e_out, eb_out, n);
  NFDPM_CHECK_LAUNCH("fold_actnorm_kernel");
  return 0;
}
