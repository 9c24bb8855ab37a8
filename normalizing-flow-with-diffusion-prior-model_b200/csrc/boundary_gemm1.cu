// Step boundary + FIRST coupling GEMM of the next StepFlow in one launch (tensor-core mode):
//     [flow_boundary: coupling of step k (or plain / squeeze source), next fused ActNorm + 1x1 conv, NCHW sink]
//     -> im2col rows of the new state built DIRECTLY in shared memory as the tcgen05 A operand (K-major, 128B swizzle)
//     -> GEMM1  h1 = relu(actnorm(A1 x W1^T))  (3x3 conv of the coupling network, transforms.py:169 via utils.py:47-69)
// One CTA per image, as nfdpm_flow_boundary.  Removes one launch per StepFlow (each launch of the chain costs ~5 us of
// fill / drain at these sizes, tools/gemm_counters.py) and the global round trip of the im2col rows; the W1 boxes are
// fetched by TMA while the CTA does the boundary arithmetic.
//   warp 0      TMA producer for W1 (boxes of 256 output channels x 64 k), ring or fully resident
//   warp 1      MMA issuer: per 128-row tile of the image 2 N-halves x K1p/64 k-blocks into 512 TMEM columns
//   warps 2..17 epilogue: TMEM -> ActNorm/ReLU -> bf16 -> swizzled staging -> TMA store of the image's rows of h1
// All 18 warps run the boundary arithmetic first (boundary_body.cuh).
#include <algorithm>

#include "tc_common.cuh"
#include "boundary_body.cuh"

namespace nfdpm {

constexpr int BG_EPI_WARPS = 16;
constexpr int BG_EPQ = BG_EPI_WARPS / 4;
constexpr int BG_THREADS = 64 + 32 * BG_EPI_WARPS;
constexpr int BG_BK = 64;
constexpr int BG_MAX_STAGES = 4;
constexpr int BG_BOX_BYTES = 256 * BG_BK * 2;           // one W1 box: 256 output channels x 64 k = 32 KB
constexpr int BG_SMEM_LIMIT = 225 * 1024;
__device__ __forceinline__ void bg_epi_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(32 * BG_EPI_WARPS) : "memory"); }

struct BG1Args {
  BoundaryArgs bd;
  const float *s1, *b1;              // inner ActNorm of the GEMM1 network (raw log-scale, bias) [512]
  int nkb;                           // K1p / 64
  int n_mt, rows;                    // 128-row tiles per image (1 or 2), valid rows per tile
  int stages, resident;              // W1 ring depth; resident: all 2*nkb boxes stay in shared memory
  int off_ring, off_stage, off_ep, off_body;   // byte offsets from the 1024-aligned base (A tile at 0)
};

template <bool COUPLING>
__global__ void __launch_bounds__(BG_THREADS, 1)
boundary_gemm1_kernel(const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmH1, const BG1Args a) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * BG_MAX_STAGES + 2];
  __shared__ uint32_t s_tmem_base;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
  const uint32_t abase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base = smem_raw + (abase - smem_u32(smem_raw));
  const uint32_t ring = abase + a.off_ring, cstage = abase + a.off_stage;
  float* ep_s = reinterpret_cast<float*>(base + a.off_ep);          // [e][e*b] x 512
  float* body_s = reinterpret_cast<float*>(base + a.off_body);
  const uint32_t bar_full = smem_u32(&bars[0]), bar_empty = smem_u32(&bars[BG_MAX_STAGES]);
  const uint32_t bar_tfull = smem_u32(&bars[2 * BG_MAX_STAGES]), bar_tempty = bar_tfull + 8;
  const int b = blockIdx.x;
  const int nkb = a.nkb, n_box = 2 * nkb;                           // W1 boxes per 128-row tile: (N half, k block)

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmW1);
    tma_prefetch_desc(&tmH1);
    for (int s = 0; s < a.stages; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    mbar_init(bar_tfull, 1);
    mbar_init(bar_tempty, BG_EPI_WARPS);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(&s_tmem_base), 512);
  pdl_trigger();
  pdl_wait();
  for (int i = tid; i < 512; i += BG_THREADS) {                     // y = max(0, e*acc + e*b)
    const float e = expf(a.s1[i]);
    ep_s[i] = e;
    ep_s[512 + i] = e * a.b1[i];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = s_tmem_base;

  // the first W1 boxes travel while the CTA does the boundary arithmetic
  const int first = min(a.stages, a.resident ? n_box : a.n_mt * n_box);
  if (warp == 0 && lane == 0) {
    for (int i = 0; i < first; ++i) {
      const int bi = i % n_box;
      mbar_arrive_expect_tx(bar_full + 8 * i, BG_BOX_BYTES);
      tma_load_2d(ring + i * BG_BOX_BYTES, &tmW1, (bi % nkb) * BG_BK, (bi / nkb) * 256, bar_full + 8 * i);
    }
  }

  // ===================== step boundary of image b; im2col rows -> A tile at `base` =====================
  flow_boundary_body<COUPLING, __nv_bfloat16>(a.bd, b, body_s, tid, BG_THREADS, nullptr, 0, base);
  fence_proxy_async();                                              // generic-proxy smem writes -> tensor-core reads
  __syncthreads();

  const int total = a.n_mt * n_box;                                 // boxes consumed by the MMA warp
  if (warp == 0) {
    // ===================== TMA producer: remaining W1 boxes (ring mode) =====================
    if (!a.resident) {
      int stage = first % a.stages;
      uint32_t phase = (first / a.stages) & 1;
      for (int i = first; i < total; ++i) {
        const int bi = i % n_box;
        mbar_wait(bar_empty + 8 * stage, phase ^ 1);
        if (lane == 0) {
          mbar_arrive_expect_tx(bar_full + 8 * stage, BG_BOX_BYTES);
          tma_load_2d(ring + stage * BG_BOX_BYTES, &tmW1, (bi % nkb) * BG_BK, (bi / nkb) * 256, bar_full + 8 * stage);
        }
        __syncwarp();
        if (++stage == a.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const uint32_t idesc = make_idesc(128, 256);
    int stage = 0;
    uint32_t phase = 0;
    for (int mt = 0; mt < a.n_mt; ++mt) {
      if (mt > 0) {                                                 // the epilogue has drained the accumulators
        mbar_wait(bar_tempty, (uint32_t)((mt - 1) & 1));
        tc_fence_after();
      }
      for (int bi = 0; bi < n_box; ++bi) {
        const int nh = bi / nkb, kb = bi - nh * nkb;
        const int st = a.resident ? bi : stage;
        mbar_wait(bar_full + 8 * st, a.resident ? 0u : phase);
        tc_fence_after();
        if (lane == 0) {
          const uint64_t adesc = make_smem_desc(abase + (uint32_t)(mt * nkb + kb) * 16384u);
          const uint64_t bdesc = make_smem_desc(ring + (uint32_t)st * BG_BOX_BYTES);
#pragma unroll
          for (int k = 0; k < BG_BK / 16; ++k)
            umma_bf16(tmem_base + nh * 256, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          if (!a.resident) umma_commit(bar_empty + 8 * stage);
          if (bi == n_box - 1) umma_commit(bar_tfull);
        }
        __syncwarp();
        if (!a.resident && ++stage == a.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue: TMEM -> ActNorm/ReLU -> bf16 -> staging -> TMA store =====================
    const int q = warp & 3, sub = (warp - 2) >> 2;
    const int trow = q * 32 + lane;
    const bool row_ok = trow < a.rows;                              // rows >= P of a short image hold garbage
    const uint32_t box_bytes = (uint32_t)a.rows * 128u;             // one store box: `rows` rows x 64 bf16
    const int P = a.bd.H * a.bd.W;
    for (int mt = 0; mt < a.n_mt; ++mt) {
      mbar_wait(bar_tfull, (uint32_t)(mt & 1));
      tc_fence_after();
      for (int nh = 0; nh < 2; ++nh) {
        // the previous TMA stores must have finished READING the staging tile before it is overwritten
        if (warp == 2 && lane == 0) tma_store_wait_read();
        bg_epi_barrier();
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + nh * 256;
        for (int c0 = sub * 16; c0 < 256; c0 += 16 * BG_EPQ) {
          uint32_t r[16];
          tmem_ld16(taddr + c0, r);
          tmem_ld_wait();
          if (row_ok) {
            const float* e = ep_s + nh * 256 + c0;
            uint32_t w[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float v0 = fmaxf(0.f, fmaf(__uint_as_float(r[2 * i]), e[2 * i], e[512 + 2 * i]));
              const float v1 = fmaxf(0.f, fmaf(__uint_as_float(r[2 * i + 1]), e[2 * i + 1], e[512 + 2 * i + 1]));
              __nv_bfloat162 t = __floats2bfloat162_rn(v0, v1);
              w[i] = *reinterpret_cast<uint32_t*>(&t);
            }
            const uint32_t box = cstage + (uint32_t)(c0 >> 6) * box_bytes + (uint32_t)trow * 128u;
            const int u0 = (c0 & 63) >> 3;
            st_shared_v4(box + (uint32_t)(((u0 + 0) ^ (trow & 7)) << 4), w[0], w[1], w[2], w[3]);
            st_shared_v4(box + (uint32_t)(((u0 + 1) ^ (trow & 7)) << 4), w[4], w[5], w[6], w[7]);
          }
        }
        if (nh == 1) {                                              // both halves read: hand tensor memory back
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_tempty);
        }
        fence_proxy_async();
        bg_epi_barrier();
        if (warp == 2 && lane == 0) {
          for (int j = 0; j < 4; ++j)
            tma_store_2d(&tmH1, cstage + (uint32_t)j * box_bytes, nh * 256 + j * 64, b * P + mt * 128);
          tma_store_commit();
        }
      }
    }
    if (warp == 2 && lane == 0) tma_store_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

static bool bg_plan(int C, int H, int W, int F, int64_t K1p, bool coupling, bool mix, BG1Args* a, size_t* smem) {
  const int P = H * W;
  if (F != 512 || K1p % BG_BK || K1p < 64 || K1p > 512 || C % 2) return false;
  if (!(P == 256 || (P <= 128 && P % 8 == 0))) return false;         // TMA store boxes of `rows` rows, swizzle atoms of 8 rows
  const int n_mt = P > 128 ? 2 : 1, rows = P > 128 ? 128 : P;
  const int nkb = (int)(K1p / BG_BK), n_box = 2 * nkb;
  const size_t a_bytes = (size_t)n_mt * nkb * 16384;
  const size_t body = (nfdpm_flow_boundary_smem(C, H, W, coupling ? 1 : 0, mix ? 1 : 0) + 15) & ~(size_t)15;
  const size_t stage_bytes = (size_t)rows * 128 * 4;                 // one N half: 4 boxes of 64 columns
  // first choice: every W1 box resident (needed for two tiles per image); else a ring whose memory (plus the A tile) is
  // reused as the output staging tile once all MMAs have retired
  for (int pass = 0; pass < 2; ++pass) {
    const bool resident = pass == 0;
    if (resident && n_box > BG_MAX_STAGES) continue;
    if (!resident && n_mt > 1) break;
    for (int stages = resident ? n_box : std::min(BG_MAX_STAGES, n_box); stages >= (resident ? n_box : 2); --stages) {
      const size_t ring = (size_t)stages * BG_BOX_BYTES;
      const size_t off_stage = resident ? a_bytes + ring : 0;
      const size_t gemm_end = resident ? off_stage + stage_bytes : std::max(a_bytes + ring, stage_bytes);
      const size_t off_ep = (gemm_end + 1023) & ~(size_t)1023;
      const size_t total = 1024 + off_ep + 4096 + body;
      if (total > (size_t)BG_SMEM_LIMIT) continue;
      if (a) {
        a->nkb = nkb; a->n_mt = n_mt; a->rows = rows; a->stages = stages; a->resident = resident ? 1 : 0;
        a->off_ring = (int)a_bytes; a->off_stage = (int)off_stage; a->off_ep = (int)off_ep; a->off_body = (int)off_ep + 4096;
      }
      if (smem) *smem = total;
      return true;
    }
  }
  return false;
}

}  // namespace nfdpm

using namespace nfdpm;

extern "C" int nfdpm_boundary_gemm1_ok(int C, int H, int W, int F, int64_t K1p) {
  return bg_plan(C, H, W, F, K1p, true, true, nullptr, nullptr) ? 1 : 0;
}

extern "C" int nfdpm_boundary_gemm1(const float* in, int64_t in_bs, int squeeze_in, const float* pm, int64_t ldp,
                                    const float* bias3, const float* logs3, float* ld_part, const float* mt,
                                    const float* beta, float* y, int64_t y_bs, float* xs, int64_t xs_bs, void* a1,
                                    const void* w1p, const float* s1, const float* b1, void* h1, int B, int C, int H, int W,
                                    int F, int64_t K1p, int inverse, nfdpm_stream_t stream) {
  NFDPM_REQUIRE(in && w1p && s1 && b1 && h1, "nfdpm_boundary_gemm1: null pointer");
  NFDPM_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0 && C % 2 == 0, "nfdpm_boundary_gemm1: bad shape B=%d C=%d H=%d W=%d", B, C, H, W);
  NFDPM_REQUIRE((mt == nullptr) == (beta == nullptr), "nfdpm_boundary_gemm1: mt/beta must both be set or both NULL");
  NFDPM_REQUIRE(pm == nullptr || (bias3 && logs3 && ldp >= 9 * (int64_t)C), "nfdpm_boundary_gemm1: coupling source needs bias3/logs3/ldp");
  NFDPM_REQUIRE(!squeeze_in || (C % 4 == 0 && in_bs % 2 == 0 && ((uintptr_t)in % 8) == 0),
                "nfdpm_boundary_gemm1: squeeze source needs C %% 4 == 0 and 8-byte alignment");
  NFDPM_REQUIRE(K1p >= 9 * (int64_t)(C / 2), "nfdpm_boundary_gemm1: K1p %lld < 9*C/2", (long long)K1p);
  NFDPM_REQUIRE((((uintptr_t)w1p | (uintptr_t)h1 | (uintptr_t)a1) % 16) == 0, "nfdpm_boundary_gemm1: operands must be 16-byte aligned");
  BG1Args a;
  size_t smem = 0;
  NFDPM_REQUIRE(bg_plan(C, H, W, F, K1p, pm != nullptr, mt != nullptr, &a, &smem),
                "nfdpm_boundary_gemm1: unsupported shape C=%d H=%d W=%d F=%d K1p=%lld (use nfdpm_flow_boundary + nfdpm_gemm_nt)",
                C, H, W, F, (long long)K1p);
  BoundaryArgs& d = a.bd;
  d.in = in; d.in_bs = in_bs; d.pm = pm; d.ldp = ldp; d.bias3 = bias3; d.logs3 = logs3; d.ld_part = ld_part;
  d.mt = mt; d.beta = beta; d.y = y; d.y_bs = y_bs; d.xs = xs; d.xs_bs = xs_bs; d.a1 = a1; d.lda1 = K1p;
  d.B = B; d.C = C; d.H = H; d.W = W; d.squeeze_in = squeeze_in; d.inverse = inverse;
  boundary_fill_div(d);
  a.s1 = s1; a.b1 = b1;
  const int64_t M = (int64_t)B * H * W;
  CUtensorMap tmW1, tmH1;
  if (make_map(&tmW1, w1p, F, K1p, K1p, 256)) return 1;
  if (make_map(&tmH1, h1, M, F, F, a.rows)) return 1;
  cudaStream_t st = as_stream(stream);
#define LAUNCH(CP)                                                                                                   \
  do {                                                                                                               \
    static bool attr_set = false;                                                                                    \
    if (!attr_set) {                                                                                                 \
      NFDPM_CUDA(cudaFuncSetAttribute(boundary_gemm1_kernel<CP>, cudaFuncAttributeMaxDynamicSharedMemorySize, BG_SMEM_LIMIT)); \
      attr_set = true;                                                                                               \
    }                                                                                                                \
    NFDPM_CUDA(launch_pdl(boundary_gemm1_kernel<CP>, dim3(B), dim3(BG_THREADS), smem, st, tmW1, tmH1, a));           \
  } while (0)
  if (pm != nullptr) LAUNCH(true); else LAUNCH(false);
#undef LAUNCH
  return 0;
}
