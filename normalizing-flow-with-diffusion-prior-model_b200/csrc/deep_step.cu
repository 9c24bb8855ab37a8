// One StepFlow of a DEEP level in ONE launch (north star items 3+4 for the levels where a 128-row GEMM tile holds
// whole images, H*W <= 64):
//     GEMM1 (3x3 conv as im2col rows) -> ActNorm/ReLU -> GEMM2 (1x1 conv) -> ActNorm/ReLU -> GEMM3 (ZeroConv, taps-as-N)
//     -> step boundary (affine coupling + log-det, next fused ActNorm/1x1 conv, NCHW + im2col sinks)
// (forward: transforms.py:169-184 / :80,132; inverse: :196-200 / :144,93).
//
// At these levels the four separate kernels are latency / L2-throughput bound (6-9 us each for M = 8192 / 2048 rows: every
// 128-row tile re-reads its A rows once per N tile and the intermediate rows make an L2 round trip per GEMM).  Here a
// thread-block CLUSTER of CS CTAs owns one 128-row tile (= 128/(H*W) whole images) and splits every GEMM along N; the
// intermediates travel CTA -> CTA through DISTRIBUTED SHARED MEMORY and never touch L2:
//     phase 0   A1 [128, K1p] (TMA) x W1 slice -> TMEM -> ActNorm/ReLU -> bf16 slice h1[:, r*NS:(r+1)*NS], written in the
//               UMMA K-major 128B-swizzle layout into the A-operand buffer of EVERY CTA of the cluster (st.shared::cluster)
//     phase 1   h1 [128, F] (resident A buffer) x W2 slice -> ... -> h2 slice -> every CTA's A buffer
//     phase 2   h2 x W3 slice (taps-as-N) -> fp32 pm rows, each row sent to the CTA that owns its image
//     phase 3   every CTA runs the step boundary (boundary_body.cuh) of its own image(s) with pm in shared memory
// Only the weight slices stream from L2 (TMA ring, prefetched across the phase boundaries: they do not depend on the
// peers); cross-CTA hand-over uses cluster-scope mbarriers:
//     xbar  (count CS)         every CTA's MMA warp commits to it in ALL CTAs (tcgen05.commit multicast): "every CTA has
//                              finished READING its A buffer / accumulating" -> the epilogues may overwrite the A buffers
//     ybar  (count CS*warps)   every epilogue warp of every CTA arrives on it in ALL CTAs once its slice is stored: "my A
//                              buffer holds the complete next operand and my accumulator has been drained"
//     pbar  (same count)       ditto for the pm rows of phase 2
// Optional global copies of h1 / h2 / pm serve as the activation stash of the training forward.
#include <algorithm>
#include <stdlib.h>

#include "tc_common.cuh"
#include "boundary_body.cuh"

namespace nfdpm {

constexpr int DS_EPI_WARPS = 16;     // four per TMEM lane quadrant
constexpr int DS_THREADS = 64 + 32 * DS_EPI_WARPS;   // warp 0 TMA producer, warp 1 MMA issuer, warps 2..17 epilogue
constexpr int DS_BK = 64;
constexpr int DS_MAX_STAGES = 8;
constexpr int DS_RING_BYTES = 64 * 1024;
constexpr int DS_SMEM_LIMIT = 225 * 1024;      // dynamic part; ~1.2 KB of static shared memory on top (227 KB per CTA)

struct DeepArgs {
  int M, K1p, F, ldp;                // rows, padded im2col width, hidden width, padded 9C
  int CS, NS, n3s, ipc;              // cluster size, N slice of GEMM1/2 and of GEMM3 (multiples of 16), images per CTA
  int stages, stage_bytes;           // weight ring
  int abuf_bytes, lds;               // A-operand buffer, row stride (floats) of the pm rows that later alias it
  const float *s1, *b1, *s2, *b2;    // inner ActNorm parameters (raw log-scale, bias)
  __nv_bfloat16 *h1, *h2;            // optional global copies [M, F] (training stash)
  float* pm; int64_t ld_pm;          // optional global copy [M, ld_pm]
  BoundaryArgs bd;
  long long* dbg;                    // optional per-CTA timeline [grid][16] (nfdpm_deep_step_debug; profiling only)
};

__device__ __forceinline__ uint32_t ds_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// cluster-wide barrier; every thread of every CTA of the cluster participates
__device__ __forceinline__ void ds_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire;" ::: "memory");
}
__device__ __forceinline__ void ds_fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ uint32_t ds_mapa(uint32_t addr, uint32_t cta) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(addr), "r"(cta));
  return remote;
}
__device__ __forceinline__ void ds_st_cluster_v4(uint32_t raddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared::cluster.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(raddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void ds_arrive_remote(uint32_t raddr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(raddr) : "memory");
}
__device__ __forceinline__ uint32_t ds_try_wait_cluster(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok;
}
// bounded like mbar_wait: a protocol bug traps instead of hanging the box
__device__ __forceinline__ void ds_wait_cluster(uint32_t bar, uint32_t parity) {
  if (ds_try_wait_cluster(bar, parity)) return;
  const uint64_t t0 = global_timer_ns();
  uint32_t spins = 0;
  while (!ds_try_wait_cluster(bar, parity)) {
    if ((++spins & 1023u) == 0 && global_timer_ns() - t0 > 2000000000ull) __trap();
  }
}
// arrive (once the MMAs issued so far by this thread have completed) on the barrier at this offset in every CTA of `mask`
__device__ __forceinline__ void ds_commit_multicast(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(mask)
               : "memory");
}

template <typename A1T>
__global__ void __launch_bounds__(DS_THREADS, 1)
deep_step_kernel(const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmW1,
                 const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmW3, const DeepArgs a) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * DS_MAX_STAGES + 4];
  __shared__ uint32_t s_tmem_base;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
  const int CS = a.CS, NS = a.NS;
  const int rank = (int)ds_ctarank();
  const int tile = blockIdx.x / CS;
  const int row0 = tile * 128;
  const uint32_t abuf = (smem_u32(smem_raw) + 1023u) & ~1023u;           // A operand: F/64 boxes [128][128 B], swizzled
  uint8_t* base = smem_raw + (abuf - smem_u32(smem_raw));
  const uint32_t ring = abuf + a.abuf_bytes;                             // weight ring
  float* pm_s = reinterpret_cast<float*>(base);                          // pm rows of this CTA's images (alias the A buffer)
  float* ep_s = reinterpret_cast<float*>(base + a.abuf_bytes + a.stages * a.stage_bytes);   // [2 nets][e, e*b][NS]
  float* body_s = ep_s + 4 * NS;
  const uint32_t bar_full = smem_u32(&bars[0]), bar_empty = smem_u32(&bars[DS_MAX_STAGES]);
  const uint32_t bar_afull = smem_u32(&bars[2 * DS_MAX_STAGES]), xbar = bar_afull + 8, ybar = bar_afull + 16,
                 pbar = bar_afull + 24;

  auto stamp = [&](int slot) {
    if (a.dbg != nullptr && tid == 64)
      a.dbg[(int64_t)blockIdx.x * 16 + slot] = (slot == 0 || slot == 15) ? (long long)global_timer_ns() : clock64();
  };
  stamp(0);
  stamp(1);
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA1); tma_prefetch_desc(&tmW1); tma_prefetch_desc(&tmW2); tma_prefetch_desc(&tmW3);
    for (int s = 0; s < a.stages; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    mbar_init(bar_afull, 1);
    mbar_init(xbar, CS);
    mbar_init(ybar, CS * DS_EPI_WARPS);
    mbar_init(pbar, CS * DS_EPI_WARPS);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(&s_tmem_base), 256);
  pdl_trigger();
  pdl_wait();
  stamp(2);
  for (int i = tid; i < NS; i += DS_THREADS) {          // y = max(0, e*acc + e*b)
    const int n = rank * NS + i;
    const float e1 = expf(a.s1[n]), e2 = expf(a.s2[n]);
    ep_s[i] = e1;
    ep_s[NS + i] = e1 * a.b1[n];
    ep_s[2 * NS + i] = e2;
    ep_s[3 * NS + i] = e2 * a.b2[n];
  }
  tc_fence_before();
  ds_cluster_sync();                  // barrier inits are visible to the peers before anybody arrives remotely
  tc_fence_after();
  const uint32_t tmem_base = s_tmem_base;
  stamp(3);

  const int nkb0 = a.K1p / DS_BK, nkb = a.F / DS_BK;
  const bool p2_active = rank * a.n3s < a.ldp;          // else: the GEMM3 slice lies entirely inside the zero padding

  if (warp == 0) {
    // ===================== TMA producer: A1 once, then the weight slices of all three GEMMs back to back ================
    if (lane == 0) {
      mbar_arrive_expect_tx(bar_afull, (uint32_t)nkb0 * 16384u);
      for (int kb = 0; kb < nkb0; ++kb) tma_load_2d(abuf + kb * 16384, &tmA1, kb * DS_BK, row0, bar_afull);
    }
    int stage = 0;
    uint32_t phase = 0;
    for (int ph = 0; ph < 3; ++ph) {
      if (ph == 2 && !p2_active) break;
      const CUtensorMap* mB = (ph == 0) ? &tmW1 : (ph == 1) ? &tmW2 : &tmW3;
      const int bn = (ph == 2) ? a.n3s : NS;
      const int n_k = (ph == 0) ? nkb0 : nkb;
      for (int kb = 0; kb < n_k; ++kb) {
        mbar_wait(bar_empty + 8 * stage, phase ^ 1);
        if (lane == 0) {
          mbar_arrive_expect_tx(bar_full + 8 * stage, (uint32_t)(bn * 128));
          tma_load_2d(ring + stage * a.stage_bytes, mB, kb * DS_BK, rank * bn, bar_full + 8 * stage);
        }
        __syncwarp();
        if (++stage == a.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    int stage = 0;
    uint32_t phase = 0;
    const uint16_t all = (uint16_t)((1u << CS) - 1u);
    for (int ph = 0; ph < 3; ++ph) {
      if (ph == 0) mbar_wait(bar_afull, 0);
      else ds_wait_cluster(ybar, (uint32_t)((ph - 1) & 1));   // every slice of the new A operand has landed here
      ds_fence_proxy_async();                                   // peers' generic-proxy stores -> tensor-core reads
      tc_fence_after();
      const bool active = (ph < 2) || p2_active;
      if (active) {
        const int bn = (ph == 2) ? a.n3s : NS;
        const int n_k = (ph == 0) ? nkb0 : nkb;
        const uint32_t idesc = make_idesc(128, bn);
        for (int kb = 0; kb < n_k; ++kb) {
          mbar_wait(bar_full + 8 * stage, phase);
          tc_fence_after();
          if (lane == 0) {
            const uint64_t adesc = make_smem_desc(abuf + kb * 16384), bdesc = make_smem_desc(ring + stage * a.stage_bytes);
#pragma unroll
            for (int k = 0; k < DS_BK / 16; ++k) umma_bf16(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
            umma_commit(bar_empty + 8 * stage);
          }
          __syncwarp();
          if (++stage == a.stages) { stage = 0; phase ^= 1; }
        }
      }
      if (lane == 0) ds_commit_multicast(xbar, all);           // "this CTA no longer reads its A buffer" -> all CTAs
      __syncwarp();
    }
  } else {
    // ===================== epilogue warps: TMEM -> registers -> the peers' shared memory =====================
    const int q = warp & 3, sub = (warp - 2) >> 2;              // TMEM lane quadrant, column-chunk owner (4 per quadrant)
    const int trow = q * 32 + lane;
    const int64_t grow = (int64_t)row0 + trow;
    const bool row_ok = grow < a.M;
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    for (int ph = 0; ph < 3; ++ph) {
      ds_wait_cluster(xbar, (uint32_t)(ph & 1));                // own accumulator complete, every A buffer free
      tc_fence_after();
      if (ph == 0) stamp(4); else if (ph == 1) stamp(7); else stamp(10);
      const bool active = (ph < 2) || p2_active;
      const int bn = (ph == 2) ? a.n3s : NS;
      const int col0 = rank * bn;
      if (active) {
        for (int c0 = sub * 16; c0 < bn; c0 += 16 * (DS_EPI_WARPS / 4)) {
          uint32_t r[16];
          tmem_ld16(taddr + c0, r);
          tmem_ld_wait();
          const int n = col0 + c0;                               // first column of the chunk in the full row
          if (ph < 2) {
            const float* e = ep_s + ph * 2 * NS + c0;
            uint32_t w[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float v0 = fmaxf(0.f, fmaf(e[2 * i], __uint_as_float(r[2 * i]), e[NS + 2 * i]));
              const float v1 = fmaxf(0.f, fmaf(e[2 * i + 1], __uint_as_float(r[2 * i + 1]), e[NS + 2 * i + 1]));
              __nv_bfloat162 t = __floats2bfloat162_rn(v0, v1);
              w[i] = *reinterpret_cast<uint32_t*>(&t);
            }
            // K-major SWIZZLE_128B box layout: 16-byte unit u of row r at r*128 + ((u ^ (r & 7)) << 4)
            const uint32_t box = abuf + (uint32_t)(n >> 6) * 16384u + (uint32_t)trow * 128u;
            const int u0 = (n & 63) >> 3;
            const uint32_t o0 = box + (uint32_t)(((u0 + 0) ^ (trow & 7)) << 4), o1 = box + (uint32_t)(((u0 + 1) ^ (trow & 7)) << 4);
            for (int c = 0; c < CS; ++c) {
              ds_st_cluster_v4(ds_mapa(o0, c), w[0], w[1], w[2], w[3]);
              ds_st_cluster_v4(ds_mapa(o1, c), w[4], w[5], w[6], w[7]);
            }
            __nv_bfloat16* hg = (ph == 0) ? a.h1 : a.h2;
            if (hg != nullptr && row_ok) {
              __nv_bfloat16* dst = hg + grow * a.F + n;
              *reinterpret_cast<uint4*>(dst) = make_uint4(w[0], w[1], w[2], w[3]);
              *reinterpret_cast<uint4*>(dst + 8) = make_uint4(w[4], w[5], w[6], w[7]);
            }
          } else {
            const int rpc = 128 / CS;                            // rows (= ipc whole images) per CTA
            const uint32_t dst = ds_mapa(abuf + (uint32_t)(((trow % rpc) * a.lds + n) * 4), (uint32_t)(trow / rpc));
#pragma unroll
            for (int j = 0; j < 4; ++j) ds_st_cluster_v4(dst + 16 * j, r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
            if (a.pm != nullptr && row_ok && n + 16 <= a.ld_pm) {
              float* pg = a.pm + grow * a.ld_pm + n;
#pragma unroll
              for (int j = 0; j < 4; ++j)
                *reinterpret_cast<uint4*>(pg + 4 * j) = make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
            }
          }
        }
      }
      if (ph == 0) stamp(5); else if (ph == 1) stamp(8); else stamp(11);
      tc_fence_before();
      ds_fence_proxy_async();
      __syncwarp();
      if (lane < CS) ds_arrive_remote(ds_mapa(ph == 2 ? pbar : ybar, (uint32_t)lane));
      __syncwarp();
      if (ph == 0) stamp(6); else if (ph == 1) stamp(9); else stamp(12);
    }
  }

  // ===================== phase 3: step boundary of this CTA's image(s), pm rows in shared memory =====================
  ds_wait_cluster(pbar, 0);
  stamp(13);
  for (int i = 0; i < a.ipc; ++i) {
    const int b = (tile * CS + rank) * a.ipc + i;
    if (b < a.bd.B)
      flow_boundary_body<true, A1T, true>(a.bd, b, body_s, tid, DS_THREADS, pm_s + (size_t)i * a.bd.H * a.bd.W * a.lds, a.lds);
    __syncthreads();
  }
  stamp(14);
  tc_fence_before();
  ds_cluster_sync();                  // nobody exits while a peer may still address its shared memory
  stamp(15);
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

// shape plan shared by the launcher and the support query
static bool ds_plan(int B, int C, int H, int W, int F, int64_t K1p, int64_t ldp, bool mix, DeepArgs* a, size_t* smem) {
  const int P = H * W;
  if (B <= 0 || C <= 0 || C % 2 || P <= 0 || P > 64 || 128 % P) return false;
  int CS = std::min(8, 128 / P);                             // portable cluster size
  if (const char* e = getenv("NFDPM_DEEP_CS")) {
    const int v = atoi(e);
    if ((v == 2 || v == 4 || v == 8) && v <= 128 / P) CS = v;
  }
  if (CS < 2) return false;
  const int ipc = 128 / (P * CS);
  if (F % 64 || F > 512 || F % (CS * 16) || K1p % DS_BK || K1p < 64 || K1p > F || ldp % 16 || ldp < 9 * C) return false;
  const int NS = F / CS;
  if (NS > 256) return false;
  const int n3s = (int)(((ldp + CS - 1) / CS + 15) / 16 * 16);
  if (n3s > 256) return false;
  const int stage_bytes = std::max(NS, n3s) * 128;
  const int stages = std::min(DS_MAX_STAGES, DS_RING_BYTES / stage_bytes);
  if (stages < 2) return false;
  const int abuf = (F / 64) * 16384;
  const int lds = CS * n3s + 4;
  if ((size_t)(128 / CS) * lds * 4 > (size_t)abuf) return false;
  const size_t body = nfdpm_flow_boundary_smem(C, H, W, 1, mix ? 1 : 0);
  const size_t total = 1024 + (size_t)abuf + (size_t)stages * stage_bytes + (size_t)4 * NS * sizeof(float) + body;
  if (total > (size_t)DS_SMEM_LIMIT) return false;
  if (a) {
    a->CS = CS; a->NS = NS; a->n3s = n3s; a->ipc = ipc; a->stages = stages; a->stage_bytes = stage_bytes;
    a->abuf_bytes = abuf; a->lds = lds;
  }
  if (smem) *smem = total;
  return true;
}

static long long* g_deep_dbg = nullptr;

}  // namespace nfdpm

using namespace nfdpm;

// profiling hook: per-CTA timeline buffer [grid][16] int64 (device memory), NULL switches it off
extern "C" int nfdpm_deep_step_debug(void* buf) {
  g_deep_dbg = reinterpret_cast<long long*>(buf);
  return 0;
}

extern "C" int nfdpm_deep_step_ok(int B, int C, int H, int W, int F, int64_t K1p, int64_t ldp) {
  return ds_plan(B, C, H, W, F, K1p, ldp, true, nullptr, nullptr) ? 1 : 0;
}

extern "C" int nfdpm_deep_step(const void* a1_in, const void* w1p, const void* w2p, const void* w3p, const float* s1,
                               const float* b1, const float* s2, const float* b2, void* h1, void* h2, float* pm,
                               int64_t ld_pm, const float* in, int64_t in_bs, const float* bias3, const float* logs3,
                               float* ld_part, const float* mt, const float* beta, float* y, int64_t y_bs, float* xs,
                               int64_t xs_bs, void* a1, int a1_dtype, int64_t lda1, int B, int C, int H, int W, int F,
                               int64_t K1p, int64_t ldp, int inverse, nfdpm_stream_t stream) {
  NFDPM_REQUIRE(a1_in && w1p && w2p && w3p && s1 && b1 && s2 && b2 && in && bias3 && logs3, "nfdpm_deep_step: null pointer");
  NFDPM_REQUIRE((mt == nullptr) == (beta == nullptr), "nfdpm_deep_step: mt/beta must both be set or both NULL");
  NFDPM_REQUIRE(y != nullptr || a1 != nullptr || xs != nullptr, "nfdpm_deep_step: no sink");
  NFDPM_REQUIRE(a1 == nullptr || (lda1 % 8 == 0 && lda1 >= 9 * (int64_t)(C / 2) && ((uintptr_t)a1 % 16) == 0),
                "nfdpm_deep_step: bad im2col sink");
  NFDPM_REQUIRE(a1 == nullptr || a1_dtype == NFDPM_F32 || a1_dtype == NFDPM_BF16, "nfdpm_deep_step: bad a1 dtype");
  NFDPM_REQUIRE((((uintptr_t)a1_in | (uintptr_t)w1p | (uintptr_t)w2p | (uintptr_t)w3p | (uintptr_t)h1 | (uintptr_t)h2 |
                  (uintptr_t)pm) % 16) == 0, "nfdpm_deep_step: operands must be 16-byte aligned");
  NFDPM_REQUIRE(pm == nullptr || (ld_pm >= ldp && ld_pm % 16 == 0), "nfdpm_deep_step: bad pm copy (ld_pm %lld)", (long long)ld_pm);
  DeepArgs a;
  size_t smem = 0;
  NFDPM_REQUIRE(ds_plan(B, C, H, W, F, K1p, ldp, mt != nullptr, &a, &smem),
                "nfdpm_deep_step: unsupported shape B=%d C=%d H=%d W=%d F=%d K1p=%lld ldp=%lld (use nfdpm_gemm_nt + "
                "nfdpm_flow_boundary)", B, C, H, W, F, (long long)K1p, (long long)ldp);
  const int64_t M = (int64_t)B * H * W;
  a.M = (int)M; a.K1p = (int)K1p; a.F = F; a.ldp = (int)ldp;
  a.s1 = s1; a.b1 = b1; a.s2 = s2; a.b2 = b2;
  a.h1 = reinterpret_cast<__nv_bfloat16*>(h1); a.h2 = reinterpret_cast<__nv_bfloat16*>(h2);
  a.pm = pm; a.ld_pm = ld_pm;
  a.dbg = g_deep_dbg;
  BoundaryArgs& d = a.bd;
  d.in = in; d.in_bs = in_bs; d.pm = nullptr; d.ldp = 0; d.bias3 = bias3; d.logs3 = logs3; d.ld_part = ld_part;
  d.mt = mt; d.beta = beta; d.y = y; d.y_bs = y_bs; d.xs = xs; d.xs_bs = xs_bs; d.a1 = a1; d.lda1 = lda1;
  d.B = B; d.C = C; d.H = H; d.W = W; d.squeeze_in = 0; d.inverse = inverse;
  boundary_fill_div(d);
  CUtensorMap tmA1, tmW1, tmW2, tmW3;
  if (make_map(&tmA1, a1_in, M, K1p, K1p, 128)) return 1;
  if (make_map(&tmW1, w1p, F, K1p, K1p, a.NS)) return 1;
  if (make_map(&tmW2, w2p, F, F, F, a.NS)) return 1;
  if (make_map(&tmW3, w3p, ldp, F, F, a.n3s)) return 1;
  const int tiles = (int)((M + 127) / 128);
  const bool bf = (a1 != nullptr && a1_dtype == NFDPM_BF16);
  static bool attr_set = false;
  if (!attr_set) {
    NFDPM_CUDA(cudaFuncSetAttribute(deep_step_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, DS_SMEM_LIMIT));
    NFDPM_CUDA(cudaFuncSetAttribute(deep_step_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, DS_SMEM_LIMIT));
    attr_set = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(tiles * a.CS);
  cfg.blockDim = dim3(DS_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = as_stream(stream);
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = a.CS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  if (bf) NFDPM_CUDA(cudaLaunchKernelEx(&cfg, deep_step_kernel<__nv_bfloat16>, tmA1, tmW1, tmW2, tmW3, a));
  else NFDPM_CUDA(cudaLaunchKernelEx(&cfg, deep_step_kernel<float>, tmA1, tmW1, tmW2, tmW3, a));
  return 0;
}
