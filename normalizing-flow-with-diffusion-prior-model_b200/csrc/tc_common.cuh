// tcgen05 / TMEM / TMA / mbarrier PTX wrappers shared by the tensor-core kernels (sm_100a).  Internal header.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace nfdpm {

constexpr int TC_EPI_WARPS = 8;      // epilogue warps: two per TMEM lane quadrant, alternating 16-column chunks

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok;
}
__device__ __forceinline__ uint64_t global_timer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// Bounded wait: a protocol bug must abort the kernel (trap -> launch failure), never hang the box.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const uint64_t t0 = global_timer_ns();
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 1023u) == 0 && global_timer_ns() - t0 > 2000000000ull) __trap();
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(map), "r"(src), "r"(c0), "r"(c1)
               : "memory");
}
// 1-D bulk copy global -> shared memory (size a multiple of 16 bytes, both addresses 16-byte aligned), completion on an mbarrier
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void epi_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(32 * TC_EPI_WARPS) : "memory"); }
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, single CTA, bf16 inputs / fp32 accumulate
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier once every previously issued tcgen05.mma of this thread has completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory descriptor, K-major operand, 128-byte swizzle, dense [rows][64] bf16 tile:
//   start address >> 4 | LBO (unused for swizzled K-major, 1) | SBO = 1024 B (8 rows x 128 B) >> 4 = 64 |
//   descriptor version 1 (Blackwell) | layout type 2 = SWIZZLE_128B
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// Instruction descriptor (kind::f16): D=F32 (bit 4), A=BF16 (bits 7..9 = 1), B=BF16 (bits 10..12 = 1),
// both K-major (bits 15,16 = 0), N>>3 at bits 17..22, M>>4 at bits 24..28
__device__ __forceinline__ uint32_t make_idesc(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// Staging-tile writers.  The tile is a sequence of 16 KB boxes [128 rows][128 bytes] in the TMA 128B-swizzle layout:
// the 16-byte unit u of row r lives at r*128 + ((u ^ (r & 7)) * 16).  One thread owns one row, so the 8 rows of a
// quarter-warp hit 8 different units: conflict-free st.shared.v4.
template <typename OutT> struct TcStage;
template <> struct TcStage<__nv_bfloat16> {
  static constexpr int kColsPerBox = 64;
  static __device__ __forceinline__ void put16(uint32_t cbase, int row, int c0, const float (&v)[16]) {
    uint32_t w[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      __nv_bfloat162 t = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t*>(&t);
    }
    const uint32_t box = cbase + (uint32_t)(c0 >> 6) * 16384u + (uint32_t)row * 128u;
    const int u0 = (c0 & 63) >> 3;
    st_shared_v4(box + (uint32_t)(((u0 + 0) ^ (row & 7)) << 4), w[0], w[1], w[2], w[3]);
    st_shared_v4(box + (uint32_t)(((u0 + 1) ^ (row & 7)) << 4), w[4], w[5], w[6], w[7]);
  }
};
template <> struct TcStage<float> {
  static constexpr int kColsPerBox = 32;
  static __device__ __forceinline__ void put16(uint32_t cbase, int row, int c0, const float (&v)[16]) {
    const uint32_t box = cbase + (uint32_t)(c0 >> 5) * 16384u + (uint32_t)row * 128u;
    const int u0 = (c0 & 31) >> 2;
#pragma unroll
    for (int i = 0; i < 4; ++i)
      st_shared_v4(box + (uint32_t)(((u0 + i) ^ (row & 7)) << 4), __float_as_uint(v[4 * i]), __float_as_uint(v[4 * i + 1]),
                   __float_as_uint(v[4 * i + 2]), __float_as_uint(v[4 * i + 3]));
  }
};

// split bf16 pairs (common.cuh): a 16 KB box [128 rows][128 bytes] holds 32 LOGICAL columns, hi in the units 0..3 and lo in
// the units 4..7 of a row
template <> struct TcStage<bf16x2_t> {
  static constexpr int kColsPerBox = 32;
  static __device__ __forceinline__ void put16(uint32_t cbase, int row, int c0, const float (&v)[16]) {
    uint32_t h[8], l[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      __nv_bfloat162 th = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
      const float2 back = __bfloat1622float2(th);
      __nv_bfloat162 tl = __floats2bfloat162_rn(v[2 * i] - back.x, v[2 * i + 1] - back.y);
      h[i] = *reinterpret_cast<uint32_t*>(&th);
      l[i] = *reinterpret_cast<uint32_t*>(&tl);
    }
    const uint32_t box = cbase + (uint32_t)(c0 >> 5) * 16384u + (uint32_t)row * 128u;
    const int u0 = (c0 & 31) >> 3;                        // 0 or 2
    st_shared_v4(box + (uint32_t)(((u0 + 0) ^ (row & 7)) << 4), h[0], h[1], h[2], h[3]);
    st_shared_v4(box + (uint32_t)(((u0 + 1) ^ (row & 7)) << 4), h[4], h[5], h[6], h[7]);
    st_shared_v4(box + (uint32_t)(((u0 + 4) ^ (row & 7)) << 4), l[0], l[1], l[2], l[3]);
    st_shared_v4(box + (uint32_t)(((u0 + 5) ^ (row & 7)) << 4), l[4], l[5], l[6], l[7]);
  }
};
// bf16 columns of the global row per logical column of the tile (TMA store coordinates)
template <typename OutT> struct TcMemCols { static constexpr int v = 1; };
template <> struct TcMemCols<bf16x2_t> { static constexpr int v = 2; };

// The MMAs of one 64-bf16-column operand stage.  Plain bf16: four K=16 slices.  Split pairs: the stage is 32 logical K,
// slices {hi0, hi1, lo0, lo1}; D += Ahi*Bhi + Alo*Bhi + Ahi*Blo.  `first` = this stage starts the accumulation.
template <bool X3>
__device__ __forceinline__ void umma_stage(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, bool first) {
  if (X3) {
    umma_bf16(tmem_d, adesc + 0, bdesc + 0, idesc, first ? 0u : 1u);
    umma_bf16(tmem_d, adesc + 2, bdesc + 2, idesc, 1u);
    umma_bf16(tmem_d, adesc + 4, bdesc + 0, idesc, 1u);
    umma_bf16(tmem_d, adesc + 6, bdesc + 2, idesc, 1u);
    umma_bf16(tmem_d, adesc + 0, bdesc + 4, idesc, 1u);
    umma_bf16(tmem_d, adesc + 2, bdesc + 6, idesc, 1u);
  } else {
#pragma unroll
    for (int k = 0; k < 4; ++k)                           // +32 bytes (= 2 x 16 B) per K=16 slice inside the swizzle row
      umma_bf16(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (first && k == 0) ? 0u : 1u);
  }
}


// ---------------------------------------------------------------- host side: tensor maps
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static inline EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 2-D row-major [rows, cols] with leading dimension ld (elements); box = [box_rows, 128 bytes of columns], 128B swizzle
static inline int make_map(CUtensorMap* map, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows,
                           bool f32 = false) {
  EncodeTiledFn enc = encode_fn();
  NFDPM_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled entry point not available");
  const int esz = f32 ? 4 : 2;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * esz};
  cuuint32_t box[2] = {(cuuint32_t)(128 / esz), (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                   const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  NFDPM_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d (rows=%lld cols=%lld ld=%lld box=%d)",
                (int)r, (long long)rows, (long long)cols, (long long)ld, box_rows);
  return 0;
}

}  // namespace nfdpm
