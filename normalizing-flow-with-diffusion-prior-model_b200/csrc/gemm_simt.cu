// Exact-fp32 CUDA-core GEMM  D[M,N] = epilogue(A[M,K] * Bw[N,K]^T)  for the coupling networks.
// This is the "fp32" precision mode of the library (bit-faithful fp32 FMA accumulation, used for the
// 1e-4 parity bar, for data-dependent initialisation, and as the on-device cross-check of the tcgen05
// path); the bf16 tensor-core kernel lives in gemm_tc.cu.  nfdpm_gemm_nt dispatches on in_dtype.
//
// Tile 128 x BN x 16, 256 threads, 8 x (BN/16) outputs per thread (split 4+4 so shared-memory reads are
// conflict-free 128-bit), register-prefetched double buffering.
#include <stdlib.h>

#include "common.cuh"

namespace nfdpm {

int gemm_nt_tc(const void* A, int64_t lda, const void* Bw, int64_t ldb, void* D, int64_t ldd, int M, int N, int K,
               int in_dtype, int out_dtype, int epilogue, const float* ep_scale, const float* ep_bias, cudaStream_t st);

constexpr int BK = 16;

template <typename OutT> struct Store4;
template <> struct Store4<float> {
  static __device__ __forceinline__ void vec(float* p, float a, float b, float c, float d) {
    *reinterpret_cast<float4*>(p) = make_float4(a, b, c, d);
  }
  static __device__ __forceinline__ void one(float* p, float a) { *p = a; }
};
template <> struct Store4<__nv_bfloat16> {
  static __device__ __forceinline__ void vec(__nv_bfloat16* p, float a, float b, float c, float d) {
    __nv_bfloat162 lo = __floats2bfloat162_rn(a, b), hi = __floats2bfloat162_rn(c, d);
    uint2 u;
    u.x = *reinterpret_cast<uint32_t*>(&lo);
    u.y = *reinterpret_cast<uint32_t*>(&hi);
    *reinterpret_cast<uint2*>(p) = u;
  }
  static __device__ __forceinline__ void one(__nv_bfloat16* p, float a) { *p = __float2bfloat16_rn(a); }
};

template <int BN, int EPI, typename OutT>
__global__ void __launch_bounds__(256) gemm_nt_f32_kernel(const float* __restrict__ A, int64_t lda,
                                                          const float* __restrict__ Bw, int64_t ldb,
                                                          OutT* __restrict__ D, int64_t ldd, int M, int N, int K,
                                                          const float* __restrict__ ep_scale,
                                                          const float* __restrict__ ep_bias) {
  constexpr int BM = 128;
  constexpr int TN = BN / 16;        // 8 or 4 outputs per thread along N
  constexpr int NG = TN / 4;         // column groups of 4 (2 or 1)
  constexpr int B_LD4 = BN * BK / 4 / 256;  // float4 loads of the B tile per thread (2 or 1)
  __shared__ __align__(16) float As[2][BK][BM + 4];
  __shared__ __align__(16) float Bs[2][BK][BN + 4];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;

  float acc[8][TN];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  float4 ra[2], rb[B_LD4];
  auto gload = [&](int k0) {
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int idx = tid + r * 256, row = idx >> 2, kq = idx & 3;
      ra[r] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (m0 + row < M) ra[r] = *reinterpret_cast<const float4*>(A + (int64_t)(m0 + row) * lda + k0 + kq * 4);
    }
#pragma unroll
    for (int r = 0; r < B_LD4; ++r) {
      const int idx = tid + r * 256, row = idx >> 2, kq = idx & 3;
      rb[r] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (n0 + row < N) rb[r] = __ldg(reinterpret_cast<const float4*>(Bw + (int64_t)(n0 + row) * ldb + k0 + kq * 4));
    }
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int idx = tid + r * 256, row = idx >> 2, kq = idx & 3;
      As[buf][kq * 4 + 0][row] = ra[r].x;
      As[buf][kq * 4 + 1][row] = ra[r].y;
      As[buf][kq * 4 + 2][row] = ra[r].z;
      As[buf][kq * 4 + 3][row] = ra[r].w;
    }
#pragma unroll
    for (int r = 0; r < B_LD4; ++r) {
      const int idx = tid + r * 256, row = idx >> 2, kq = idx & 3;
      Bs[buf][kq * 4 + 0][row] = rb[r].x;
      Bs[buf][kq * 4 + 1][row] = rb[r].y;
      Bs[buf][kq * 4 + 2][row] = rb[r].z;
      Bs[buf][kq * 4 + 3][row] = rb[r].w;
    }
  };

  const int nk = K / BK;
  gload(0);
  sstore(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) gload((kt + 1) * BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float b[TN];
#pragma unroll
      for (int g = 0; g < NG; ++g) {
        const float4 bv = *reinterpret_cast<const float4*>(&Bs[buf][k][g * 64 + tx * 4]);
        b[g * 4 + 0] = bv.x; b[g * 4 + 1] = bv.y; b[g * 4 + 2] = bv.z; b[g * 4 + 3] = bv.w;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      sstore(buf ^ 1);
      __syncthreads();
    }
  }

  // ---- epilogue
  const bool vec_ok = (ldd % 4 == 0) && ((uintptr_t)D % 16 == 0);
#pragma unroll
  for (int g = 0; g < NG; ++g) {
    const int n = n0 + g * 64 + tx * 4;
    float es[4] = {1.f, 1.f, 1.f, 1.f}, eb[4] = {0.f, 0.f, 0.f, 0.f};
    if (EPI == NFDPM_EPI_ACTNORM_RELU) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (n + j < N) { es[j] = expf(__ldg(ep_scale + n + j)); eb[j] = __ldg(ep_bias + n + j); }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int m = m0 + (i >> 2) * 64 + ty * 4 + (i & 3);
      if (m >= M) continue;
      float v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        v[j] = acc[i][g * 4 + j];
        if (EPI == NFDPM_EPI_ACTNORM_RELU) v[j] = fmaxf(0.f, es[j] * (v[j] + eb[j]));
      }
      OutT* dp = D + (int64_t)m * ldd + n;
      if (vec_ok && n + 3 < N) {
        Store4<OutT>::vec(dp, v[0], v[1], v[2], v[3]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (n + j < N) Store4<OutT>::one(dp + j, v[j]);
      }
    }
  }
}

template <int BN, int EPI, typename OutT>
static int launch_simt(const float* A, int64_t lda, const float* Bw, int64_t ldb, void* D, int64_t ldd, int M, int N,
                       int K, const float* es, const float* eb, cudaStream_t st) {
  dim3 grid((N + BN - 1) / BN, (M + 127) / 128);
  gemm_nt_f32_kernel<BN, EPI, OutT><<<grid, 256, 0, st>>>(A, lda, Bw, ldb, (OutT*)D, ldd, M, N, K, es, eb);
  NFDPM_CHECK_LAUNCH("gemm_nt_f32_kernel");
  return 0;
}

}  // namespace nfdpm

using namespace nfdpm;

extern "C" int nfdpm_gemm_nt(const void* A, int64_t lda, const void* Bw, int64_t ldb, void* D, int64_t ldd, int M,
                             int N, int K, int in_dtype, int out_dtype, int epilogue, const float* ep_scale,
                             const float* ep_bias, nfdpm_stream_t stream) {
  NFDPM_REQUIRE(A && Bw && D, "nfdpm_gemm_nt: null pointer");
  NFDPM_REQUIRE(M > 0 && N > 0 && K > 0, "nfdpm_gemm_nt: bad shape M=%d N=%d K=%d", M, N, K);
  NFDPM_REQUIRE(lda >= K && ldb >= K && ldd >= N, "nfdpm_gemm_nt: leading dimension too small");
  NFDPM_REQUIRE(lda % 8 == 0 && ldb % 8 == 0 && ldd % 8 == 0, "nfdpm_gemm_nt: lda/ldb/ldd must be multiples of 8");
  NFDPM_REQUIRE(epilogue == NFDPM_EPI_RAW || epilogue == NFDPM_EPI_ACTNORM_RELU, "nfdpm_gemm_nt: bad epilogue %d", epilogue);
  NFDPM_REQUIRE(epilogue == NFDPM_EPI_RAW || (ep_scale && ep_bias), "nfdpm_gemm_nt: epilogue needs scale/bias");
  NFDPM_REQUIRE(out_dtype == NFDPM_F32 || out_dtype == NFDPM_BF16 || out_dtype == NFDPM_BF16X2, "nfdpm_gemm_nt: bad out_dtype %d", out_dtype);
  cudaStream_t st = as_stream(stream);
  if (in_dtype == NFDPM_BF16 || in_dtype == NFDPM_BF16X2) {
    if (in_dtype == NFDPM_BF16) NFDPM_REQUIRE(K % 64 == 0, "nfdpm_gemm_nt: bf16 path needs K %% 64 == 0 (K=%d)", K);
    else NFDPM_REQUIRE(K % 32 == 0 && lda % 32 == 0 && ldb % 32 == 0, "nfdpm_gemm_nt: split-pair path needs K, lda, ldb %% 32 == 0 (K=%d)", K);
    return gemm_nt_tc(A, lda, Bw, ldb, D, ldd, M, N, K, in_dtype, out_dtype, epilogue, ep_scale, ep_bias, st);
  }
  NFDPM_REQUIRE(out_dtype != NFDPM_BF16X2, "nfdpm_gemm_nt: the fp32 CUDA-core path writes fp32 or bf16");
  NFDPM_REQUIRE(in_dtype == NFDPM_F32, "nfdpm_gemm_nt: bad in_dtype %d", in_dtype);
  NFDPM_REQUIRE(K % 16 == 0, "nfdpm_gemm_nt: fp32 path needs K %% 16 == 0 (K=%d)", K);
  NFDPM_REQUIRE(((uintptr_t)A % 16 == 0) && ((uintptr_t)Bw % 16 == 0), "nfdpm_gemm_nt: operands must be 16-byte aligned");
  const float* a = (const float*)A;
  const float* b = (const float*)Bw;
  const bool wide = N > 64;
#define GO(BN, EPI, T) return launch_simt<BN, EPI, T>(a, lda, b, ldb, D, ldd, M, N, K, ep_scale, ep_bias, st)
  if (out_dtype == NFDPM_F32) {
    if (epilogue == NFDPM_EPI_RAW) { if (wide) GO(128, NFDPM_EPI_RAW, float); else GO(64, NFDPM_EPI_RAW, float); }
    else { if (wide) GO(128, NFDPM_EPI_ACTNORM_RELU, float); else GO(64, NFDPM_EPI_ACTNORM_RELU, float); }
  } else {
    if (epilogue == NFDPM_EPI_RAW) { if (wide) GO(128, NFDPM_EPI_RAW, __nv_bfloat16); else GO(64, NFDPM_EPI_RAW, __nv_bfloat16); }
    else { if (wide) GO(128, NFDPM_EPI_ACTNORM_RELU, __nv_bfloat16); else GO(64, NFDPM_EPI_ACTNORM_RELU, __nv_bfloat16); }
  }
#undef GO
}
