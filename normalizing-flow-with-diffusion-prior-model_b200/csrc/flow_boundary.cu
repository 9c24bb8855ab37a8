// Fused "step boundary" kernel: everything that happens to the flow state between the last GEMM of one StepFlow
// and the first GEMM of the next one, in ONE launch, one CTA per image:
//
//   source   SRC_PLAIN     x = in[b]                       (optionally with squeeze addressing, transforms.py:226)
//            SRC_COUPLING  x = affine coupling of in[b] with the taps-as-N ZeroConv rows pm (forward:
//                          transforms.py:179-184 incl. the per-image log-det partial; inverse: :196-200)
//   mix      u = Mt^T x + beta     fused ActNorm + invertible 1x1 conv of the NEXT step (forward, transforms.py:80,132)
//                                  or of THIS step (inverse, :144,:93); optional
//   sinks    y[b]  <- u            NCHW fp32 flow state (optional)
//            a1    <- im2col3x3(u[:C/2])  pixel-major rows for the next coupling network's first GEMM (optional)
//
// All global traffic is coalesced: the image is staged channel-major in shared memory, pm rows are read with the
// channel index fastest across lanes (each pm element is consumed exactly once), im2col rows are written as 16-byte
// (bf16) / 32-byte (fp32) pieces with the column group fastest across lanes.
// Replaces, per StepFlow: nfdpm_coupling_apply + nfdpm_channel_mix + nfdpm_im2col3x3 (+ nfdpm_squeeze at level entry).
#include "boundary_body.cuh"

namespace nfdpm {

template <bool COUPLING, typename A1T>
__global__ void __launch_bounds__(1024) flow_boundary_kernel(const BoundaryArgs a) {
  extern __shared__ __align__(16) float sm[];
  pdl_trigger();      // PDL: the next kernel of the chain may be scheduled as soon as SMs free up ...
  pdl_wait();         // ... and this one reads nothing before its predecessor has completed
  flow_boundary_body<COUPLING, A1T>(a, blockIdx.x, sm, threadIdx.x, blockDim.x);
}

}  // namespace nfdpm

using namespace nfdpm;

extern "C" size_t nfdpm_flow_boundary_smem(int C, int H, int W, int coupling, int mix) {
  const size_t P = (size_t)H * W, PS = P + 1, Cp = (C + 3) & ~3;
  size_t fl = (size_t)C * PS * (mix ? 2 : 1);
  if (mix) fl += (size_t)C * Cp + Cp;
  if (coupling) fl += 2 * (size_t)C + P * (C / 2);
  return fl * sizeof(float);
}

static int flow_boundary_impl(const float* in, int64_t in_bs, int squeeze_in, const float* pm, int64_t ldp,
                              const float* bias3, const float* logs3, float* ld_part, const float* mt,
                              const float* beta, float* y, int64_t y_bs, float* xs, int64_t xs_bs, void* a1, int a1_dtype,
                              int64_t lda1, int B, int C, int H, int W, int inverse, nfdpm_stream_t stream) {
  NFDPM_REQUIRE(in != nullptr, "nfdpm_flow_boundary: null input");
  NFDPM_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0 && C % 2 == 0, "nfdpm_flow_boundary: bad shape B=%d C=%d H=%d W=%d", B, C, H, W);
  NFDPM_REQUIRE((mt == nullptr) == (beta == nullptr), "nfdpm_flow_boundary: mt/beta must both be set or both NULL");
  NFDPM_REQUIRE(pm == nullptr || (bias3 && logs3 && ldp >= 9 * (int64_t)C), "nfdpm_flow_boundary: coupling source needs bias3/logs3/ldp");
  NFDPM_REQUIRE(!squeeze_in || (C % 4 == 0 && in_bs % 2 == 0 && ((uintptr_t)in % 8) == 0), "nfdpm_flow_boundary: squeeze source needs C %% 4 == 0 and 8-byte alignment");
  NFDPM_REQUIRE(a1 == nullptr || (lda1 % 8 == 0 && lda1 >= 9 * (int64_t)(C / 2) && ((uintptr_t)a1 % 16) == 0), "nfdpm_flow_boundary: bad im2col sink");
  NFDPM_REQUIRE(a1 == nullptr || a1_dtype == NFDPM_F32 || a1_dtype == NFDPM_BF16, "nfdpm_flow_boundary: bad a1 dtype");
  NFDPM_REQUIRE(y != nullptr || a1 != nullptr || xs != nullptr, "nfdpm_flow_boundary: no sink");
  const size_t smem = nfdpm_flow_boundary_smem(C, H, W, pm != nullptr, mt != nullptr);
  NFDPM_REQUIRE(smem <= 200 * 1024, "nfdpm_flow_boundary: image too large for the fused path (%zu bytes of shared memory); "
                "use the unfused kernels", smem);
  BoundaryArgs a;
  a.in = in; a.in_bs = in_bs; a.pm = pm; a.ldp = ldp; a.bias3 = bias3; a.logs3 = logs3; a.ld_part = ld_part;
  a.mt = mt; a.beta = beta; a.y = y; a.y_bs = y_bs; a.xs = xs; a.xs_bs = xs_bs; a.a1 = a1; a.lda1 = lda1;
  a.B = B; a.C = C; a.H = H; a.W = W; a.squeeze_in = squeeze_in; a.inverse = inverse;
  cudaStream_t st = as_stream(stream);
  const bool bf = (a1 != nullptr && a1_dtype == NFDPM_BF16);
  // one CTA per image: give it as many warps as its largest phase has work items (latency hiding), up to 1024 threads
  int64_t items = (int64_t)H * W * (C / 2);
  if (a1 != nullptr && (int64_t)H * W * (lda1 / 8) > items) items = (int64_t)H * W * (lda1 / 8);
  int threads = (int)((items + 31) / 32 * 32);
  if (threads > 1024) threads = 1024;
  if (threads < 128) threads = 128;
#define LAUNCH(CP, T)                                                                                              \
  do {                                                                                                             \
    static bool attr_set = false;                                                                                  \
    if (!attr_set) {                                                                                               \
      NFDPM_CUDA(cudaFuncSetAttribute(flow_boundary_kernel<CP, T>, cudaFuncAttributeMaxDynamicSharedMemorySize,    \
                                      200 * 1024));                                                                \
      attr_set = true;                                                                                             \
    }                                                                                                              \
    NFDPM_CUDA(launch_pdl(flow_boundary_kernel<CP, T>, dim3(B), dim3(threads), smem, st, a));                      \
  } while (0)
  if (pm != nullptr) { if (bf) LAUNCH(true, __nv_bfloat16); else LAUNCH(true, float); }
  else { if (bf) LAUNCH(false, __nv_bfloat16); else LAUNCH(false, float); }
#undef LAUNCH
  NFDPM_CHECK_LAUNCH("flow_boundary_kernel");
  return 0;
}

extern "C" int nfdpm_flow_boundary(const float* in, int64_t in_bs, int squeeze_in, const float* pm, int64_t ldp,
                                   const float* bias3, const float* logs3, float* ld_part, const float* mt,
                                   const float* beta, float* y, int64_t y_bs, void* a1, int a1_dtype, int64_t lda1,
                                   int B, int C, int H, int W, int inverse, nfdpm_stream_t stream) {
  return flow_boundary_impl(in, in_bs, squeeze_in, pm, ldp, bias3, logs3, ld_part, mt, beta, y, y_bs, nullptr, 0, a1,
                            a1_dtype, lda1, B, C, H, W, inverse, stream);
}

extern "C" int nfdpm_flow_boundary_stash(const float* in, int64_t in_bs, int squeeze_in, const float* pm, int64_t ldp,
                                         const float* bias3, const float* logs3, float* ld_part, const float* mt,
                                         const float* beta, float* y, int64_t y_bs, float* xs, int64_t xs_bs, void* a1,
                                         int a1_dtype, int64_t lda1, int B, int C, int H, int W, nfdpm_stream_t stream) {
  return flow_boundary_impl(in, in_bs, squeeze_in, pm, ldp, bias3, logs3, ld_part, mt, beta, y, y_bs, xs, xs_bs, a1,
                            a1_dtype, lda1, B, C, H, W, 0, stream);
}
