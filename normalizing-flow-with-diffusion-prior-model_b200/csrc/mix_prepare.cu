// K-LU + fold: batched fp64 partial-pivot LU of the invertible 1x1-conv weights, giving log|det W|
// (replaces torch.slogdet(W.double()), normalizing_flow/transforms.py:131), W^-1 (replaces
// W.inverse(), transforms.py:144) and the ActNorm-folded forward / inverse mixing matrices consumed by
// nfdpm_channel_mix.  One CTA per StepFlow, up to 64 StepFlows per launch (items passed by value, so the
// call is CUDA-graph capturable and needs no device-side pointer table).
#include "common.cuh"

namespace nfdpm {

constexpr int kPrepBatch = 64;     // 64 x 88-byte items by value: needs the >4 KB kernel-parameter space (CUDA 12.1+, sm_70+)
struct PrepBatch {
  nfdpm_mix_item it[kPrepBatch];
};

__global__ void __launch_bounds__(256) mix_prepare_kernel(const PrepBatch batch, int smem_doubles) {
  extern __shared__ __align__(16) double dsm[];
  __shared__ int s_piv;
  __shared__ double s_logdet;
  const nfdpm_mix_item it = batch.it[blockIdx.x];
  const int C = it.C;
  const int t = threadIdx.x;
  double* A = nullptr;    // LU factors, row-major [C][C]
  double* X = nullptr;    // inverse, row-major [C][C]
  if (it.weight != nullptr) {
    const bool in_smem = 2 * C * C <= smem_doubles;
    A = in_smem ? dsm : it.lu_ws;
    X = A + C * C;
    for (int i = t; i < C * C; i += 256) A[i] = (double)it.weight[i];
    if (t == 0) s_logdet = 0.0;
    __syncthreads();
    // perm[k] is kept implicitly by physically swapping rows of A and of the right-hand side (X starts as I)
    for (int i = t; i < C * C; i += 256) X[i] = ((i / C) == (i % C)) ? 1.0 : 0.0;
    __syncthreads();
    for (int k = 0; k < C; ++k) {
      if (t < 32) {  // pivot search by warp 0
        double best = -1.0;
        int bi = k;
        for (int r = k + t; r < C; r += 32) {
          const double v = fabs(A[r * C + k]);
          if (v > best) { best = v; bi = r; }
        }
        for (int o = 16; o > 0; o >>= 1) {
          const double ob = __shfl_xor_sync(0xffffffffu, best, o);
          const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
          if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
        }
        if (t == 0) s_piv = bi;
      }
      __syncthreads();
      const int piv = s_piv;
      if (piv != k) {
        for (int c = t; c < C; c += 256) {
          double a = A[k * C + c]; A[k * C + c] = A[piv * C + c]; A[piv * C + c] = a;
          double x = X[k * C + c]; X[k * C + c] = X[piv * C + c]; X[piv * C + c] = x;
        }
      }
      __syncthreads();
      const double d = A[k * C + k];
      if (t == 0) s_logdet += log(fabs(d));
      const double inv_d = 1.0 / d;
      __syncthreads();
      for (int r = k + 1 + t; r < C; r += 256) A[r * C + k] *= inv_d;
      __syncthreads();
      const int n = C - k - 1;
      for (int i = t; i < n * n; i += 256) {
        const int r = k + 1 + i / n, c = k + 1 + i % n;
        A[r * C + c] -= A[r * C + k] * A[k * C + c];
      }
      __syncthreads();
    }
    // Solve L U X = P I column by column: thread j owns column j of X.
    for (int j = t; j < C; j += 256) {
      for (int r = 1; r < C; ++r) {  // forward substitution (unit lower)
        double s = X[r * C + j];
        for (int c = 0; c < r; ++c) s -= A[r * C + c] * X[c * C + j];
        X[r * C + j] = s;
      }
      for (int r = C - 1; r >= 0; --r) {  // back substitution
        double s = X[r * C + j];
        for (int c = r + 1; c < C; ++c) s -= A[r * C + c] * X[c * C + j];
        X[r * C + j] = s / A[r * C + r];
      }
    }
    __syncthreads();
  }
  // ---- fold ActNorm into the mixing matrices
  auto Wv = [&](int o, int i) -> float { return it.weight ? it.weight[o * C + i] : (o == i ? 1.f : 0.f); };
  auto Wi = [&](int o, int i) -> float { return it.weight ? (float)X[o * C + i] : (o == i ? 1.f : 0.f); };
  auto Sv = [&](int i) -> float { return it.scale ? it.scale[i] : 0.f; };
  auto Bv = [&](int i) -> float { return it.bias ? it.bias[i] : 0.f; };
  for (int idx = t; idx < C * C; idx += 256) {
    const int i = idx / C, o = idx - i * C;
    if (it.fwd_mt) it.fwd_mt[idx] = Wv(o, i) * expf(Sv(i));
    if (it.inv_mt) it.inv_mt[idx] = expf(-Sv(o)) * Wi(o, i);
    if (it.winv) it.winv[idx] = Wi(i, o);  // idx = row*C + col with row=i, col=o
  }
  for (int o = t; o < C; o += 256) {
    if (it.fwd_beta) {
      float s = 0.f;
      for (int i = 0; i < C; ++i) s = fmaf(Wv(o, i) * expf(Sv(i)), Bv(i), s);
      it.fwd_beta[o] = s;
    }
    if (it.inv_beta) it.inv_beta[o] = -Bv(o);
  }
  if (t == 0 && it.logdet) {
    double s = it.weight ? s_logdet : 0.0;
    float ss = 0.f;  // reference sums the fp32 scales in fp32 (transforms.py:81)
    for (int i = 0; i < C; ++i) ss += Sv(i);
    it.logdet[0] = (float)s + ss;
  }
}

}  // namespace nfdpm

using namespace nfdpm;

extern "C" int nfdpm_mix_prepare(const nfdpm_mix_item* items, int n, nfdpm_stream_t stream) {
  NFDPM_REQUIRE(items != nullptr && n > 0, "nfdpm_mix_prepare: no items");
  int maxC = 0;
  for (int i = 0; i < n; ++i) {
    NFDPM_REQUIRE(items[i].C > 0 && items[i].C <= 1024, "nfdpm_mix_prepare: item %d has C=%d", i, items[i].C);
    NFDPM_REQUIRE(items[i].weight == nullptr || items[i].lu_ws != nullptr, "nfdpm_mix_prepare: item %d needs lu_ws", i);
    if (items[i].C > maxC) maxC = items[i].C;
  }
  int smem_doubles = 2 * maxC * maxC;
  if ((size_t)smem_doubles * sizeof(double) > 96 * 1024) smem_doubles = 0;  // large C: work in lu_ws (global)
  const size_t smem = (size_t)smem_doubles * sizeof(double);
  static bool attr_set = false;
  if (!attr_set) {
    NFDPM_CUDA(cudaFuncSetAttribute(mix_prepare_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    attr_set = true;
  }
  for (int base = 0; base < n; base += kPrepBatch) {
    PrepBatch pb;
    const int cnt = (n - base < kPrepBatch) ? n - base : kPrepBatch;
    for (int i = 0; i < cnt; ++i) pb.it[i] = items[base + i];
    mix_prepare_kernel<<<cnt, 256, smem, as_stream(stream)>>>(pb, smem_doubles);
    NFDPM_CHECK_LAUNCH("mix_prepare_kernel");
  }
  return 0;
}
