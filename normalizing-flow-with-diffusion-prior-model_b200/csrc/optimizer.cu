// Fused optimiser step of the reference's training recipe (normalizing_flow/trainer.py:165-167):
//     torch.nn.utils.clip_grad_value_(flow.parameters(), 1); torch.nn.utils.clip_grad_norm_(flow.parameters(), 1);
//     optimizer.step()            # Adam / AdamW (utils.py:120-137)
// as THREE launches over all parameter tensors (multi-tensor: a device table of tensor references + a chunk table)
// instead of ~1300 foreach / elementwise launches:
//   opt_clip_sumsq_kernel   g <- clamp(g, -c, c) in place for the clip group; per-chunk sum of squares
//   opt_finalize_kernel     total norm (fixed summation order), clip coefficient min(1, max_norm/(norm+1e-6)), step += 1
//   opt_adam_kernel         g <- g*coef (clip group, in place, like clip_grad_norm_); Adam moments and parameter update
// All of it is HBM-bound streaming: 4 fp32 reads + 4 writes per parameter element.
#include "common.cuh"

namespace nfdpm {

constexpr int OPT_CHUNK = 4096;       // elements per CTA
constexpr int OPT_THREADS = 256;

struct OptRef {                       // one parameter tensor (device table, 32 bytes)
  float* p;                           // parameter
  float* g;                           // gradient (NULL: tensor skipped this step)
  int64_t state_off;                  // offset of exp_avg / exp_avg_sq inside the flat state buffers
  int32_t numel;
  int32_t clip;                       // 1: member of the clip group
};

__device__ __forceinline__ float block_sum_256(float v, float* sh) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) sh[w] = v;
  __syncthreads();
  float t = 0.f;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < OPT_THREADS / 32; ++i) t += sh[i];
  }
  return t;   // valid in thread 0
}

// clip_grad_value_ semantics of torch.clamp: a NaN gradient stays NaN (fminf / fmaxf would return the non-NaN operand and
// turn a diverged step into a silent update by -clip_value); the NaN then reaches the norm and the parameters, as in the
// reference's torch.nn.utils.clip_grad_value_ / clip_grad_norm_ (trainer.py:165-166).
__device__ __forceinline__ float clampv(float v, float c) { return v < -c ? -c : (v > c ? c : v); }

__global__ void __launch_bounds__(OPT_THREADS) opt_clip_sumsq_kernel(const OptRef* __restrict__ refs,
                                                                    const int2* __restrict__ chunks, float clip_value,
                                                                    float* __restrict__ partial) {
  __shared__ float sh[OPT_THREADS / 32];
  const int2 ck = chunks[blockIdx.x];
  const OptRef r = refs[ck.x];
  float acc = 0.f;
  if (r.g != nullptr && r.clip) {
    const int n = min(OPT_CHUNK, r.numel - ck.y);
    float* g = r.g + ck.y;
    if ((((uintptr_t)g) & 15) == 0) {
      const int n4 = n >> 2;
      for (int i = threadIdx.x; i < n4; i += OPT_THREADS) {
        float4 v = reinterpret_cast<float4*>(g)[i];
        v.x = clampv(v.x, clip_value);
        v.y = clampv(v.y, clip_value);
        v.z = clampv(v.z, clip_value);
        v.w = clampv(v.w, clip_value);
        reinterpret_cast<float4*>(g)[i] = v;
        acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
      }
      for (int i = (n4 << 2) + threadIdx.x; i < n; i += OPT_THREADS) {
        const float v = clampv(g[i], clip_value);
        g[i] = v;
        acc += v * v;
      }
    } else {
      for (int i = threadIdx.x; i < n; i += OPT_THREADS) {
        const float v = clampv(g[i], clip_value);
        g[i] = v;
        acc += v * v;
      }
    }
  }
  const float t = block_sum_256(acc, sh);
  if (threadIdx.x == 0) partial[blockIdx.x] = t;
}

// scal: [0] clip coefficient, [1] total gradient norm of the clip group, [2] step count (float, like torch capturable Adam)
__global__ void __launch_bounds__(1024) opt_finalize_kernel(const float* __restrict__ partial, int n_chunks, float max_norm,
                                                            float* __restrict__ scal) {
  __shared__ double sh[1024];
  double acc = 0.0;
  for (int i = threadIdx.x; i < n_chunks; i += 1024) acc += (double)partial[i];
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 512; s > 0; s >>= 1) {
    if (threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const float norm = (float)sqrt(sh[0]);
    scal[1] = norm;
    scal[0] = (max_norm > 0.f) ? fminf(1.f, max_norm / (norm + 1e-6f)) : 1.f;
    scal[2] += 1.f;
  }
}

struct AdamHyper { double lr, beta1, beta2, eps, weight_decay; int decoupled; };

__global__ void __launch_bounds__(OPT_THREADS) opt_adam_kernel(const OptRef* __restrict__ refs, const int2* __restrict__ chunks,
                                                              float* __restrict__ exp_avg, float* __restrict__ exp_avg_sq,
                                                              const float* __restrict__ scal, const AdamHyper hp) {
  const int2 ck = chunks[blockIdx.x];
  const OptRef r = refs[ck.x];
  if (r.g == nullptr) return;
  const float coef = r.clip ? scal[0] : 1.f;
  const double step = (double)scal[2];
  // scalars are formed in double and rounded once, the way torch's Python-side Adam forms them
  const float step_size = (float)(hp.lr / (1.0 - pow(hp.beta1, step)));
  const float bc2_sqrt = (float)sqrt(1.0 - pow(hp.beta2, step));
  const float decay = hp.decoupled ? (float)(1.0 - hp.lr * hp.weight_decay) : 1.f;
  const float b2 = (float)hp.beta2, omb1 = (float)(1.0 - hp.beta1), omb2 = (float)(1.0 - hp.beta2);
  const float eps = (float)hp.eps, wd = (float)hp.weight_decay;
  const int n = min(OPT_CHUNK, r.numel - ck.y);
  float* g = r.g + ck.y;
  float* p = r.p + ck.y;
  float* m = exp_avg + r.state_off + ck.y;
  float* v = exp_avg_sq + r.state_off + ck.y;
  auto upd = [&](float& pv, float& gv, float& mv, float& vv) {
    gv *= coef;
    float ge = gv;
    if (!hp.decoupled && wd != 0.f) ge = fmaf(wd, pv, ge);
    pv *= decay;
    mv = fmaf(omb1, ge - mv, mv);                             // exp_avg.lerp_(grad, 1 - beta1)
    vv = fmaf(omb2 * ge, ge, b2 * vv);                        // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
    const float denom = sqrtf(vv) / bc2_sqrt + eps;
    pv -= step_size * (mv / denom);
  };
  const bool al = (((uintptr_t)g | (uintptr_t)p | (uintptr_t)m | (uintptr_t)v) & 15) == 0;
  int done = 0;
  if (al) {
    const int n4 = n >> 2;
    for (int i = threadIdx.x; i < n4; i += OPT_THREADS) {
      float4 pv = reinterpret_cast<float4*>(p)[i], gv = reinterpret_cast<float4*>(g)[i];
      float4 mv = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
      upd(pv.x, gv.x, mv.x, vv.x);
      upd(pv.y, gv.y, mv.y, vv.y);
      upd(pv.z, gv.z, mv.z, vv.z);
      upd(pv.w, gv.w, mv.w, vv.w);
      reinterpret_cast<float4*>(p)[i] = pv;
      if (r.clip) reinterpret_cast<float4*>(g)[i] = gv;
      reinterpret_cast<float4*>(m)[i] = mv;
      reinterpret_cast<float4*>(v)[i] = vv;
    }
    done = n4 << 2;
  }
  for (int i = done + threadIdx.x; i < n; i += OPT_THREADS) {
    float pv = p[i], gv = g[i], mv = m[i], vv = v[i];
    upd(pv, gv, mv, vv);
    p[i] = pv;
    if (r.clip) g[i] = gv;
    m[i] = mv;
    v[i] = vv;
  }
}

}  // namespace nfdpm

using namespace nfdpm;

extern "C" int nfdpm_opt_chunk(void) { return OPT_CHUNK; }

extern "C" int nfdpm_fused_clip_adam(const void* refs, const int32_t* chunks, int n_chunks, float* exp_avg,
                                     float* exp_avg_sq, float* partial, float* scal, float clip_value, float max_norm,
                                     double lr, double beta1, double beta2, double eps, double weight_decay, int decoupled,
                                     nfdpm_stream_t stream) {
  NFDPM_REQUIRE(refs && chunks && exp_avg && exp_avg_sq && partial && scal, "nfdpm_fused_clip_adam: null pointer");
  NFDPM_REQUIRE(n_chunks > 0, "nfdpm_fused_clip_adam: no chunks");
  static_assert(sizeof(OptRef) == 32, "OptRef must be 32 bytes (the host builds it as 4 x int64)");
  cudaStream_t st = as_stream(stream);
  const OptRef* r = reinterpret_cast<const OptRef*>(refs);
  const int2* ck = reinterpret_cast<const int2*>(chunks);
  opt_clip_sumsq_kernel<<<n_chunks, OPT_THREADS, 0, st>>>(r, ck, clip_value > 0.f ? clip_value : 3.0e38f, partial);
  NFDPM_CHECK_LAUNCH("opt_clip_sumsq_kernel");
  opt_finalize_kernel<<<1, 1024, 0, st>>>(partial, n_chunks, max_norm, scal);
  NFDPM_CHECK_LAUNCH("opt_finalize_kernel");
  AdamHyper hp{lr, beta1, beta2, eps, weight_decay, decoupled};
  opt_adam_kernel<<<n_chunks, OPT_THREADS, 0, st>>>(r, ck, exp_avg, exp_avg_sq, scal, hp);
  NFDPM_CHECK_LAUNCH("opt_adam_kernel");
  return 0;
}
