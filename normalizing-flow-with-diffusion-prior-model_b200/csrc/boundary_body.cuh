// The arithmetic of one step boundary for ONE image, shared by flow_boundary_kernel (one CTA per image) and the fused
// deep-level StepFlow kernel (deep_step.cu).  See flow_boundary.cu for the description.  Internal header.
#pragma once
#include "common.cuh"

namespace nfdpm {


struct BoundaryArgs {
  const float* in; int64_t in_bs;       // source state [B,C,P] (or [B,C/4,2H,2W] when squeeze_in)
  const float* pm; int64_t ldp;         // SRC_COUPLING: taps-as-N rows [B*P, ldp]
  const float* bias3; const float* logs3;
  float* ld_part;                       // forward coupling: [B] per-image log-det partial (may be null)
  const float* mt; const float* beta;   // mix (null = identity)
  float* y; int64_t y_bs;               // NCHW sink (may be null)
  float* xs; int64_t xs_bs;             // NCHW sink of the PRE-mix state (training stash of the next step's input; may be null)
  void* a1; int64_t lda1;               // im2col sink (may be null)
  int B, C, H, W;
  int squeeze_in, inverse;
};

template <typename T> __device__ __forceinline__ void store8(T* p, const float (&v)[8]);
template <> __device__ __forceinline__ void store8<float>(float* p, const float (&v)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
template <> __device__ __forceinline__ void store8<__nv_bfloat16>(__nv_bfloat16* p, const float (&v)[8]) {
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __nv_bfloat162 t = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    w[i] = *reinterpret_cast<uint32_t*>(&t);
  }
  *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
}


// `sm`: shared-memory scratch of nfdpm_flow_boundary_smem() bytes; all `nt` threads of the CTA must call this together.
// PM_SMEM: the taps-as-N rows of image b are already in shared memory at `pm_img` (row stride `pm_ld` floats) instead of
// a.pm in global memory (deep_step.cu).
template <bool COUPLING, typename A1T, bool PM_SMEM = false>
__device__ __forceinline__ void flow_boundary_body(const BoundaryArgs& a, const int b, float* sm, const int tid, const int nt,
                                                   const float* pm_img = nullptr, const int pm_ld = 0) {

  const int C = a.C, H = a.H, W = a.W, P = H * W, Ch = C >> 1;
  const int PS = P + 1;                       // padded pixel stride: conflict-free for lanes over channels
  const int Cp = (C + 3) & ~3;
  float* x_s = sm;                            // [C][PS]  source / coupling result
  float* u_s = x_s + C * PS;                  // [C][PS]  mixed result (aliases x_s when there is no mix)
  float* m_s = (a.mt != nullptr) ? u_s + C * PS : u_s;      // [C][Cp] + beta [Cp]
  float* par_s = m_s + ((a.mt != nullptr) ? (C * Cp + Cp) : 0);   // [2C] bias3, exp(3 logs3)
  float* ls_s = par_s + (COUPLING ? 2 * C : 0);                   // [P*Ch] log-det terms
  if (a.mt == nullptr) u_s = x_s;

  // ---- parameters
  if (a.mt != nullptr) {
    for (int i = tid; i < C * Cp; i += nt) {
      const int r = i / Cp, c = i - r * Cp;
      m_s[i] = (c < C) ? a.mt[r * C + c] : 0.f;
    }
    for (int i = tid; i < Cp; i += nt) m_s[C * Cp + i] = (i < C) ? a.beta[i] : 0.f;
  }
  if (COUPLING) {
    for (int i = tid; i < C; i += nt) {
      par_s[i] = a.bias3[i];
      par_s[C + i] = expf(3.f * a.logs3[i]);
    }
  }
  // ---- phase 0: stage the image channel-major (lanes over pixels -> coalesced)
  const float* inb = a.in + (int64_t)b * a.in_bs;
  if (a.squeeze_in) {
    // in is [C/4, 2H, 2W]; channel c = cc*4 + h1*2 + w1 reads in[cc, 2y+h1, 2x+w1]
    const int W2 = 2 * W;
    for (int i = tid; i < (C >> 2) * P; i += nt) {
      const int cc = i / P, p = i - cc * P;
      const int py = p / W, px = p - py * W;
      const float* s = inb + ((int64_t)cc * 2 * H + 2 * py) * W2 + 2 * px;
      const float2 t0 = *reinterpret_cast<const float2*>(s), t1 = *reinterpret_cast<const float2*>(s + W2);
      x_s[(cc * 4 + 0) * PS + p] = t0.x;
      x_s[(cc * 4 + 1) * PS + p] = t0.y;
      x_s[(cc * 4 + 2) * PS + p] = t1.x;
      x_s[(cc * 4 + 3) * PS + p] = t1.y;
    }
  } else {
    for (int i = tid; i < C * P; i += nt) {
      const int c = i / P, p = i - c * P;
      x_s[c * PS + p] = inb[i];
    }
  }
  __syncthreads();

  // ---- phase 1: affine coupling, item = (pixel, j) with j fastest (pm rows read contiguously)
  if (COUPLING) {
    const float* pmb = PM_SMEM ? pm_img : a.pm + (int64_t)b * P * a.ldp;
    const int64_t ldp = PM_SMEM ? (int64_t)pm_ld : a.ldp;
    for (int it = tid; it < P * Ch; it += nt) {
      const int p = it / Ch, j = it - p * Ch;
      const int py = p / W, px = p - py * W;
      // all 18 loads are issued before the first use (each pm element is consumed exactly once: plain streaming reads)
      float lv[9], tv[9];
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        const int yy = py + tap / 3 - 1, xx = px + tap % 3 - 1;
        const bool ok = (yy >= 0) && (yy < H) && (xx >= 0) && (xx < W);
        const float* r = pmb + (int64_t)(ok ? yy * W + xx : p) * ldp + tap * C + j;
        const float l0 = PM_SMEM ? r[0] : __ldg(r), t0 = PM_SMEM ? r[Ch] : __ldg(r + Ch);
        lv[tap] = ok ? l0 : 0.f;
        tv[tap] = ok ? t0 : 0.f;
      }
      float ls = 0.f, tt = 0.f;
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        ls += lv[tap];
        tt += tv[tap];
      }
      const float log_s = (ls + par_s[j]) * par_s[C + j];
      const float sh_t = (tt + par_s[Ch + j]) * par_s[C + Ch + j];
      const float s = 1.f / (1.f + expf(-(log_s + 2.f)));
      const float xb = x_s[(Ch + j) * PS + p];
      if (a.inverse) {
        x_s[(Ch + j) * PS + p] = xb / (s + 1e-6f) - sh_t;
      } else {
        x_s[(Ch + j) * PS + p] = (xb + sh_t) * s;
        ls_s[it] = logf(s + 1e-6f);
      }
    }
    __syncthreads();
    if (!a.inverse && a.ld_part != nullptr && tid < 32) {
      // deterministic per-image sum: fixed lane-strided order + shuffle tree
      float acc = 0.f;
      for (int i = tid; i < P * Ch; i += 32) acc += ls_s[i];
      acc = warp_sum(acc);
      if (tid == 0) a.ld_part[b] = acc;
    }
  }

  // ---- phase 1b: stash the pre-mix state (the next StepFlow's input, needed by its backward)
  if (a.xs != nullptr) {
    float* xb = a.xs + (int64_t)b * a.xs_bs;
    for (int i = tid; i < C * P; i += nt) {
      const int c = i / P, p = i - c * P;
      xb[i] = x_s[c * PS + p];
    }
  }

  // ---- phase 2: channel mix, item = (group of 4 outputs, pixel), lanes over pixels
  if (a.mt != nullptr) {
    const int n_og = Cp >> 2;
    for (int it = tid; it < n_og * P; it += nt) {
      const int og = it / P, p = it - og * P;
      const float4 b4 = *reinterpret_cast<const float4*>(m_s + C * Cp + og * 4);
      float a0 = b4.x, a1 = b4.y, a2 = b4.z, a3 = b4.w;
      for (int c = 0; c < C; ++c) {
        const float xv = x_s[c * PS + p];
        const float4 w = *reinterpret_cast<const float4*>(m_s + c * Cp + og * 4);
        a0 = fmaf(w.x, xv, a0);
        a1 = fmaf(w.y, xv, a1);
        a2 = fmaf(w.z, xv, a2);
        a3 = fmaf(w.w, xv, a3);
      }
      const int o = og * 4;
      u_s[o * PS + p] = a0;
      if (o + 1 < C) u_s[(o + 1) * PS + p] = a1;
      if (o + 2 < C) u_s[(o + 2) * PS + p] = a2;
      if (o + 3 < C) u_s[(o + 3) * PS + p] = a3;
    }
    __syncthreads();
  }

  // ---- phase 3a: NCHW sink (lanes over pixels)
  if (a.y != nullptr) {
    float* yb = a.y + (int64_t)b * a.y_bs;
    for (int i = tid; i < C * P; i += nt) {
      const int c = i / P, p = i - c * P;
      yb[i] = u_s[c * PS + p];
    }
  }
  // ---- phase 3b: im2col sink, item = (pixel, 8-column group), group fastest -> 128-byte runs per row
  if (a.a1 != nullptr) {
    const int K = Ch * 9;
    const int n_g = (int)(a.lda1 >> 3);
    A1T* a1b = reinterpret_cast<A1T*>(a.a1) + (int64_t)b * P * a.lda1;
    for (int it = tid; it < P * n_g; it += nt) {
      const int p = it / n_g, g = it - p * n_g;
      const int py = p / W, px = p - py * W;
      float v[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int k = g * 8 + e;
        float val = 0.f;
        if (k < K) {
          const int c = k / 9, tap = k - c * 9;
          const int ky = tap / 3, kx = tap - ky * 3;
          const int yy = py + ky - 1, xx = px + kx - 1;
          if (yy >= 0 && yy < H && xx >= 0 && xx < W) val = u_s[c * PS + yy * W + xx];
        }
        v[e] = val;
      }
      store8<A1T>(a1b + (int64_t)p * a.lda1 + g * 8, v);
    }
  }
}

}  // namespace nfdpm
