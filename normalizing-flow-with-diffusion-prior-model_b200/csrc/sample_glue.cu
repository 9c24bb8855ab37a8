// Data formats either side of the flow: HBM-bound byte / permutation work.
//   latent_format_kernel  CatFormater.process_latents / .postprocess (diffusion_prior/latent_formaters.py:163-236):
//                         every squeeze^k / unsqueeze^k + the channel concat (or split) of ALL latent parts in ONE launch
//   postprocess_u8_kernel postprocess_batch (normalizing_flow/utils.py:210): fp32 model space -> uint8 pixels on the
//                         device, so the D2H copy moves 1 byte per value instead of 4
//   preprocess_kernel     preprocess_batch (+ the dequantisation noise add of trainer.py:155) in one pass
#include "common.cuh"

namespace nfdpm {

struct LatentTable {
  nfdpm_latent_part p[NFDPM_MAX_LATENT_PARTS];
  int n;
};

// One thread per element of the concatenated tensor cat [B, Ct, Ht, Wt] (coalesced on that side).
// For a part with degree k > 0 the latent is the FINE tensor [C, Ht<<k, Wt<<k] and its slice of `cat` the k-fold
// squeeze of it; for degree -k the latent is the COARSE tensor [C, Ht>>k, Wt>>k] and the slice its k-fold unsqueeze.
// With "b c (h h1) (w w1) -> b (c h1 w1) h w" applied k times the coarse channel is c*4^k + D, where the base-4 digits
// of D from most to least significant hold bits 0 .. k-1 of the fine (row, column) offset inside the 2^k x 2^k cell.
template <bool TO_CAT>
__global__ void __launch_bounds__(256) latent_format_kernel(const LatentTable tab, float* __restrict__ cat, int Ct, int Ht,
                                                            int Wt, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int w = (int)(i % Wt);
  int64_t r = i / Wt;
  const int h = (int)(r % Ht);
  r /= Ht;
  const int cc = (int)(r % Ct);
  const int64_t b = r / Ct;
  int pi = 0;
#pragma unroll
  for (int q = 1; q < NFDPM_MAX_LATENT_PARTS; ++q)
    if (q < tab.n && cc >= tab.p[q].ch_offset) pi = q;
  const nfdpm_latent_part P = tab.p[pi];
  const int q = cc - P.ch_offset;          // channel inside this part's slice of cat
  int64_t li;                              // element index inside the latent [B, C, H, W]
  if (P.degree > 0) {
    const int k = P.degree;
    const int c = q >> (2 * k);
    int lh = 0, lw = 0;
    for (int j = 0; j < k; ++j) {          // digit j from the top of D <-> bit j
      const int dig = (q >> (2 * (k - 1 - j))) & 3;
      lh |= (dig >> 1) << j;
      lw |= (dig & 1) << j;
    }
    li = ((b * P.C + c) * P.H + ((h << k) | lh)) * (int64_t)P.W + ((w << k) | lw);
  } else if (P.degree < 0) {
    const int k = -P.degree;
    int D = 0;
    for (int j = 0; j < k; ++j) D |= ((((h >> j) & 1) << 1) | ((w >> j) & 1)) << (2 * (k - 1 - j));
    li = ((b * P.C + ((q << (2 * k)) | D)) * P.H + (h >> k)) * (int64_t)P.W + (w >> k);
  } else {
    li = ((b * P.C + q) * P.H + h) * (int64_t)P.W + w;
  }
  if (TO_CAT)
    cat[i] = __ldg(P.ptr + li);
  else
    P.ptr[li] = __ldg(cat + i);
}

// out = (uint8) clip(floor((x + 0.5) * n_bins) * k, 0, 255), the same fp32 operation order as the reference
__device__ __forceinline__ uint32_t post1(float x, float n_bins, float k) {
  float v = floorf((x + 0.5f) * n_bins) * k;
  v = fminf(fmaxf(v, 0.f), 255.f);        // NaN -> 0 (fmaxf returns the non-NaN operand)
  return (uint32_t)v;                     // truncation, like Tensor.to(torch.uint8)
}

__global__ void __launch_bounds__(256) postprocess_u8_kernel(const float* __restrict__ x, uint8_t* __restrict__ out,
                                                             int64_t n, float n_bins, float k) {
  // 16 values per thread: four 16-byte loads, one 16-byte store
  const int64_t i0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 16;
  if (i0 >= n) return;
  if (i0 + 16 <= n) {
    uint32_t w[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float4 v = ldg_stream4(x + i0 + 4 * j);
      w[j] = post1(v.x, n_bins, k) | (post1(v.y, n_bins, k) << 8) | (post1(v.z, n_bins, k) << 16) |
             (post1(v.w, n_bins, k) << 24);
    }
    *reinterpret_cast<uint4*>(out + i0) = make_uint4(w[0], w[1], w[2], w[3]);
  } else {
    for (int64_t i = i0; i < n; ++i) out[i] = (uint8_t)post1(x[i], n_bins, k);
  }
}

// y = floor(x * 255 / 2^(8 - n_bits)) / n_bins - 0.5  [+ noise / n_bins]; IEEE division and no FMA contraction so the
// result is bit-identical to the reference's sequence of torch ops
__device__ __forceinline__ float pre1(float x, float inv_q, int quantise, float n_bins) {
  float v = __fmul_rn(x, 255.f);
  if (quantise) v = floorf(__fmul_rn(v, inv_q));      // 2^(8-n_bits) is a power of two: the product is exact
  return __fsub_rn(__fdiv_rn(v, n_bins), 0.5f);
}

__global__ void __launch_bounds__(256) preprocess_kernel(const float* __restrict__ x, const float* __restrict__ noise,
                                                         float* __restrict__ y, int64_t n, float inv_q, int quantise,
                                                         float n_bins) {
  const int64_t i0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i0 >= n) return;
  if (i0 + 4 <= n) {
    const float4 v = ldg_stream4(x + i0);
    float4 o = make_float4(pre1(v.x, inv_q, quantise, n_bins), pre1(v.y, inv_q, quantise, n_bins),
                           pre1(v.z, inv_q, quantise, n_bins), pre1(v.w, inv_q, quantise, n_bins));
    if (noise) {
      const float4 u = ldg_stream4(noise + i0);
      o.x = __fadd_rn(o.x, __fdiv_rn(u.x, n_bins));
      o.y = __fadd_rn(o.y, __fdiv_rn(u.y, n_bins));
      o.z = __fadd_rn(o.z, __fdiv_rn(u.z, n_bins));
      o.w = __fadd_rn(o.w, __fdiv_rn(u.w, n_bins));
    }
    stg_stream4(y + i0, o);
  } else {
    for (int64_t i = i0; i < n; ++i) {
      float o = pre1(x[i], inv_q, quantise, n_bins);
      if (noise) o = __fadd_rn(o, __fdiv_rn(noise[i], n_bins));
      y[i] = o;
    }
  }
}

}  // namespace nfdpm

using namespace nfdpm;

extern "C" int nfdpm_latent_format(const nfdpm_latent_part* parts, int n_parts, float* cat, int B, int Ct, int Ht, int Wt,
                                   int to_cat, nfdpm_stream_t stream) {
  NFDPM_REQUIRE(parts && cat, "nfdpm_latent_format: null pointer");
  NFDPM_REQUIRE(n_parts >= 1 && n_parts <= NFDPM_MAX_LATENT_PARTS, "nfdpm_latent_format: n_parts=%d outside 1..%d", n_parts,
                NFDPM_MAX_LATENT_PARTS);
  NFDPM_REQUIRE(B > 0 && Ct > 0 && Ht > 0 && Wt > 0, "nfdpm_latent_format: bad shape");
  LatentTable tab;
  tab.n = n_parts;
  int off = 0;
  for (int i = 0; i < n_parts; ++i) {
    const nfdpm_latent_part& p = parts[i];
    NFDPM_REQUIRE(p.ptr, "nfdpm_latent_format: part %d has a null pointer", i);
    NFDPM_REQUIRE(p.C > 0 && p.H > 0 && p.W > 0, "nfdpm_latent_format: part %d has a bad shape", i);
    NFDPM_REQUIRE(p.degree >= -8 && p.degree <= 8, "nfdpm_latent_format: part %d degree %d", i, p.degree);
    int want_c, ok;
    if (p.degree >= 0) {
      ok = p.H == (Ht << p.degree) && p.W == (Wt << p.degree);
      want_c = p.C << (2 * p.degree);
    } else {
      const int k = -p.degree;
      ok = (p.H << k) == Ht && (p.W << k) == Wt && p.C % (1 << (2 * k)) == 0;
      want_c = p.C >> (2 * k);
    }
    NFDPM_REQUIRE(ok, "nfdpm_latent_format: part %d [%d,%d,%d] with degree %d does not map onto %dx%d", i, p.C, p.H, p.W,
                  p.degree, Ht, Wt);
    NFDPM_REQUIRE(p.ch_offset == off && p.ch_count == want_c,
                  "nfdpm_latent_format: part %d covers channels [%d,%d), expected [%d,%d)", i, p.ch_offset,
                  p.ch_offset + p.ch_count, off, off + want_c);
    off += want_c;
    tab.p[i] = p;
  }
  NFDPM_REQUIRE(off == Ct, "nfdpm_latent_format: parts cover %d channels, cat has %d", off, Ct);
  const int64_t n = (int64_t)B * Ct * Ht * Wt;
  const unsigned grid = (unsigned)cdiv64(n, 256);
  if (to_cat)
    latent_format_kernel<true><<<grid, 256, 0, as_stream(stream)>>>(tab, cat, Ct, Ht, Wt, n);
  else
    latent_format_kernel<false><<<grid, 256, 0, as_stream(stream)>>>(tab, cat, Ct, Ht, Wt, n);
  NFDPM_CHECK_LAUNCH("latent_format_kernel");
  return 0;
}

extern "C" int nfdpm_postprocess_u8(const float* x, uint8_t* out, int64_t n, float n_bins, float out_scale,
                                    nfdpm_stream_t stream) {
  NFDPM_REQUIRE(x && out, "nfdpm_postprocess_u8: null pointer");
  NFDPM_REQUIRE(n > 0, "nfdpm_postprocess_u8: empty tensor");
  NFDPM_REQUIRE(((uintptr_t)x % 16) == 0 && ((uintptr_t)out % 16) == 0, "nfdpm_postprocess_u8: 16-byte alignment required");
  postprocess_u8_kernel<<<(unsigned)cdiv64(n, 256 * 16), 256, 0, as_stream(stream)>>>(x, out, n, n_bins, out_scale);
  NFDPM_CHECK_LAUNCH("postprocess_u8_kernel");
  return 0;
}

extern "C" int nfdpm_preprocess(const float* x, const float* noise, float* y, int64_t n, int n_bits, float n_bins,
                                nfdpm_stream_t stream) {
  NFDPM_REQUIRE(x && y, "nfdpm_preprocess: null pointer");
  NFDPM_REQUIRE(n > 0, "nfdpm_preprocess: empty tensor");
  NFDPM_REQUIRE(n_bits >= 1 && n_bits <= 8, "nfdpm_preprocess: n_bits=%d outside 1..8", n_bits);
  NFDPM_REQUIRE(n_bins > 0.f, "nfdpm_preprocess: n_bins must be positive");
  NFDPM_REQUIRE(((uintptr_t)x % 16) == 0 && ((uintptr_t)y % 16) == 0 && (!noise || ((uintptr_t)noise % 16) == 0),
                "nfdpm_preprocess: 16-byte alignment required");
  const float inv_q = 1.0f / (float)(1 << (8 - n_bits));
  preprocess_kernel<<<(unsigned)cdiv64(n, 256 * 4), 256, 0, as_stream(stream)>>>(x, noise, y, n, inv_q, n_bits < 8 ? 1 : 0,
                                                                               n_bins);
  NFDPM_CHECK_LAUNCH("preprocess_kernel");
  return 0;
}
