// Persistent tile schedule of the tcgen05 GEMM (gemm_tc.cu), as plain functions so that the host-side unit test
// (tests/test_sched_cpu.py, compiled with g++) exercises exactly the code the kernel runs.
//
// One unit of work of a CTA: a BM x bn output tile at column n0.  Every CTA runs the same number of full tiles
// (tile = cta + it * grid); when the tile count is not a multiple of the grid (512 tiles on 148 SMs: 68 CTAs would run a 4th
// tile while 80 idle for a whole tile time), the REMAINING tiles are split into two half-width tiles each and handed to
// 2 * rem CTAs (`split` mode: the B tensor-map box is BN/2 rows, a full tile takes two B loads per stage), so the last wave
// costs half a tile on (almost) every SM instead of a whole tile on some.
#pragma once
#if defined(__CUDACC__)
#define TC_SCHED_HD __host__ __device__ __forceinline__
#else
#define TC_SCHED_HD inline
#endif

namespace nfdpm {

struct TcWork { int m_blk, n0, bn; };

// it-th unit of work of CTA `cta` of `grid`; false when the CTA is done
TC_SCHED_HD bool tc_work_for(int cta, int grid, int it, int num_tiles, int num_n, int BN, int split, TcWork& w) {
  int tile = cta + it * grid, half = -1;
  if (split) {
    const int full_per = num_tiles / grid;
    if (it > full_per) return false;
    if (it == full_per) {
      const int rem = num_tiles - full_per * grid;
      if (cta >= 2 * rem) return false;
      tile = full_per * grid + (cta >> 1);
      half = cta & 1;
    }
  }
  if (tile >= num_tiles) return false;
  w.m_blk = tile / num_n;
  const int n_blk = tile - w.m_blk * num_n;
  w.n0 = n_blk * BN + (half > 0 ? (BN >> 1) : 0);
  w.bn = half >= 0 ? (BN >> 1) : BN;
  return true;
}

// may the last round be split?  (half tiles must be whole 64-column store groups and whole UMMA N steps; every N block full)
TC_SCHED_HD bool tc_split_ok(int tiles, int grid, int BN, int N) {
  const int rem = tiles % grid;
  return tiles > grid && rem > 0 && 2 * rem <= grid && BN % 128 == 0 && N % BN == 0;
}

}  // namespace nfdpm
