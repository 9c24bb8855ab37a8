"""Host-side unit test of the tcgen05 GEMM's persistent tile schedule (csrc/tc_sched.h — the same functions the kernel
calls, compiled here with g++): for every shape, every output element of the [num_m*128, N] result is produced by
exactly one unit of work, every CTA's work list is the same for its three warp roles by construction, half tiles appear
only in the last round and only where the host enabled them, and the last-wave split never makes the longest CTA longer."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HDR = os.path.join(ROOT, "normalizing-flow-with-diffusion-prior-model_b200", "csrc")

HARNESS = r"""
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "tc_sched.h"
using namespace nfdpm;
int main(int argc, char** argv) {
  const int grids[] = {1, 2, 7, 64, 132, 148};
  const int bns[] = {16, 64, 112, 128, 256};
  long long cases = 0, split_cases = 0;
  for (int grid_sms : grids)
    for (int BN : bns)
      for (int num_n = 1; num_n <= 4; ++num_n)
        for (int num_m = 1; num_m <= 300; num_m += (num_m < 40 ? 1 : 37)) {
          const int N = num_n * BN, tiles = num_m * num_n, grid = tiles < grid_sms ? tiles : grid_sms;
          for (int want_split = 0; want_split <= 1; ++want_split) {
            const int split = want_split && tc_split_ok(tiles, grid, BN, N);
            if (want_split && !split) continue;
            ++cases; split_cases += split;
            std::vector<int> cover((size_t)num_m * N, 0);
            long long longest = 0;                       // in half-tile units
            for (int cta = 0; cta < grid; ++cta) {
              TcWork w; long long cost = 0; int it = 0, seen_half = 0;
              for (; tc_work_for(cta, grid, it, tiles, num_n, BN, split, w); ++it) {
                if (w.bn != BN) {
                  if (!split || w.bn * 2 != BN || (w.bn % 64) != 0) { printf("bad half tile\n"); return 1; }
                  seen_half = 1; cost += 1;
                } else {
                  if (seen_half) { printf("full tile after a half tile\n"); return 1; }
                  cost += 2;
                }
                if (w.m_blk < 0 || w.m_blk >= num_m || w.n0 < 0 || w.n0 + w.bn > N) { printf("out of range\n"); return 1; }
                for (int c = 0; c < w.bn; ++c) cover[(size_t)w.m_blk * N + w.n0 + c] += 1;
              }
              TcWork w2;                                  // once exhausted, the schedule stays exhausted
              if (tc_work_for(cta, grid, it + 1, tiles, num_n, BN, split, w2)) { printf("not exhausted\n"); return 1; }
              if (cost > longest) longest = cost;
            }
            for (int v : cover) if (v != 1) { printf("coverage %d (grid %d BN %d num_n %d num_m %d split %d)\n", v, grid, BN, num_n, num_m, split); return 1; }
            const long long unsplit_longest = 2LL * ((tiles + grid - 1) / grid);
            if (longest > unsplit_longest) { printf("split made the longest CTA longer\n"); return 1; }
            if (split && longest >= unsplit_longest) { printf("split did not shorten the last wave\n"); return 1; }
          }
        }
  // the case the split was written for: M = 32768, N = 512, BN = 256 on 148 SMs
  if (!tc_split_ok(512, 148, 256, 512)) { printf("headline shape not split\n"); return 1; }
  printf("ok %lld cases, %lld split\n", cases, split_cases);
  return 0;
}
"""


def test_tile_schedule_covers_every_output_exactly_once(tmp_path):
    src = tmp_path / "sched.cpp"
    src.write_text(HARNESS)
    exe = tmp_path / "sched"
    r = subprocess.run(["g++", "-std=c++17", "-O1", "-I", HDR, str(src), "-o", str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:]
    assert r.stdout.startswith("ok") and int(r.stdout.split()[3]) > 0, r.stdout
