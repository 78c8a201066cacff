"""The device introsort replay (lvt_b200/csrc/introsort.cuh) walks std::sort's permutation."""
import os
import subprocess

from helpers import ROOT

SRC = r"""
#include "introsort.cuh"
#include <algorithm>
#include <cstdio>
#include <random>
#include <vector>
static bool cmp(uint32_t x, uint32_t y) { return (x >> 24) > (y >> 24); }
int main() {
  std::mt19937 rng(1); long bad = 0, cases = 0;
  for (int it = 0; it < 6000; it++) {
    int n = 1 + rng() % (it % 50 == 0 ? 6000 : 700), spread = 1 + rng() % 230, mode = it % 7;
    std::vector<uint32_t> a(n);
    for (int i = 0; i < n; i++) {
      uint32_t r = 20 + rng() % spread;
      if (mode == 3) r = 20 + (long)i * spread / n;
      if (mode == 4) r = 20 + (long)(n - 1 - i) * spread / n;
      if (mode == 5) r = 20 + (long)(i < n / 2 ? i : n - i) * spread / n;
      if (mode == 6) r = 20 + rng() % 3;
      a[i] = (r << 24) | i;
    }
    std::vector<uint32_t> b = a, c = a, orig = a;
    std::sort(a.begin(), a.end(), cmp);
    lvtb::isort::sort(b.data(), n);
    std::vector<lvtb::isort::LevelRange> q0(n / 8 + 2), q1(n / 8 + 2);
    lvtb::isort::sort_levels(c.data(), n, q0.data(), q1.data());   // the level-synchronous schedule of tile_kernel
    std::vector<uint32_t> d2 = orig;
    std::vector<uint16_t> pl(n + 1), pr(n + 1);
    lvtb::isort::sort_levels(d2.data(), n, q0.data(), q1.data(), pl.data(), pr.data(), 17);  // list-based partition (warp_partition)
    cases++; bad += a != b; bad += a != c; bad += a != d2;
  }
  // the heapsort fallback, exercised directly against std::make_heap + std::sort_heap
  for (int it = 0; it < 500; it++) {
    int n = 2 + rng() % 400;
    std::vector<uint32_t> a(n);
    for (int i = 0; i < n; i++) a[i] = ((20 + rng() % 40) << 24) | i;
    std::vector<uint32_t> b = a;
    std::make_heap(a.begin(), a.end(), cmp); std::sort_heap(a.begin(), a.end(), cmp);
    lvtb::isort::heap_sort(b.data(), b.data() + n);
    cases++; bad += a != b;
  }
  printf("%ld %ld\n", cases, bad);
  return bad != 0;
}
"""


def test_introsort_replays_std_sort(tmp_path):
    src = tmp_path / "t.cpp"
    src.write_text(SRC)
    exe = tmp_path / "t"
    subprocess.run(["g++", "-O2", "-std=c++17", "-I", os.path.join(ROOT, "lvt_b200", "csrc"), "-o", str(exe), str(src)], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout
    assert r.stdout.split()[1] == "0"
