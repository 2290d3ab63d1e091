"""The summation plan of the wide chi-square scan (slam-1_b200/csrc/chi2_plan.h) on the CPU: leaves summed the way an 8-lane
group sums them (accumulator j per lane, XOR-shuffle combination) and inner nodes added level by level must give np.sum's
result BIT for bit -- numpy's pairwise summation is what bag_of_words.py:30-31 relies on.  Compiled with g++."""
import os
import shutil
import struct
import subprocess

import numpy as np
import pytest

from conftest import ROOT

HARNESS = r"""
#include "chi2_plan.h"
#include <cstdio>
#include <cstdint>
#include <cstring>
// what chi2_leaf8 (bow.cu) computes, lane by lane
static double leaf8(const double *a, int n)
{
    if (n < 8) { double r = 0.0; for (int i = 0; i < n; ++i) r = r + a[i]; return r; }
    const int n8 = n - n % 8;
    double r[8], s[8];
    for (int j = 0; j < 8; ++j) r[j] = a[j];
    for (int i = 8; i < n8; i += 8) for (int j = 0; j < 8; ++j) r[j] = r[j] + a[i + j];
    for (int o = 1; o < 8; o <<= 1) { for (int j = 0; j < 8; ++j) s[j] = r[j] + r[j ^ o]; memcpy(r, s, sizeof(r)); }
    double res = r[0];
    for (int i = n8; i < n; ++i) res = res + a[i];
    return res;
}
int main(int argc, char **argv)
{
    FILE *f = fopen(argv[1], "rb");
    int n_cases = 0;
    if (fread(&n_cases, 4, 1, f) != 1) return 2;
    for (int c = 0; c < n_cases; ++c) {
        int k = 0;
        if (fread(&k, 4, 1, f) != 1) return 2;
        std::vector<double> a(k);
        if (fread(a.data(), 8, k, f) != (size_t)k) return 2;
        const Chi2Plan p = chi2_plan(k);
        if (!p.ok) return 3;
        std::vector<double> slot(p.n_leaves + p.n_inner, -1.0);
        std::vector<char> done(p.n_leaves + p.n_inner, 0);
        int covered = 0;
        for (int l = 0; l < p.n_leaves; ++l) {
            if (p.table[l].x != covered || p.table[l].y < 1 || p.table[l].y > 128) return 4;     // leaves tile [0, k) in order
            covered += p.table[l].y;
            slot[l] = leaf8(a.data() + p.table[l].x, p.table[l].y);
            done[l] = 1;
        }
        if (covered != k) return 4;
        if (p.levels.start[p.levels.n_levels] != p.n_inner) return 5;
        for (int h = 0; h < p.levels.n_levels; ++h) {
            if (p.levels.start[h + 1] <= p.levels.start[h]) return 5;
            // a level only reads slots finished by earlier levels (so its nodes can run in parallel)
            for (int i = p.levels.start[h]; i < p.levels.start[h + 1]; ++i)
                if (!done[p.table[p.n_leaves + i].x] || !done[p.table[p.n_leaves + i].y]) return 6;
            for (int i = p.levels.start[h]; i < p.levels.start[h + 1]; ++i)
                slot[p.n_leaves + i] = slot[p.table[p.n_leaves + i].x] + slot[p.table[p.n_leaves + i].y];
            for (int i = p.levels.start[h]; i < p.levels.start[h + 1]; ++i) done[p.n_leaves + i] = 1;
        }
        const double v = p.n_inner ? slot[p.n_leaves + p.n_inner - 1] : slot[0];
        uint64_t bits;
        memcpy(&bits, &v, 8);
        printf("%d %016llx %d %d %d\n", k, (unsigned long long)bits, p.n_leaves, p.n_inner, p.levels.n_levels);
    }
    return 0;
}
"""


def test_chi2_plan_reproduces_numpy_pairwise_sum(tmp_path):
    cxx = shutil.which("g++")
    if cxx is None:
        pytest.skip("g++ not found")
    src = tmp_path / "plan.cpp"
    src.write_text(HARNESS)
    exe = tmp_path / "plan"
    r = subprocess.run([cxx, "-O1", "-std=c++17", "-ffp-contract=off", "-I", os.path.join(ROOT, "slam-1_b200", "csrc"), str(src),
                        "-o", str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    rng = np.random.default_rng(5)
    ks = [1, 5, 8, 9, 127, 128, 129, 130, 255, 256, 257, 300, 1000, 1024, 1025, 4099, 8192, 8193, 12288, 12289, 65536, 100003,
          500000, 1 << 19] + [int(x) for x in rng.integers(129, 200000, 60)]
    cases = []
    with open(tmp_path / "cases.bin", "wb") as f:
        f.write(struct.pack("<i", len(ks)))
        for k in ks:
            x, y = rng.integers(0, 50, k), rng.integers(0, 50, k)
            a = (2 * (x - y) ** 2 / np.maximum(1, x + y)).astype(np.float64)          # the terms of bag_of_words.py:30-31
            cases.append(a)
            f.write(struct.pack("<i", k))
            f.write(a.tobytes())
    r = subprocess.run([str(exe), str(tmp_path / "cases.bin")], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, (r.returncode, r.stdout[-500:], r.stderr[-500:])
    lines = r.stdout.split("\n")
    for k, a, line in zip(ks, cases, lines):
        kk, bits, n_leaves, n_inner, n_levels = line.split()
        assert int(kk) == k
        want = struct.unpack("<Q", struct.pack("<d", float(np.sum(a))))[0]
        assert int(bits, 16) == want, (k, bits, hex(want))
        assert int(n_inner) == int(n_leaves) - 1
    # config 4's vocabulary: a perfect tree, 512 leaves of 128 words, 9 levels
    assert lines[ks.index(65536)].split()[2:] == ["512", "511", "9"]
