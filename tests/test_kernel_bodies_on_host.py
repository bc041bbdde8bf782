"""CPU: the bodies of the kernels written after this round's GPU minutes were spent (count_emit / count_flag of csrc/hga_count.cu,
enr_merge2_keys / enr_purge2 of csrc/hga_enrich.cu), compiled FOR THE HOST and run as one thread with a one-thread grid: the text of
each kernel is cut out of the .cu file at test time (nothing is copied into the repo), `__global__`, `blockIdx`, `atomicMax`, ... are
defined away in a few lines, and the results are compared with numpy. This checks the arithmetic and the indexing of the kernel
code itself; it says nothing about launches, CUB calls or buffers - the GPU tests (tests/test_zz_gpu_tail_block.py) do that.
Test infrastructure only: the product never runs this way."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import datagen
from test_second_merge_rule import second_merge_rule
from test_sdk_selection_cpu import exact_counts

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "hybrid-genome-assembler_b200", "csrc")

PRELUDE = r"""
#include <cstdint>
#include <algorithm>
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
struct Dim { unsigned x, y, z; };
static Dim blockIdx{0, 0, 0}, threadIdx{0, 0, 0}, blockDim{1, 1, 1}, gridDim{1, 1, 1};
static inline uint32_t atomicMax(uint32_t *p, uint32_t v) { uint32_t o = *p; if (v > o) *p = v; return o; }
using std::max; using std::min;
"""


def _cut(text, start_pattern):
    """the source of one function / template / constant, from its first line to the closing brace at column 0"""
    m = re.search(start_pattern, text, flags=re.M)
    assert m, start_pattern
    end = text.index("\n}\n", m.start()) + 3
    return text[m.start():end]


@pytest.fixture(scope="module")
def host_kernels(tmp_path_factory):
    count = open(os.path.join(CSRC, "hga_count.cu")).read()
    enrich = open(os.path.join(CSRC, "hga_enrich.cu")).read()
    parts = [PRELUDE,
             re.search(r"^constexpr unsigned long long COUNT_SENTINEL.*$", count, flags=re.M).group(0),
             _cut(count, r"^__device__ __forceinline__ int count_base_code"),
             _cut(count, r"^__global__ void count_emit_kernel"),
             _cut(count, r"^__global__ void count_flag_kernel"),
             _cut(enrich, r"^__global__ void enr_merge2_keys_kernel"),
             _cut(enrich, r"^template<bool FILL>\n__global__ void enr_purge2_kernel"),
             r"""
extern "C" {
void run_count_emit(const char *bases, const uint64_t *read_off, uint64_t n_reads, int k, uint64_t p0, uint64_t p1, unsigned long long *out) {
    count_emit_kernel(bases, read_off, n_reads, k, p0, p1, out);
}
void run_count_flag(const unsigned long long *keys, const uint32_t *cnt, uint64_t n, uint32_t min_count, uint8_t *flag) { count_flag_kernel(keys, cnt, n, min_count, flag); }
void run_merge2_keys(const uint64_t *ukeys, uint64_t n, const uint32_t *map, const uint32_t *surv_old, const uint32_t *surv_new, uint32_t *R2, uint64_t *out) {
    enr_merge2_keys_kernel(ukeys, n, map, surv_old, surv_new, R2, out);
}
void run_purge2_count(const uint32_t *p_off, const uint32_t *p_row, uint32_t n_slots, const uint32_t *R2, const uint8_t *once, uint32_t *cnt) {
    enr_purge2_kernel<false>(p_off, p_row, n_slots, R2, once, cnt, nullptr, nullptr);
}
void run_purge2_fill(const uint32_t *p_off, const uint32_t *p_row, uint32_t n_slots, const uint32_t *R2, const uint8_t *once, const uint32_t *out_off, uint32_t *out_row) {
    enr_purge2_kernel<true>(p_off, p_row, n_slots, R2, once, nullptr, out_off, out_row);
}
}
"""]
    d = tmp_path_factory.mktemp("host_kernels")
    src, so = str(d / "k.cpp"), str(d / "k.so")
    open(src, "w").write("\n".join(parts))
    subprocess.run(["g++", "-std=c++17", "-O1", "-fPIC", "-shared", "-o", so, src], check=True)
    return C.CDLL(so)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.mark.parametrize("k", [1, 5, 19, 31, 32])
def test_count_emit_and_flag_bodies(host_kernels, k):
    g = datagen.random_genome(1500, 10 + k)
    reads = [datagen.to_ascii(r) for r in datagen.sample_reads(g, 60, 120, 20 + k, error_rate=0.02)]
    reads[2] = reads[2][:30] + "N" + reads[2][31:]
    reads[4] = reads[4].lower()
    reads[6] = reads[6][:max(k - 1, 0)]
    reads[8] = ""
    reads[9] = ""
    seq = "".join(reads).encode()
    off = np.zeros(len(reads) + 1, dtype=np.uint64)
    np.cumsum([len(r) for r in reads], out=off[1:])
    n = len(seq)
    out = np.zeros(n, dtype=np.uint64)
    # two chunks, as hga_count_kmers launches them
    host_kernels.run_count_emit(seq, _p(off), C.c_uint64(len(reads)), k, C.c_uint64(0), C.c_uint64(n // 3), _p(out))
    tail = np.zeros(n - n // 3, dtype=np.uint64)
    host_kernels.run_count_emit(seq, _p(off), C.c_uint64(len(reads)), k, C.c_uint64(n // 3), C.c_uint64(n), _p(tail))
    out[n // 3:] = tail
    keys, cnt = np.unique(out, return_counts=True)                 # sort + run-length encode
    flag = np.zeros(len(keys), dtype=np.uint8)
    host_kernels.run_count_flag(_p(keys), _p(cnt.astype(np.uint32)), C.c_uint64(len(keys)), 2, _p(flag))
    wk, wc = exact_counts(seq, off, k, 2)
    assert np.array_equal(keys[flag.astype(bool)], wk) and np.array_equal(cnt[flag.astype(bool)].astype(np.uint32), wc)
    assert keys[-1] == np.uint64(0xFFFFFFFFFFFFFFFF) or k == 1     # windows over read ends / N produce the sentinel, which sorts last
    if k < 32:
        assert int(keys[:-1].max()) < (1 << (2 * k))               # ... also when only 2k bits are sorted


def test_second_merge_bodies_match_the_rule(host_kernels):
    """random purged index + unions + clusters: enr_merge2_keys_kernel and enr_purge2_kernel (count and fill) against the numpy rule
    that tests/test_second_merge_rule.py pins against the reference"""
    rng = np.random.default_rng(5)
    for trial in range(20):
        n_rows, n_slots, n_cores = 400, 300, int(rng.integers(3, 12))
        surv = np.sort(rng.choice(n_rows, n_cores, replace=False)).astype(np.uint32)          # survivor rows, ascending
        unions = [np.sort(rng.choice(n_slots, int(rng.integers(5, 120)), replace=False)) for _ in range(n_cores)]
        # purged lists: rows outside every core (here: any non-survivor row) and stale survivor copies, ascending, with duplicates
        lists = []
        for s in range(n_slots):
            e = rng.choice(n_rows, int(rng.integers(0, 12)))
            holders = [c for c in range(n_cores) if s in set(unions[c].tolist())]
            stale = [int(surv[c]) for c in holders for _ in range(int(rng.integers(0, 3)))]
            e = [int(v) for v in e if v not in set(surv.tolist())] + stale
            lists.append(np.sort(np.array(e, dtype=np.uint32)))
        p_off = np.zeros(n_slots + 1, dtype=np.uint32)
        np.cumsum([len(l) for l in lists], out=p_off[1:])
        p_row = np.concatenate(lists).astype(np.uint32) if p_off[-1] else np.zeros(1, np.uint32)
        # clusters over the cores: a few multi-member ones, singletons, and cores in no cluster
        perm = rng.permutation(n_cores)
        clusters, i = [], 0
        while i < n_cores - 1:
            sz = int(rng.integers(1, 4))
            clusters.append([int(surv[c]) for c in perm[i:i + sz]])
            i += sz
        # host glue as in hga_enrich_run: into / multi / new numbering / map / flags
        index_of = {int(s): c for c, s in enumerate(surv)}
        into = list(range(n_cores)); multi = [0] * n_cores
        for cl in clusters:
            if len(cl) < 2:
                continue
            for s in cl:
                into[index_of[s]] = index_of[cl[0]]; multi[index_of[s]] = 1
        new_idx, surv_new = [0] * n_cores, []
        for c in range(n_cores):
            if into[c] == c:
                new_idx[c] = len(surv_new); surv_new.append(int(surv[c]))
        cmap = np.array([new_idx[into[c]] | (multi[c] << 31) for c in range(n_cores)], dtype=np.uint32)
        once = np.zeros(n_rows + 1, dtype=np.uint8)
        for c in range(n_cores):
            if multi[c]:
                once[surv[c]] = 1
        ukeys = np.concatenate([(np.uint64(c) << np.uint64(32)) | unions[c].astype(np.uint64) for c in range(n_cores)])
        R2 = np.zeros(n_slots + 1, dtype=np.uint32)
        rel = np.zeros(len(ukeys), dtype=np.uint64)
        host_kernels.run_merge2_keys(_p(ukeys), C.c_uint64(len(ukeys)), _p(cmap), _p(surv), _p(np.array(surv_new + [0], dtype=np.uint32)), _p(R2), _p(rel))
        new_keys = np.unique(rel)                                                             # sort + unique
        cnt = np.zeros(n_slots + 1, dtype=np.uint32)
        host_kernels.run_purge2_count(_p(p_off), _p(p_row), n_slots, _p(R2), _p(once), _p(cnt))
        out_off = np.zeros(n_slots + 1, dtype=np.uint32)
        np.cumsum(cnt[:n_slots], out=out_off[1:])
        out_row = np.zeros(max(int(out_off[-1]), 1), dtype=np.uint32)
        host_kernels.run_purge2_fill(_p(p_off), _p(p_row), n_slots, _p(R2), _p(once), _p(out_off), _p(out_row))
        # the rule (rows as ids: first_id = 0)
        ids, want_unions, want_off, want_rows = second_merge_rule(surv.astype(np.int64), [u.astype(np.int64) for u in unions], p_off.astype(np.uint64),
                                                                  p_row[:p_off[-1]], clusters)
        assert ids.tolist() == surv_new
        assert np.array_equal(out_off.astype(np.uint64), want_off) and np.array_equal(out_row[:out_off[-1]], want_rows)
        got_unions = [np.sort((new_keys[(new_keys >> np.uint64(32)) == np.uint64(c)] & np.uint64(0xFFFFFFFF)).astype(np.int64)) for c in range(len(surv_new))]
        assert all(np.array_equal(a, b) for a, b in zip(got_unions, want_unions))
