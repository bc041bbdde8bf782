# First GPU call of round 2: everything that was written after round 1's GPU minutes were spent, then the usual numbers.
# usage (from the repo root, through gpurun):   gpurun --timeout 1500 -- 'bash scripts/gpu_round2_first.sh r2a'
TAG=${1:-r2a}
mkdir -p gpurun_out
# 1. the blind code: hga_enrich_full (tail / spectral block, second merge), --spectral, hga_count_kmers, jf_occurrences
timeout 600 python -m pytest tests/test_zx_gpu_forced_spectral.py tests/test_zy_gpu_tail_block.py tests/test_zz_gpu_sdk_selection.py -m gpu -q > gpurun_out/zz_${TAG}.log 2>&1
echo "zz tests rc=$?: $(tail -1 gpurun_out/zz_${TAG}.log)"
# 2. phase times of the block on a case with many scaffold components (HGA_ENRICH_TIMING prints every phase on stderr)
HGA_ENRICH_TIMING=1 timeout 300 python - > gpurun_out/tail_block_${TAG}.log 2>&1 <<'PY'
import sys, time
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import numpy as np, golden_util, hga_b200
c = golden_util.load_case("full_short")
with hga_b200.Handle(c["kmers"], c["k"]) as h:
    h.scan(c["bases"], c["seq_off"]); h.build_index(); h.pair_count(min_score=1); h.select_edges(fraction=c["fraction"])
    for _ in range(2):
        t = time.perf_counter()
        h.enrich_full(c["seq_off"], min_size=c["min_size"], enrichment_min_score=c["enrich"])
        print("hga_enrich_full wall ms", (time.perf_counter() - t) * 1e3, h.metrics()["n_cores"], "cores")
PY
echo "tail block timing rc=$?"
# 3. the whole GPU suite and the default bench line
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/gpu_tests_${TAG}.log 2>&1
echo "gpu tests rc=$?: $(tail -1 gpurun_out/gpu_tests_${TAG}.log)"
timeout 600 python bench.py --gpus 1 --steps 3 --warmup 3 > gpurun_out/bench_${TAG}.log 2>&1
echo "bench rc=$?: $(tail -c 600 gpurun_out/bench_${TAG}.log)"
