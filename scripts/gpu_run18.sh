TAG=${1:-x}
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q -k "enrich or final or run_clustering" 2>&1 | tail -30) > gpurun_out/pytest_${TAG}.log 2>&1; tail -3 gpurun_out/pytest_${TAG}.log | cut -c1-300
HGA_ENRICH_TIMING=1 timeout 900 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/bench100_${TAG}.log 2>&1; echo "bench rc=$?"
grep "hga_enrich:" gpurun_out/bench100_${TAG}.log | tail -6; grep -o '"enrich": {[^}]*}' gpurun_out/bench100_${TAG}.log
