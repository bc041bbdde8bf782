TAG=${1:-x}
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -30) > gpurun_out/pytest_${TAG}.log 2>&1; tail -3 gpurun_out/pytest_${TAG}.log | cut -c1-300
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench100_${TAG}.log 2>&1; echo "bench rc=$?"
grep -o '"stages_ms": {[^}]*}' gpurun_out/bench100_${TAG}.log; grep -o '"enrich": {[^}]*}' gpurun_out/bench100_${TAG}.log; tail -3 gpurun_out/bench100_${TAG}.log | cut -c1-400
