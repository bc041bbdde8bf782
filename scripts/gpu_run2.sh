# usage: bash scripts/gpu_run2.sh <tag> [ncu]
TAG=${1:-x}
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25) > gpurun_out/pytest_$TAG.log 2>&1
tail -4 gpurun_out/pytest_$TAG.log
timeout 600 python bench.py --genome-mbp 20 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench20_$TAG.log 2>&1; tail -2 gpurun_out/bench20_$TAG.log
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench100_$TAG.log 2>&1; tail -2 gpurun_out/bench100_$TAG.log
if [ "$2" = "ncu" ]; then
  timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/plain_$TAG.log 2>&1 && \
  timeout 1200 ncu --set full --clock-control none --import-source on -k 'regex:scan_probe_kernel|pair_count_kernel' -c 3 -o gpurun_out/prof_$TAG -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_$TAG.log 2>&1
  tail -3 gpurun_out/ncu_$TAG.log
fi
