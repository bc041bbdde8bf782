TAG=${1:-r2l}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1
echo "gpu tests rc=$?: $(tail -3 gpurun_out/pytest_$TAG.log | tr '\n' ' ')"
(timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2) > gpurun_out/smoke_$TAG.log 2>&1; tail -1 gpurun_out/smoke_$TAG.log | cut -c1-300
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench100_$TAG.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench100_$TAG.log | cut -c1-1500
timeout 900 python bench.py --impl reference --steps 2 --warmup 0 > gpurun_out/bench_ref_$TAG.log 2>&1; tail -1 gpurun_out/bench_ref_$TAG.log | cut -c1-400
