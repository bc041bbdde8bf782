# final check of a build on one GPU: the whole GPU suite, smoke(), the default bench line
TAG=${1:-final}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/pytest_$TAG.log)"
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$? $(tail -1 gpurun_out/smoke_$TAG.log)"
timeout 900 python bench.py > gpurun_out/bench_$TAG.log 2>gpurun_out/bench_$TAG.err; echo "bench rc=$?"
grep -o '"value": [0-9.]*\|"stages_ms": {[^}]*}\|"frac": [0-9.]*\|"e2e": {[^}]*}\|"gpu_launches": [0-9]*\|"clocks": {[^}]*}' gpurun_out/bench_$TAG.log | head -8
