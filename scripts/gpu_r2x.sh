# r2x (2 GPUs): N-rank parity incl. the gathered handle, the CLI at --gpus 2, traced bench
TAG=${1:-r2x}; N=${2:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -x -q > gpurun_out/pytest_multi_$TAG.log 2>&1; echo "pytest multi rc=$? $(tail -1 gpurun_out/pytest_multi_$TAG.log)"
tail -30 gpurun_out/pytest_multi_$TAG.log | grep -n "Error\|assert\|differ\|error" | head
HGA_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 2 --warmup 2 --no-e2e > gpurun_out/bench_${TAG}_n$N.log 2>&1; echo "bench rc=$?"
grep -o '"value": [0-9.]*\|"stages_ms": {[^}]*}\|"exchange_ms": [0-9.]*\|"parity_n": "[A-Za-z]*"' gpurun_out/bench_${TAG}_n$N.log
grep "hga trace r0" gpurun_out/bench_${TAG}_n$N.log | tail -5
HGA_PARTIALS_UNPACKED=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus $N --steps 1 --warmup 1 --no-e2e > gpurun_out/bench_${TAG}_unpacked_n$N.log 2>&1; echo "bench (unpacked partials) rc=$?"
grep -o '"stages_ms": {[^}]*}\|"parity_n": "[A-Za-z]*"' gpurun_out/bench_${TAG}_unpacked_n$N.log
