# N-rank bench of the final build with the single-GPU comparison on rank 0 (parity_n) and sub-stage traces
TAG=${1:-r3n}; N=${2:-2}
mkdir -p gpurun_out
env ${HGA_TRACE_ON:+HGA_TRACE=1} timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_${TAG}_n$N.log 2>&1; echo "bench rc=$?"
grep -o '"value": [0-9.]*\|"stages_ms": {[^}]*}\|"exchange_ms": [0-9.]*\|"parity_n": "[A-Za-z]*"\|"e2e": {[^}]*}' gpurun_out/bench_${TAG}_n$N.log
grep "hga trace r0" gpurun_out/bench_${TAG}_n$N.log | tail -4
