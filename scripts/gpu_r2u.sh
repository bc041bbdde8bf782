# r2u: N-rank bench with HGA_TRACE sub-stage times (no parity leg: quick), N = $2
TAG=${1:-r2u}; N=${2:-2}
mkdir -p gpurun_out
HGA_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 2 --warmup 2 --no-e2e --no-parity > gpurun_out/bench_${TAG}_n$N.log 2>&1; echo "bench rc=$?"
grep -o '"value": [0-9.]*\|"stages_ms": {[^}]*}\|"exchange_ms": [0-9.]*' gpurun_out/bench_${TAG}_n$N.log
grep "hga trace r0" gpurun_out/bench_${TAG}_n$N.log | tail -8
