# r2v: deterministic table + slot-owner partition. 1 GPU: parity suite + bench; run with --gpus 2: N-rank parity + traced bench
TAG=${1:-r2v}; N=${2:-1}
mkdir -p gpurun_out
if [ "$N" = "1" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1
  echo "pytest rc=$? $(tail -1 gpurun_out/pytest_$TAG.log)"
  bash scripts/gpu_ab.sh $TAG base ""
else
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tests/multi_gpu_parity.py > gpurun_out/multi_parity_${TAG}_n$N.log 2>&1; echo "parity rc=$?"; grep -n "parity ok\|Error\|differ\|assert" gpurun_out/multi_parity_${TAG}_n$N.log | head
  HGA_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 2 --warmup 2 --no-e2e > gpurun_out/bench_${TAG}_n$N.log 2>&1; echo "bench rc=$?"
  grep -o '"value": [0-9.]*\|"stages_ms": {[^}]*}\|"exchange_ms": [0-9.]*\|"parity_n": "[A-Za-z]*"' gpurun_out/bench_${TAG}_n$N.log
  grep "hga trace r0" gpurun_out/bench_${TAG}_n$N.log | tail -4
fi
