# N-rank parity (tests/multi_gpu_parity.py) + bench with the single-GPU comparison on rank 0
TAG=${1:-x}; N=${2:-2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tests/multi_gpu_parity.py > gpurun_out/multi_parity_${TAG}_n$N.log 2>&1; echo "parity rc=$?"; grep -n "parity ok\|Error\|differ" gpurun_out/multi_parity_${TAG}_n$N.log | head
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 3 --warmup 2 --no-e2e > gpurun_out/bench100_${TAG}_n$N.log 2>&1; echo "bench rc=$?"; grep -o '"value": [0-9.]*\|"pairs": [0-9.]*\|"stages_ms": {[^}]*}\|"exchange_ms": [0-9.]*\|"parity_n": "[A-Za-z]*"' gpurun_out/bench100_${TAG}_n$N.log; tail -3 gpurun_out/bench100_${TAG}_n$N.log | cut -c1-300
