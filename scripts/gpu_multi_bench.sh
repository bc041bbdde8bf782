TAG=${1:-x}; N=${2:-8}
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 3 --warmup 2 --no-e2e > gpurun_out/bench100_${TAG}_n$N.log 2>&1; echo "bench rc=$?"; grep -o '"value": [0-9.]*\|"stages_ms": {[^}]*}\|"exchange_ms": [0-9.]*' gpurun_out/bench100_${TAG}_n$N.log
