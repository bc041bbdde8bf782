# usage: bash scripts/gpu_multi.sh <tag> <ngpus>
TAG=${1:-x}; N=${2:-2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tests/multi_gpu_parity.py > gpurun_out/multi_parity_${TAG}.log 2>&1; echo "parity rc=$?"; tail -25 gpurun_out/multi_parity_${TAG}.log | cut -c1-400
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 3 --warmup 2 --no-e2e > gpurun_out/bench100_${TAG}_n$N.log 2>&1; echo "bench rc=$?"; tail -5 gpurun_out/bench100_${TAG}_n$N.log | cut -c1-1500
