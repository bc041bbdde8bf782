TAG=${1:-x}; N=${2:-2}
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25) > gpurun_out/pytest_$TAG.log 2>&1; tail -3 gpurun_out/pytest_$TAG.log
run() { NAME=$1; shift; env "$@" timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu-baseline $GENOME > gpurun_out/bench_${TAG}_$NAME.log 2>&1
  echo "$NAME: $(grep -o '"stages_ms": {[^}]*}' gpurun_out/bench_${TAG}_$NAME.log) $(grep -o '"e2e": {[^}]*}' gpurun_out/bench_${TAG}_$NAME.log)"; }
GENOME="--genome-mbp 20"; run g20 A=1
GENOME=""; run g100 A=1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tests/multi_gpu_parity.py > gpurun_out/multi_parity_${TAG}.log 2>&1; echo "parity rc=$?"; grep -n "parity ok\|Error\|differ" gpurun_out/multi_parity_${TAG}.log | head
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 3 --warmup 2 > gpurun_out/bench100_${TAG}_n$N.log 2>&1; echo "bench rc=$?"; grep -o '"value": [0-9.]*\|"pairs": [0-9.]*\|"stages_ms": {[^}]*}\|"exchange_ms": [0-9.]*\|"e2e": {[^}]*}' gpurun_out/bench100_${TAG}_n$N.log
