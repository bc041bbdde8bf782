# r3x: NCCL p2p channel count against the all-to-all times of the index and partial-pair exchanges (traces), N = $2
TAG=${1:-r3x}; N=${2:-4}
mkdir -p gpurun_out
for V in default ch16 ch32; do
  E=""; [ $V = ch16 ] && E="NCCL_MIN_P2P_NCHANNELS=16 NCCL_MAX_P2P_NCHANNELS=32"; [ $V = ch32 ] && E="NCCL_MIN_P2P_NCHANNELS=32 NCCL_MAX_P2P_NCHANNELS=32"
  env $E HGA_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 2 --warmup 2 --no-e2e --no-parity > gpurun_out/bench_${TAG}_${V}_n$N.log 2>&1
  echo "$V rc=$? $(grep -o '"value": [0-9.]*' gpurun_out/bench_${TAG}_${V}_n$N.log | head -1)"
  grep "hga trace r0 index\|hga trace r0 partials\] pack" gpurun_out/bench_${TAG}_${V}_n$N.log | tail -2
done
