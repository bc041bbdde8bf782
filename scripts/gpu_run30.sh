TAG=${1:-x}
mkdir -p gpurun_out
run() { NAME=$1; shift; env "$@" timeout 150 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-enrich > gpurun_out/bench_${TAG}_$NAME.log 2>&1
  echo "$NAME rc=$?: $(grep -o 'pad [0-9]* B -> [0-9]* CTAs' gpurun_out/bench_${TAG}_$NAME.log | head -1) $(grep -o '"pair_count": [0-9.]*' gpurun_out/bench_${TAG}_$NAME.log)"; }
run pad0 HGA_PAIR_PAD_KB=0
run pad12 HGA_PAIR_PAD_KB=12
run pad24 HGA_PAIR_PAD_KB=24
run pad42 HGA_PAIR_PAD_KB=42
run pad80 HGA_PAIR_PAD_KB=80
