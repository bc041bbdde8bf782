TAG=${1:-x}
mkdir -p gpurun_out
run() { NAME=$1; shift; env "$@" timeout 150 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-enrich > gpurun_out/bench_${TAG}_$NAME.log 2>&1
  echo "$NAME rc=$?: $(grep -o '"scan": [0-9.]*' gpurun_out/bench_${TAG}_$NAME.log) $(grep -o '"pair_count": [0-9.]*' gpurun_out/bench_${TAG}_$NAME.log)"; }
run pad0 HGA_SCAN_PAD_KB=0
run pad7 HGA_SCAN_PAD_KB=7
run pad14 HGA_SCAN_PAD_KB=14
run pad26 HGA_SCAN_PAD_KB=26
run pad45 HGA_SCAN_PAD_KB=45
