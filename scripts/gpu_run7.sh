TAG=${1:-x}
mkdir -p gpurun_out
run() { NAME=$1; shift; env "$@" timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e $GENOME > gpurun_out/bench_${TAG}_$NAME.log 2>&1
  echo "$NAME: $(grep -o '"stages_ms": {[^}]*}' gpurun_out/bench_${TAG}_$NAME.log)"; }
GENOME="--genome-mbp 20"; run g20_occ6 HGA_SCAN_MIN_CTAS=6; run g20_occ7 HGA_SCAN_MIN_CTAS=7
GENOME=""; run g100_occ6 HGA_SCAN_MIN_CTAS=6; run g100_occ7 HGA_SCAN_MIN_CTAS=7
bash scripts/gpu_prof.sh $TAG 2>&1 | tail -4 | cut -c1-300
