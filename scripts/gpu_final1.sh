TAG=${1:-x}
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -30) > gpurun_out/pytest_${TAG}.log 2>&1; tail -3 gpurun_out/pytest_${TAG}.log | cut -c1-300
run() { NAME=$1; shift; env "$@" timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e $GENOME > gpurun_out/bench_${TAG}_$NAME.log 2>&1
  echo "$NAME: $(grep -o '"scan": [0-9.]*' gpurun_out/bench_${TAG}_$NAME.log)"; }
GENOME="--genome-mbp 20"; run g20_d0 HGA_SCAN_DIAG=0; run g20_d1 HGA_SCAN_DIAG=1; run g20_d2 HGA_SCAN_DIAG=2
GENOME=""; run g100_d0 HGA_SCAN_DIAG=0; run g100_d1 HGA_SCAN_DIAG=1; run g100_d2 HGA_SCAN_DIAG=2
bash scripts/gpu_prof.sh $TAG 2>&1 | tail -5 | cut -c1-2500
