# quick: tests + 20 Mbp bench with diag variants
TAG=${1:-x}
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25) > gpurun_out/pytest_$TAG.log 2>&1
tail -4 gpurun_out/pytest_$TAG.log
for D in 0 1 2 3; do
  HGA_SCAN_DIAG=$D timeout 600 python bench.py --genome-mbp 20 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/bench20_${TAG}_d$D.log 2>&1
  python - <<PY
import json
l=[x for x in open("gpurun_out/bench20_${TAG}_d$D.log") if x.startswith("{")]
if l:
    j=json.loads(l[-1]); print("diag $D", j["stages_ms"], j.get("diagnostics"))
else:
    print(open("gpurun_out/bench20_${TAG}_d$D.log").read()[-2000:])
PY
done
HGA_L2_PERSIST=0 timeout 600 python bench.py --genome-mbp 20 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/bench20_${TAG}_nopersist.log 2>&1; tail -1 gpurun_out/bench20_${TAG}_nopersist.log | cut -c1-100; grep -o '"stages_ms": {[^}]*}' gpurun_out/bench20_${TAG}_nopersist.log
timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/bench100_$TAG.log 2>&1; grep -o '"stages_ms": {[^}]*}' gpurun_out/bench100_$TAG.log;  grep -o '"diagnostics": {[^}]*}' gpurun_out/bench100_$TAG.log
HGA_L2_PERSIST=0 timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/bench100_${TAG}_nopersist.log 2>&1; grep -o '"stages_ms": {[^}]*}' gpurun_out/bench100_${TAG}_nopersist.log
