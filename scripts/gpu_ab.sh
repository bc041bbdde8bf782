# A/B helper: bench a few env variants under tight timeouts. usage: gpu_ab.sh TAG name1 "ENV=.. ENV=.." name2 "..." ...
TAG=$1; shift
mkdir -p gpurun_out
while [ $# -ge 2 ]; do
  NAME=$1; ENVS=$2; shift 2
  env $ENVS timeout 150 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --no-enrich > gpurun_out/bench_${TAG}_$NAME.log 2>&1
  echo "$NAME rc=$?: $(grep -o '"stages_ms": {[^}]*}' gpurun_out/bench_${TAG}_$NAME.log)"
done
