# Evidence set of one build on one GPU: GPU tests, the default bench line, ncu --set full of the scan and pair kernels, launch list, the
# reference arm, the other BASELINE configs.  usage: gpu_evidence.sh TAG
TAG=${1:-r3f}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/pytest_$TAG.log)"
timeout 900 python bench.py > gpurun_out/bench_$TAG.log 2>gpurun_out/bench_$TAG.err; echo "bench rc=$?"
grep -o '"value": [0-9.]*\|"stages_ms": {[^}]*}\|"frac": [0-9.]*\|"e2e": {[^}]*}' gpurun_out/bench_$TAG.log | head -6
B="--steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-enrich"
timeout 1200 ncu --set full --clock-control none --import-source on -k 'regex:scan_probe_kernel|pair_count_warp_kernel' -s 2 -c 2 -o gpurun_out/prof_$TAG -f python bench.py $B > gpurun_out/ncu_$TAG.log 2>&1; echo "ncu full rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:scan_|pair_count|cc_|enr_|table_|hist_from|count_chunks|chunk_selected|write_selected|write_ties|count_le|expand_rows|run_offsets|group_offsets|index_local|increments|row_minhash|mark_pivots|flag_min|RadixSort|DeviceScan|DeviceSelect|DeviceRunLength|DeviceReduce|split_keys|add_u32|low32|kid_list' -c 400 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py $B > gpurun_out/ncu_launch_$TAG.log 2>&1; echo "launch list rc=$? $(wc -l < gpurun_out/launches_$TAG.csv)"
timeout 600 python bench.py --impl reference --steps 3 --warmup 0 > gpurun_out/bench_ref_$TAG.log 2>&1; echo "reference arm rc=$?"
bash scripts/gpu_configs.sh $TAG > gpurun_out/configs_$TAG.txt 2>&1; echo "configs rc=$?"
