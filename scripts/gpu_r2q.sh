# round-2 evidence for profiles/: ncu --set full of the final scan and pair kernels (10 Gbases), launch list of one step
TAG=${1:-r2q}
mkdir -p gpurun_out
B="--steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-enrich"
timeout 600 python bench.py $B > gpurun_out/plain_$TAG.log 2>&1 || exit 1
timeout 1200 ncu --set full --clock-control none --import-source on -k 'regex:scan_probe_kernel|pair_count_warp_kernel' -s 3 -c 2 -o gpurun_out/prof_$TAG -f python bench.py $B > gpurun_out/ncu_$TAG.log 2>&1
echo "ncu full rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:scan_|pair_count|cc_|enr_|table_|hist_from|count_chunks|chunk_selected|write_selected|write_ties|count_le|expand_rows|run_offsets|increments|row_minhash|mark_pivots|flag_min|RadixSort|DeviceScan|DeviceSelect|DeviceRunLength|DeviceReduce|split_keys|add_u32|low32|kid_list' -c 400 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py $B > gpurun_out/ncu_launch_$TAG.log 2>&1
echo "launch list rc=$? $(wc -l < gpurun_out/launches_$TAG.csv)"
