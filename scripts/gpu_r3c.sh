# r3c: launch list of the library's kernels (two steps), after a plain run
TAG=${1:-r3c}
mkdir -p gpurun_out
B="--steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-enrich"
timeout 600 python bench.py $B > gpurun_out/plain_$TAG.log 2>&1 || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:scan_|pair_count|cc_|enr_|table_|hist_from|count_chunks|chunk_selected|write_selected|write_ties|count_le|expand_rows|run_offsets|group_offsets|index_local|increments|row_minhash|row_label|row_neighbor|mark_pivots|flag_min|CUB_|split_keys|add_u32|low32|kid_list' -c 400 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py $B > gpurun_out/ncu_launch_$TAG.log 2>&1
echo "launch list rc=$? $(wc -l < gpurun_out/launches_$TAG.csv)"
