TAG=${1:-r2o}
mkdir -p gpurun_out
timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/plain_$TAG.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:scan_|pair_count|cc_|enr_|table_|hist_from|count_chunks|chunk_selected|write_selected|write_ties|count_le|expand_rows|run_offsets|increments|row_minhash|mark_pivots|flag_min|RadixSort|DeviceScan|DeviceSelect|DeviceRunLength|DeviceReduce|split_keys|add_u32|low32|kid_list' -c 700 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launch_$TAG.log 2>&1
wc -l gpurun_out/launches_$TAG.csv
