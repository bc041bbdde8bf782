# single GPU: smoke, tests, default bench, ncu full on the scan + pair kernels, launch list of one step (+ hga_enrich)
TAG=${1:-x}
mkdir -p gpurun_out
(timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2) > gpurun_out/smoke_$TAG.log 2>&1; tail -1 gpurun_out/smoke_$TAG.log | cut -c1-300
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25) > gpurun_out/pytest_$TAG.log 2>&1; tail -2 gpurun_out/pytest_$TAG.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench100_$TAG.log 2>&1; tail -1 gpurun_out/bench100_$TAG.log | cut -c1-600
timeout 900 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref_$TAG.log 2>&1; tail -1 gpurun_out/bench_ref_$TAG.log | cut -c1-400
timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/plain_$TAG.log 2>&1 && \
timeout 1500 ncu --set full --clock-control none --import-source on -k 'regex:scan_probe_kernel|pair_count_warp_kernel' -c 3 -o gpurun_out/prof_$TAG -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-enrich > gpurun_out/ncu_$TAG.log 2>&1
tail -2 gpurun_out/ncu_$TAG.log | cut -c1-200
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:scan_|pair_count|cc_|enr_|table_|hist_from|count_chunks|chunk_selected|write_selected|write_ties|count_le|expand_rows|run_offsets|increments|mark_pivots|RadixSort|DeviceScan|DeviceSelect|DeviceRunLength|DeviceReduce|slots_to_kids|split_keys|add_u32|low32|kid_list' -c 600 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launch_$TAG.log 2>&1
wc -l gpurun_out/launches_$TAG.csv
