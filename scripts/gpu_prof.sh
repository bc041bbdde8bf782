# single GPU: tests, bench 100, ncu full on scan + pair tier-1 kernels (100 Mbp)
TAG=${1:-x}
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25) > gpurun_out/pytest_$TAG.log 2>&1; tail -3 gpurun_out/pytest_$TAG.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench100_$TAG.log 2>&1; tail -1 gpurun_out/bench100_$TAG.log | cut -c1-3000
timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/plain_$TAG.log 2>&1 && \
timeout 1500 ncu --set full --clock-control none --import-source on -k 'regex:scan_probe_kernel|pair_count_warp_kernel' -c 3 -o gpurun_out/prof_$TAG -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_$TAG.log 2>&1
tail -2 gpurun_out/ncu_$TAG.log | cut -c1-200
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:scan_|pair_count|cc_|table_|hist_from|count_chunks|chunk_selected|write_selected|write_ties|count_le|expand_rows|run_offsets|increments|mark_pivots|RadixSort|DeviceScan|slots_to_kids|split_keys|add_u32' -c 400 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launch_$TAG.log 2>&1
wc -l gpurun_out/launches_$TAG.csv
