mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/pytest_r1b.log 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi_r1b.log
(time timeout 600 python bench.py --genome-mbp 20 --steps 3 --warmup 3 --no-cpu-baseline) > gpurun_out/bench20_r1b.log 2>&1
(time timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline) > gpurun_out/bench100_r1b.log 2>&1
tail -3 gpurun_out/pytest_r1b.log; tail -5 gpurun_out/bench20_r1b.log; tail -5 gpurun_out/bench100_r1b.log
