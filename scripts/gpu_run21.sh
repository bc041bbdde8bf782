TAG=${1:-x}
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5) > gpurun_out/pytest_${TAG}.log 2>&1; tail -2 gpurun_out/pytest_${TAG}.log | cut -c1-300
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/memcheck_${TAG}.log python -m pytest tests -m gpu -x -q -k "kat2 or config1_like or enrichment or empty_inputs or tetraploid or pivot_subset or heavy_rows" > gpurun_out/memcheck_pytest_${TAG}.log 2>&1; echo "memcheck rc=$?"
tail -3 gpurun_out/memcheck_pytest_${TAG}.log | cut -c1-300; tail -5 gpurun_out/memcheck_${TAG}.log | cut -c1-300
