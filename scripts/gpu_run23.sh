TAG=${1:-x}
mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -x -q -k "sc_score or cli or score_threshold" 2>&1 | tail -5) > gpurun_out/pytest_${TAG}.log 2>&1; tail -2 gpurun_out/pytest_${TAG}.log | cut -c1-300
timeout 900 python scripts/cli_e2e.py --genome-mbp 10 --no-reference --out gpurun_out/cli_e2e_${TAG}.json 2>&1 | tail -1 | cut -c1-1800
