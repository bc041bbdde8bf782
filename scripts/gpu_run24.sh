TAG=${1:-x}
mkdir -p gpurun_out
HGA_ENRICH_TIMING=1 timeout 900 python scripts/cli_e2e.py --genome-mbp 10 --no-reference --out gpurun_out/cli_e2e_${TAG}.json 2>&1 | tail -1 | grep -o '"cli_stderr_tail.*' | tr ',' '\n' | cut -c1-200
