TAG=${1:-x}
mkdir -p gpurun_out
timeout 900 python scripts/cli_e2e.py --genome-mbp 4 --no-reference --out gpurun_out/cli_e2e_${TAG}.json 2>&1 | tail -3 | cut -c1-2500
