# ncu --set full of the scan kernel (20 Mbp diploid = 2 Gbases): full, no key probes (diag 1), no filter probes (diag 2)
TAG=${1:-r2b}
mkdir -p gpurun_out
B="--steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-enrich --genome-mbp 20"
for V in 0 1 2; do
  HGA_SCAN_DIAG=$V timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_probe_kernel -s 3 -c 1 -o gpurun_out/scan_${TAG}_d$V -f python bench.py $B > gpurun_out/ncu_${TAG}_d$V.log 2>&1
  echo "d$V rc=$? $(ls -la gpurun_out/scan_${TAG}_d$V.ncu-rep | awk '{print $5}')"
done
