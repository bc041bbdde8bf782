# r3k: L2 sectors / DRAM bytes of the scan kernel for four forms of the key-sector load
TAG=${1:-r3k}
mkdir -p gpurun_out
B="--steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-enrich"
M="dram__bytes_read.sum,gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum"
for V in 4 8 16; do
  env HGA_SCAN_DIAG=$V timeout 600 ncu --metrics $M --clock-control none -k 'regex:scan_probe_kernel' -s 1 -c 1 --csv --log-file gpurun_out/dram_${TAG}_v$V.csv python bench.py $B > gpurun_out/dram_${TAG}_v$V.log 2>&1
  echo "$V rc=$?"
done
