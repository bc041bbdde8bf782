TAG=${1:-x}
mkdir -p gpurun_out
m() { NAME=$1; shift; env "$@" timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,lts__t_sectors.sum,lts__t_sectors_srcunit_tex.sum,lts__t_sectors_srcunit_tex_lookup_hit.sum,lts__t_sectors_srcunit_tex_lookup_miss.sum,gpu__time_duration.sum --clock-control none -k regex:scan_probe_kernel -s 1 -c 1 --csv --log-file gpurun_out/dram_${TAG}_$NAME.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > /dev/null 2>&1
echo "$NAME:"; grep scan_probe gpurun_out/dram_${TAG}_$NAME.csv | awk -F'","' '{print "   ", $(NF-2), $(NF-1), $NF}'; }
m d2 HGA_SCAN_DIAG=2
m d5 HGA_SCAN_DIAG=5
m d5_b4 HGA_SCAN_DIAG=5 HGA_FILTER_BITS_PER_KEY=4
m d0 HGA_SCAN_DIAG=0
