# `ncu --set full` of the transfer kernels that sit under 70 % of the HBM peak in profiles/r01_bench_levels.md:
# P and R of level 2 (789 331 x 120 107, 62.4 M entries: P on 16 lanes per row -- spmv_vec_kernel<16, ...>, the only
# kernel ptxas spills on -- R on 128 threads per row).  One GPU, after the same command has exited 0 without ncu.
#   bash tools/profile_transfers.sh            -> gpurun_out/r02_transfers.ncu-rep + r02_transfers_raw.csv
# EPI_PLAIN = 0 (R), EPI_SUB = 5 (P fused with the correction).
set -x
python tools/profile_kernels.py 256 > gpurun_out/profile_kernels_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k 'regex:spmv_(vec|rowgroup)_kernel<\(int\)(16|128), \(int\)(0|5)' -c 8 -o gpurun_out/r02_transfers \
    python tools/profile_kernels.py 256 > gpurun_out/ncu_transfers.log 2>&1
tail -3 gpurun_out/ncu_transfers.log
ncu -i gpurun_out/r02_transfers.ncu-rep --page raw --csv \
    --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct \
    > gpurun_out/r02_transfers_raw.csv 2>/dev/null
cut -c1-220 gpurun_out/r02_transfers_raw.csv | head -12
