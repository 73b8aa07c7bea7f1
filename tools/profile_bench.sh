# Round-1 evidence run (one B200): GPU tests, then the two ncu passes of B200_PROFILING.md.
# gpurun runs each ncu command once without ncu first (its own guard), so no explicit plain run here.
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
K='regex:spmv_|halo_pack|dot_kernel|pcg_|cheb_first|negate_copy|coarsest_kernel|carry_scalar|cg_p_|scale_vector|widen_ghost'
# 1. every launch of one timed solve with its device time (3 warm-up solves x ~792 launches skipped)
ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -s 2380 -c 800 --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 2 --no-cpu-baseline > gpurun_out/ncu1.log 2>&1
tail -2 gpurun_out/ncu1.log | cut -c1-300
# 2. the dominant kernel, full section set (launches 0,1 = level 0; 2,3 = level 1; 4,5 = level 2)
ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k 'regex:spmv_sell_kernel<\(int\)3>' -c 6 -o gpurun_out/r01_sell_cheb python tools/profile_kernels.py 256 > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
