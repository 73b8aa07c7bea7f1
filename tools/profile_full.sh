# one `ncu --set full` capture of the dominant kernel (fused Chebyshev sweep, sliced layout):
# launches 0,1 = level 0, 2,3 = level 1, 4,5 = level 2 of the first V-cycle
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -6
ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k 'regex:spmv_sell_kernel<\(int\)3>' -c 6 -o gpurun_out/r01_sell_cheb python tools/profile_kernels.py 256 > gpurun_out/ncu_full.log 2>&1
tail -5 gpurun_out/ncu_full.log
ls -la gpurun_out/
SAENA_BENCH_VERBOSE=1 python bench.py 2> gpurun_out/bench256.err | tee gpurun_out/bench256.json | cut -c1-300
grep -E "aggregation|RAP|setup" gpurun_out/bench256.err
