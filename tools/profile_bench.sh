set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -8
K='regex:spmv_|halo_pack|dot_kernel|pcg_|cheb_first|negate_copy|coarsest_kernel|carry_scalar|cg_p_'
python bench.py --steps 2 --no-cpu-baseline > gpurun_out/plain.log 2> gpurun_out/plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -s 2376 -c 800 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --no-cpu-baseline > gpurun_out/ncu1.log 2>&1
tail -3 gpurun_out/ncu1.log
python bench.py --steps 2 --no-cpu-baseline > gpurun_out/plain2.log 2> gpurun_out/plain2.err && \
ncu --set full --clock-control none --import-source on -k 'regex:spmv_sell_kernelILi3E' -s 2 -c 3 -o gpurun_out/prof_sell_cheb python bench.py --steps 2 --no-cpu-baseline > gpurun_out/ncu2.log 2>&1
tail -3 gpurun_out/ncu2.log
cat gpurun_out/plain.log | cut -c1-400
