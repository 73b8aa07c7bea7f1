# Round 2: the whole -m gpu suite WITHOUT -x (every failure in one call), one GPU
mkdir -p gpurun_out
set -x
timeout 1200 python -m pytest tests -q -m gpu -s > gpurun_out/r02e_pytest.log 2>&1; tail -6 gpurun_out/r02e_pytest.log | cut -c1-300; grep -E "^n=|^FAILED|^ERROR|^E  " gpurun_out/r02e_pytest.log | cut -c1-300 | head -30
