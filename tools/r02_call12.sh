# Round 2, last call: the driver's own GPU test command at HEAD, then smoke()
mkdir -p gpurun_out
set -x
timeout 170 python -m pytest tests -x -q -m gpu > gpurun_out/r02h_pytest.log 2>&1; tail -3 gpurun_out/r02h_pytest.log | cut -c1-300
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
