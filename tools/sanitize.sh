# compute-sanitizer passes over the parity tests (one GPU; under gpurun --timeout 1500):
#   bash tools/sanitize.sh            memcheck on the golden parity tests + racecheck / synccheck on the smoke
# Output: gpurun_out/sanitize_*.log; "ERROR SUMMARY: 0 errors" is the pass criterion of each log.
# (Never a source of timings: the sanitizer serialises and instruments every kernel.)
set -x
export SAENA_B200_GRAPHS=${SAENA_B200_GRAPHS:-1}
CS=/usr/local/cuda/bin/compute-sanitizer
mkdir -p gpurun_out
timeout 400 $CS --tool memcheck --leak-check no --print-limit 20 --log-file gpurun_out/sanitize_memcheck.log \
  python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "golden or edge or empty or ragged" > gpurun_out/sanitize_memcheck.out 2>&1
tail -3 gpurun_out/sanitize_memcheck.out; grep -E "ERROR SUMMARY|Invalid|out of bounds" gpurun_out/sanitize_memcheck.log | head
for tool in racecheck synccheck initcheck; do
  timeout 200 $CS --tool $tool --print-limit 20 --log-file gpurun_out/sanitize_$tool.log \
    python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitize_$tool.out 2>&1
  tail -1 gpurun_out/sanitize_$tool.out; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard" gpurun_out/sanitize_$tool.log | head -5
done
