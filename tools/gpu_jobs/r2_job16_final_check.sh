# round 2, GPU job 16 (1 GPU): the committed final state -- smoke(), full GPU tier, memcheck over the small-scalar MSM and the Lagrange-basis prover test
set -x
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/j16_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/j16_smoke.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/j16_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/j16_pytest_gpu.log
tail -4 gpurun_out/j16_pytest_gpu.log
timeout 420 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_prover.py -m gpu -x -q -k "small_scalar and 377 and not 200000 or lagrange" > gpurun_out/j16_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/j16_memcheck.log
tail -6 gpurun_out/j16_memcheck.log
