set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/j1_smi.txt 2>&1
(cd tools && ./build/ubench) > gpurun_out/ubench_r2_raw.txt 2>&1
python tools/quick_perf.py 22,26 0 "msm_madd_call=0/msm_madd_call=1/msm_madd_call=1,msm_acc_blocks=4/msm_madd_call=1,msm_acc_blocks=3,msm_window_max=23/msm_window_max=22" > gpurun_out/j1_quick_perf.txt 2>&1
python -m pytest tests -m gpu -x -q > gpurun_out/j1_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/j1_pytest_gpu.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_prover.py tests/test_gpu_witness.py -m gpu -x -q -k "golden or fips197 or wrong_length or ragged or (assignment and 1)" > gpurun_out/j1_memcheck.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/j1_memcheck.log
tail -3 gpurun_out/j1_pytest_gpu.log; tail -5 gpurun_out/j1_memcheck.log; cat gpurun_out/j1_quick_perf.txt
