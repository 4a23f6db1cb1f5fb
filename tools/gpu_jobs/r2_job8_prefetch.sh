# round 2, GPU job 8 (1 GPU): staging the next point in shared memory (cp.async / cp.async.bulk + mbarrier) -- parity with the knob on, then timing
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_fq52.py -m gpu -x -q -k "tuning or fq52 or out_of_line" > gpurun_out/j8_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/j8_pytest.log
tail -4 gpurun_out/j8_pytest.log
timeout 600 python tools/quick_perf.py 22,26 0 "msm_prefetch=0/msm_prefetch=1/msm_prefetch=2/msm_prefetch=0" > gpurun_out/r2_quick_perf_prefetch.txt 2>&1
cat gpurun_out/r2_quick_perf_prefetch.txt
