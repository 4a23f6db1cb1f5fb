# round 2, GPU job 13 (1 GPU): final state -- full GPU tier, the default bench line (with the CPU baseline), phase trace,
# ncu launch list of the bench command, ncu --set full capture of one bucket-accumulation launch (MSM 2^26)
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/j13_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/j13_pytest_gpu.log
tail -4 gpurun_out/j13_pytest_gpu.log
timeout 1200 python bench.py > gpurun_out/j13_bench_4k_default.json 2> gpurun_out/j13_bench_4k_default.err; echo "bench rc=$?"
cut -c1-600 gpurun_out/j13_bench_4k_default.json
ZKAES_TRACE=1 timeout 900 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/j13_trace_bench.json 2> gpurun_out/j13_phase_trace_4k.txt
tail -22 gpurun_out/j13_phase_trace_4k.txt | cut -c1-90
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_msm_accumulate -s 2 -c 1 -f -o gpurun_out/j13_msm_acc26_full python tools/quick_perf.py 26 > gpurun_out/j13_ncu_full.log 2>&1
python tools/ncu_summary.py full gpurun_out/j13_msm_acc26_full.ncu-rep > gpurun_out/j13_ncu_full_summary.txt 2>&1; head -30 gpurun_out/j13_ncu_full_summary.txt
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/j13_launches_bench_4k.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/j13_ncu_launches.log 2>&1
python tools/ncu_summary.py launches gpurun_out/j13_launches_bench_4k.csv > gpurun_out/j13_launches_summary.txt 2>&1; head -24 gpurun_out/j13_launches_summary.txt
