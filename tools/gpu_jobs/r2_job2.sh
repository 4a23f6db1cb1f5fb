# round 2, GPU job 2: full GPU test tier, default bench line (4 KiB) with the new reference pairing, ncu launch list + --set full capture
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/j2_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/j2_pytest_gpu.log
python bench.py > gpurun_out/r2_bench_4k.json 2> gpurun_out/r2_bench_4k.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err; echo "ref rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_msm_accumulate -s 2 -c 1 -f -o gpurun_out/r2_msm_acc26_full python tools/quick_perf.py 26 > gpurun_out/j2_ncu_full.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_bench_4k.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/j2_ncu_launches.log 2>&1
ZKAES_TRACE=1 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/j2_trace_bench.json 2> gpurun_out/r2_phase_trace_4k.txt
tail -3 gpurun_out/j2_pytest_gpu.log; cat gpurun_out/r2_bench_4k.json; cat gpurun_out/r2_bench_reference.json; tail -20 gpurun_out/r2_phase_trace_4k.txt
