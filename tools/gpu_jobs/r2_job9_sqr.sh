# round 2, GPU job 9 (1 GPU): dedicated Montgomery squaring in the mixed addition -- parity, then timing
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/j9_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/j9_pytest_gpu.log
tail -4 gpurun_out/j9_pytest_gpu.log
timeout 600 python tools/quick_perf.py 22,26 > gpurun_out/r2_quick_perf_sqr.txt 2>&1
cat gpurun_out/r2_quick_perf_sqr.txt
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/j9_bench_4k.json 2> gpurun_out/j9_bench_4k.err
python -c "
import json
d=json.loads(open('gpurun_out/j9_bench_4k.json').read()); print(d['ms_per_step'], d['config']['verified'], d['config']['proof_sha256'][:16], d['roofline']['avg_launch_ms'], d['roofline']['alu'])
"
