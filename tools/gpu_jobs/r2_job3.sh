# round 2, GPU job 3 (1 GPU): GPU tier after the AllocatedBool::or change / key files / coset waves; MSM window cap 22 vs 23 at 4 KiB
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/j3_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/j3_pytest_gpu.log
tail -3 gpurun_out/j3_pytest_gpu.log
ZKAES_MSM_WINDOW_MAX=22 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/j3_bench_c22.json 2> gpurun_out/j3_bench_c22.err
ZKAES_MSM_WINDOW_MAX=23 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/j3_bench_c23.json 2> gpurun_out/j3_bench_c23.err
python -c "
import json
for n in ('c22','c23'):
    d=json.loads(open('gpurun_out/j3_bench_%s.json'%n).read()); print(n, d['ms_per_step'], d['config']['verified'], d['config']['proof_sha256'][:16], d['roofline']['avg_launch_ms'], d['roofline']['launches_per_step'])
"
tail -3 gpurun_out/j3_bench_c23.err
