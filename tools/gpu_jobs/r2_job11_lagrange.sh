# round 2, GPU job 11 (1 GPU): private memory pool per context + Lagrange-basis round-1 commitments (small-scalar MSM):
# full GPU tier, 4 KiB bench (with and without the Lagrange path), phase trace
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/j11_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/j11_pytest_gpu.log
tail -5 gpurun_out/j11_pytest_gpu.log
ZKAES_TRACE=1 timeout 900 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/j11_trace_bench.json 2> gpurun_out/j11_phase_trace_4k.txt
grep -E "keys:|r1:" gpurun_out/j11_phase_trace_4k.txt | tail -12
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/j11_bench_4k.json 2> gpurun_out/j11_bench_4k.err
ZKAES_LAGRANGE=0 timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/j11_bench_4k_nolag.json 2> gpurun_out/j11_bench_4k_nolag.err
python -c "
import json
for n in ('j11_bench_4k','j11_bench_4k_nolag'):
    d=json.loads(open('gpurun_out/%s.json'%n).read()); print(n, d['ms_per_step'], d['config']['verified'], d['config']['proof_sha256'][:16], d['roofline']['avg_launch_ms'], d['roofline'].get('launches_per_step'))
"
tail -3 gpurun_out/j11_bench_4k.err
