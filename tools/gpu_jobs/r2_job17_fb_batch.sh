# round 2, GPU job 17 (1 GPU): batched normalisation of the fixed-base multiplications (SRS powers, Lagrange points) -- full GPU tier, key synthesis timing at 4 KiB
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/j17_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/j17_pytest_gpu.log
tail -4 gpurun_out/j17_pytest_gpu.log
ZKAES_TRACE=1 timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/j17_bench_4k.json 2> gpurun_out/j17_trace.txt; echo "bench rc=$?"
grep "keys:" gpurun_out/j17_trace.txt | cut -c1-120
python -c "
import json
d=json.loads(open('gpurun_out/j17_bench_4k.json').read()); print(d['ms_per_step'], d['config']['verified'], d['config']['proof_sha256'][:16], d['config']['key_setup_s'])"
