# round 2, GPU job 14 (1 GPU): the scratch arena after the private pool / key-size / trace changes -- golden + arena tests, an untraced bench with ZKAES_ALLOC_STATS=1
# (is the arena active at 4 KiB next to the 12.9 GB of Lagrange points?), and a traced one
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_prover.py -m gpu -x -q -k "golden or arena or lagrange" > gpurun_out/j14_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/j14_pytest.log
tail -3 gpurun_out/j14_pytest.log
ZKAES_ALLOC_STATS=1 timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/j14_bench_4k.json 2> gpurun_out/j14_bench_4k.err; echo "bench rc=$?"
grep "zkaes" gpurun_out/j14_bench_4k.err | cut -c1-200
python -c "
import json
d=json.loads(open('gpurun_out/j14_bench_4k.json').read()); print(d['ms_per_step'], d['config']['verified'], d['config']['proof_sha256'][:16], d['roofline']['avg_launch_ms'], d['roofline']['launches_per_step'])"
ZKAES_TRACE=1 timeout 900 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/j14_trace_bench.json 2> gpurun_out/j14_phase_trace_4k.txt
tail -17 gpurun_out/j14_phase_trace_4k.txt | cut -c1-175
