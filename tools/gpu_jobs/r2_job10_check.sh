# round 2, GPU job 10 (1 GPU): round 2 on three cosets -- golden-proof parity and timing; racecheck of the 16-byte prover kernels
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/j10_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/j10_pytest_gpu.log
tail -4 gpurun_out/j10_pytest_gpu.log
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/j10_bench_4k.json 2> gpurun_out/j10_bench_4k.err
python -c "
import json
d=json.loads(open('gpurun_out/j10_bench_4k.json').read()); print(d['ms_per_step'], d['config']['verified'], d['config']['proof_sha256'][:16], d['roofline']['avg_launch_ms'])
"
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_prover.py -m gpu -x -q -k "golden" > gpurun_out/j10_racecheck.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/j10_racecheck.log
tail -6 gpurun_out/j10_racecheck.log
