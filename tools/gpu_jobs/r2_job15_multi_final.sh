# round 2, GPU job 15 (N GPUs, N = $1): final multi-GPU state -- world-N tests, the 4 KiB proof untraced (the number) and traced (the phase split)
N=$1
set -x
mkdir -p gpurun_out
nvidia-smi -L | head -8
if [ "${2:-}" = "tests" ]; then
timeout -k 10 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -k "$N" > gpurun_out/j15_pytest_multi_${N}gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/j15_pytest_multi_${N}gpu.log
tail -4 gpurun_out/j15_pytest_multi_${N}gpu.log
fi
timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/j15_bench_4k_${N}gpu.json 2> gpurun_out/j15_bench_4k_${N}gpu.err; echo "bench rc=$?"
cut -c1-420 gpurun_out/j15_bench_4k_${N}gpu.json
if [ "${3:-trace}" = "trace" ]; then
ZKAES_TRACE=1 timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29572 bench.py --gpus $N --steps 1 --warmup 3 > gpurun_out/j15_trace_bench_${N}gpu.json 2> gpurun_out/j15_phase_trace_4k_${N}gpu_raw.txt; echo "trace rc=$?"
python tools/trace_summary.py gpurun_out/j15_phase_trace_4k_${N}gpu_raw.txt $N
fi
