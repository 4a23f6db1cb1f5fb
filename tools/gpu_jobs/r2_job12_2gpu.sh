# round 2, GPU job 12 (N GPUs, N = $1, default 2): multi-GPU tests and the 4 KiB proof with the Lagrange-basis round 1, the rank-sharded
# polynomial preparation (w / z_A / z_B owners + broadcast, r_alpha and f slices + all-gather) and the chunked heavy columns of t
N=${1:-2}
set -x
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout -k 10 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/j12_pytest_multi_${N}gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/j12_pytest_multi_${N}gpu.log
tail -4 gpurun_out/j12_pytest_multi_${N}gpu.log
ZKAES_TRACE=1 timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/j12_bench_4k_${N}gpu.json 2> gpurun_out/j12_phase_trace_4k_${N}gpu_raw.txt; echo "bench rc=$?"
cut -c1-500 gpurun_out/j12_bench_4k_${N}gpu.json
grep "zkaes" gpurun_out/j12_phase_trace_4k_${N}gpu_raw.txt | tail -40 | cut -c1-80
