# round 2, GPU job 6 (N GPUs, N = $1): world-N multi-GPU tests, 4 KiB proof on N GPUs with a phase trace, sharded MSM at 2^22 / 2^24 / 2^26
N=$1
set -x
mkdir -p gpurun_out
nvidia-smi -L | head -8
python -m pytest tests/test_gpu_multi.py -m gpu -x -q -k "$N" > gpurun_out/j6_pytest_multi_${N}gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/j6_pytest_multi_${N}gpu.log
tail -4 gpurun_out/j6_pytest_multi_${N}gpu.log
ZKAES_TRACE=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r2_bench_4k_${N}gpu.json 2> gpurun_out/r2_phase_trace_4k_${N}gpu_raw.txt; echo "bench rc=$?"
cut -c1-400 gpurun_out/r2_bench_4k_${N}gpu.json
: > gpurun_out/r2_sweep_${N}gpu.jsonl
for L in 22 24 26; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$((L-20)) bench.py --gpus $N --workload msm --curve 381 --log-n $L --steps 5 --warmup 3 >> gpurun_out/r2_sweep_${N}gpu.jsonl 2>> gpurun_out/j6_sweep_${N}.err
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --workload msm --curve 377 --log-n 24 --steps 5 --warmup 3 >> gpurun_out/r2_sweep_${N}gpu.jsonl 2>> gpurun_out/j6_sweep_${N}.err
python tools/sweep_table.py gpurun_out/r2_sweep_${N}gpu.jsonl
