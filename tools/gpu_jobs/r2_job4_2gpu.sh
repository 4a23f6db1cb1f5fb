# round 2, GPU job 4 (2 GPUs): multi-rank and single-process multi-GPU tests, 4 KiB proof on 2 GPUs with a phase trace
set -x
mkdir -p gpurun_out
nvidia-smi -L
python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/j4_pytest_multi_2gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/j4_pytest_multi_2gpu.log
tail -5 gpurun_out/j4_pytest_multi_2gpu.log
ZKAES_TRACE=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/r2_bench_4k_2gpu.json 2> gpurun_out/r2_phase_trace_4k_2gpu_raw.txt; echo "bench rc=$?"
cat gpurun_out/r2_bench_4k_2gpu.json | cut -c1-700
grep "zkaes" gpurun_out/r2_phase_trace_4k_2gpu_raw.txt | tail -30 | cut -c1-90
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload msm --log-n 24 --steps 5 --warmup 3 > gpurun_out/r2_bench_msm24_2gpu.json 2> gpurun_out/j4_msm.err; cat gpurun_out/r2_bench_msm24_2gpu.json | cut -c1-300
