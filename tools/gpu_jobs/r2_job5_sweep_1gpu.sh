# round 2, GPU job 5 (1 GPU): BASELINE configs[4] sweep through bench.py, both curves, with the CPU oracle next to it up to 2^22
set -x
mkdir -p gpurun_out
: > gpurun_out/r2_sweep_1gpu.jsonl
for curve in 381 377; do
  for L in 16 18 20 22 24; do
    extra=""; if [ $L -le 22 ]; then extra="--cpu-check"; fi
    python bench.py --workload msm --curve $curve --log-n $L --steps 5 --warmup 3 $extra >> gpurun_out/r2_sweep_1gpu.jsonl 2>> gpurun_out/j5_sweep.err
    python bench.py --workload ntt --curve $curve --log-n $L --steps 20 --warmup 3 $extra >> gpurun_out/r2_sweep_1gpu.jsonl 2>> gpurun_out/j5_sweep.err
  done
done
python tools/sweep_table.py gpurun_out/r2_sweep_1gpu.jsonl
