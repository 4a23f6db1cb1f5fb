set -x
mkdir -p gpurun_out
(cd tools && ./build/ubench) > gpurun_out/ubench_r2b_raw.txt 2>&1
grep "fq product" gpurun_out/ubench_r2b_raw.txt
