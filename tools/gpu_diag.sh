#!/bin/bash
# Run the GPU tests in separate processes (a trapped kernel poisons the CUDA context of its own process only)
# and keep the logs under gpurun_out/.  Usage (on the GPU box):  bash tools/gpu_diag.sh
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/nvidia_smi.txt 2>&1
run() {  # name, env, pytest args...
  local name=$1; shift; local envs=$1; shift
  env $envs timeout 1200 python -m pytest "$@" -q -m gpu --no-header -p no:cacheprovider > gpurun_out/$name.log 2>&1
  echo "$name exit=$? $(tail -n 1 gpurun_out/$name.log)" | tee -a gpurun_out/summary.txt
}
run kernels_simt X=1 tests/test_gpu_kernels.py -k simt
run kernels_tc X=1 tests/test_gpu_kernels.py -k "tc"
run t3_simt "CBX_GEMM=simt CBX_ATTN=simt" tests/test_gpu_t3.py
run s3gen_simt "CBX_GEMM=simt CBX_ATTN=simt" tests/test_gpu_s3gen.py
run t3 X=1 tests/test_gpu_t3.py
run s3gen X=1 tests/test_gpu_s3gen.py
run e2e X=1 tests/test_gpu_e2e.py
for f in kernels_simt kernels_tc t3_simt s3gen_simt t3 s3gen e2e; do echo "=== $f"; grep -E "^(FAILED|ERROR|E  )|passed|failed" gpurun_out/$f.log | head -40; done
