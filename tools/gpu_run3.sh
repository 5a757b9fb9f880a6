#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
t() { local name=$1; shift; local lim=$1; shift; timeout $lim "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?"; tail -n 6 gpurun_out/$name.log; }
t precise 180 python -m pytest tests/test_gpu_kernels.py -v -m gpu -p no:cacheprovider -k "precise" --timeout 120
t bn256 180 python -m pytest tests/test_gpu_kernels.py -v -m gpu -p no:cacheprovider -k "bn256" --timeout 120
t s3gen 300 python -m pytest tests/test_gpu_s3gen.py -v -m gpu -p no:cacheprovider --timeout 120
t e2e 300 python -m pytest tests/test_gpu_e2e.py -v -m gpu -p no:cacheprovider --timeout 200
t t3 400 python -m pytest tests/test_gpu_t3.py -v -m gpu -p no:cacheprovider --timeout 200
t bench16 600 python bench.py --batch 16 --steps 1 --warmup 1
