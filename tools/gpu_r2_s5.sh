#!/bin/bash
# Round 2, GPU session 5: staged-tile HiFT ResBlock convolutions, fp16-activation T3 decode mode, sanitizer pass on the new kernels.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
t() { local name=$1; shift; local lim=$1; shift; timeout $lim "$@" > gpurun_out/$name.log 2>&1; echo "$name exit=$?"; tail -n ${TAILN:-4} gpurun_out/$name.log | cut -c1-500; }
PT="python -m pytest -q -m gpu -p no:cacheprovider --timeout 300 --no-header"
TAILN=8 t s5_hift_tests 300 $PT tests/test_gpu_s3gen.py tests/test_hift_drift.py -k "hift or drift or f0"
t s5_hift_new 200 python tools/hift_only.py
CBX_HIFT=legacy t s5_hift_legacy 200 python tools/hift_only.py
TAILN=16 t s5_tests 900 $PT -s tests
TCLS=none t s5_t3only_f16 200 python tools/t3_only.py
TAILN=6 t s5_sanitizer 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest -q -m gpu -p no:cacheprovider --no-header tests/test_gpu_kernels.py -k "weight_resident or splitk or (paged and bf16 and 4)" tests/test_gpu_s3gen.py -k "weight_resident or splitk or (paged and bf16 and 4) or hift_decode"
TAILN=3 t s5_bench 900 python bench.py --steps 1 --warmup 1 --no-extra --cpu-sample none
tail -n 1 gpurun_out/s5_bench.log | cut -c1-900
