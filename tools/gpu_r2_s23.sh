#!/bin/bash
# Round 2, GPU session 23 (2 GPUs): bench.py exactly as the driver launches it for N = 2 (weak scaling + strong_256 + mtl + turbo legs), both arms.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
SECONDS=0
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 1 --warmup 3 > gpurun_out/s23_bench2.log 2> gpurun_out/s23_bench2.err; echo "bench N=2 exit=$? wall=${SECONDS}s"
tail -n 1 gpurun_out/s23_bench2.log > gpurun_out/s23_bench2_line.json
grep -E "^\[bench" gpurun_out/s23_bench2.err | cut -c1-420 | tail -14
python - <<'PY'
import json
d = json.load(open("gpurun_out/s23_bench2_line.json"))
print({k: d[k] for k in ("value", "n_gpus", "ms_per_step", "scaling")}, d["e2e"]["value"])
print("strong_256", d["config"].get("strong_256"))
PY
SECONDS=0
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/s23_ref2.log 2> gpurun_out/s23_ref2.err; echo "reference N=2 exit=$? wall=${SECONDS}s"; tail -n 1 gpurun_out/s23_ref2.log | cut -c1-200
