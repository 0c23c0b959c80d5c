#!/bin/bash
# Multi-GPU evidence (run under `gpurun --gpus 8` from the repo root):   bash tools/evidence_multi.sh <tag> "8 4 2"
# bench.py under torchrun exactly as the driver launches it, headline workload only (configs[4], strong-scaled).
TAG=${1:-evidence}
NS=${2:-"8 4 2"}
OUT=gpurun_out/$TAG
mkdir -p $OUT
PORT=29511
for N in $NS; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT bench.py \
      --gpus $N --steps 20 --warmup 5 --no-per-config --no-cpu > $OUT/bench_c5_n$N.json 2> $OUT/bench_c5_n$N.err
  PORT=$((PORT + 1))
  python - <<PY
import json
d = json.loads(open("$OUT/bench_c5_n$N.json").read().strip().splitlines()[-1])
print("N=%d value %.4g solves/s  %.4f ms/pass  e2e %.4g (%.2f ms)  parity %s" % (d["n_gpus"], d["value"], d["ms_per_step"] / d["config"]["passes_per_step"],
      d["e2e"]["value"], d["e2e"]["ms_per_step"], d.get("parity", {}).get("ok")))
PY
done
