#!/bin/bash
# Multi-GPU box visit: the multi-GPU parity tests, then the 1/2/4/N series on config 2 (frames) and the 8K stand-in (tiles).
# Usage: scripts/gpu_scale.sh <tag> <N>
set -u
TAG=${1:-r2final}; NMAX=${2:-8}; OUT=gpurun_out; mkdir -p $OUT
timeout 500 python -m pytest tests/test_multigpu_gpu.py -m gpu -x -q 2>&1 | tail -3 | tee $OUT/${TAG}_multigpu_pytest_${NMAX}gpu.log
for n in 8 4 2; do
  [ $n -le $NMAX ] || continue
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 60 --warmup 5 --tf-changes 50 --no-cpu-baseline > $OUT/${TAG}_scale_n$n.json 2> $OUT/${TAG}_scale_n$n.err
done
python bench.py --steps 60 --warmup 5 --tf-changes 50 --no-cpu-baseline --quick > $OUT/${TAG}_scale_n1.json 2>/dev/null
for n in 8 4 2; do
  [ $n -le $NMAX ] || continue
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n bench.py --gpus $n --workload c5s --steps 20 --warmup 3 --tf-changes 20 --quick --no-cpu-baseline > $OUT/${TAG}_c5s_n$n.json 2> $OUT/${TAG}_c5s_n$n.err
done
python bench.py --workload c5s --steps 20 --warmup 3 --tf-changes 20 --quick --no-cpu-baseline > $OUT/${TAG}_c5s_n1.json 2>/dev/null
python - <<PY
import json
for f in ["scale_n1","scale_n2","scale_n4","scale_n8","c5s_n1","c5s_n2","c5s_n4","c5s_n8"]:
    try:
        d=json.loads(open("$OUT/${TAG}_"+f+".json").read().strip().splitlines()[-1])
        print(f, round(d["value"]), round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"]), "rebuild", round(d["ess_rebuild_ms"]["median"],4), d.get("sharded_rebuild_matches"), d.get("tiles_match_single_gpu_frame"), (d.get("tiles_8k") or {}).get("speedup"))
    except Exception as e: print(f, "ERR", e)
PY
