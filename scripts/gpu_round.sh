#!/bin/bash
# One GPU-box visit: parity tests, the default bench (both arms), then the ncu launch list + full captures.
# Usage: scripts/gpu_round.sh <tag>
set -u
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/${TAG}_pytest.log
tail -3 $OUT/${TAG}_pytest.log
timeout 600 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_reference.json 2> $OUT/${TAG}_bench_reference.err; echo "ref rc=$?"
KERNELS="raycast_kernel raycast_long_kernel occupancy_tma_kernel gradient_walk_kernel xpass_lanes_kernel ysweep_ring_kernel zwalk_kernel" timeout 1200 bash scripts/gpu_profile.sh $TAG > /dev/null 2>&1
python scripts/show_bench.py $OUT/${TAG}_bench.json
