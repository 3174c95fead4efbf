#!/bin/bash
# Run on the GPU box (via gpurun): launch list + one full ncu capture per hot kernel.
# Usage: scripts/gpu_profile.sh <tag>     (outputs under gpurun_out/<tag>_*)
set -u
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
BENCH="python bench.py --steps 6 --warmup 3 --tf-changes 3 --quick --no-cpu-baseline"
# 1) every launch with its device time (cold-cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv $BENCH > $OUT/${TAG}_launches.log 2>&1
# 2) full captures of the hot kernels
for K in ${KERNELS:-raycast_kernel occupancy_tma_kernel gradient_int_kernel xpass_vec4_kernel ysweep_kernel zwalk_kernel tf_masks_kernel}; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 1 -c 2 -f -o $OUT/${TAG}_$K $BENCH > $OUT/${TAG}_$K.log 2>&1
done
ls -la $OUT
