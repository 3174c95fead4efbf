mkdir -p gpurun_out
TAG=${1:-r1q}
GRAD_PROBE_ONLY=${2:-walk} ncu --set full --clock-control none --import-source on -k regex:gradient_walk -s 1 -c 1 -f -o gpurun_out/${TAG}_gradient_walk python scripts/grad_probe.py > gpurun_out/${TAG}_gradient_walk.log 2>&1
tail -3 gpurun_out/${TAG}_gradient_walk.log
