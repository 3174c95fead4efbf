mkdir -p gpurun_out
TAG=${1:-r1q}
GRAD_PROBE_ONLY=${2:-flat_surf_c4} ncu --set full --clock-control none --import-source on -k regex:gradient_flat -s 1 -c 1 -f -o gpurun_out/${TAG}_gradient_flat python scripts/grad_probe.py > gpurun_out/${TAG}_gradient_flat.log 2>&1
tail -3 gpurun_out/${TAG}_gradient_flat.log
