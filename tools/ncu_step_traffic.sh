#!/bin/bash
# DRAM bytes of ONE bench step (every kernel of the timed step, one ncu pass per kernel): the `traffic` of bench.py's whole-step
# roofline.  The first SKIP launches (peak probes, warm-up steps) run unprofiled; the summary takes the launches between the last
# two segment-mean kernels it sees, i.e. the last full step.  Run under gpurun; writes gpurun_out/<tag>_step_traffic.csv, summarised by tools/step_traffic.py into profiles/.
set -u
TAG=${1:-r02}
export VAD_BENCH_NO_INCEPTION=1 VAD_BENCH_NO_SMOOTH=1 VAD_BENCH_NO_TF32=1
python bench.py --steps 1 --warmup 3 --sustain-seconds 0 --no-cpu-baseline > gpurun_out/${TAG}_traffic_plain.log 2>&1 || { echo "plain bench failed"; exit 1; }
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --launch-skip ${SKIP:-900} -c ${COUNT:-1200} --csv \
    --log-file gpurun_out/${TAG}_step_traffic.csv python bench.py --steps 1 --warmup 3 --sustain-seconds 0 --no-cpu-baseline \
    > gpurun_out/${TAG}_traffic_ncu.log 2>&1
echo "rc=$?"
