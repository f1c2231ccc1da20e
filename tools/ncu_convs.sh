#!/bin/bash
# Section-level ncu pass over every conv launch of one forward (op order), raw CSV only.  Run under gpurun; one GPU.
#   tools/ncu_convs.sh <tag> [ENV=VALUE ...]
set -u
TAG=$1; shift
for kv in "$@"; do export "$kv"; done
python tools/ncu_target.py > gpurun_out/${TAG}_plain.log 2>&1 || { echo "plain target failed"; exit 1; }
ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section WarpStateStats --section SchedulerStats --section LaunchStats \
    --clock-control none -k regex:'conv_|stem_' -c 60 -o gpurun_out/${TAG}_convs -f python tools/ncu_target.py > gpurun_out/${TAG}_ncu.log 2>&1
echo "ncu rc=$?"
ncu -i gpurun_out/${TAG}_convs.ncu-rep --page raw --csv > gpurun_out/${TAG}_convs_raw.csv 2>/dev/null
rm -f gpurun_out/${TAG}_convs.ncu-rep
ls -la gpurun_out/${TAG}_convs_raw.csv
