#!/bin/bash
# Section-level ncu pass over every launch of one InceptionI3d forward (op order), raw CSV only.  Run under gpurun; one GPU.
set -u
TAG=${1:-r01_inception}
python tools/ncu_target_inception.py > gpurun_out/${TAG}_plain.log 2>&1 || { echo "plain target failed"; exit 1; }
ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats \
    --clock-control none -k regex:'conv_|stem_|maxpool|avgpool' -c 90 -o gpurun_out/${TAG} -f python tools/ncu_target_inception.py > gpurun_out/${TAG}_ncu.log 2>&1
echo "ncu rc=$?"
ncu -i gpurun_out/${TAG}.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw.csv 2>/dev/null
rm -f gpurun_out/${TAG}.ncu-rep
ls -la gpurun_out/${TAG}_raw.csv
