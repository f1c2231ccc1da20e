# --set full of ONE launch of the InceptionI3d forward: bash tools/ncu_one.sh <tag> <kernel regex> <skip>
set -u
ncu --set full --clock-control none --import-source on -k regex:"$2" -s $3 -c 1 -o gpurun_out/$1 -f python tools/ncu_target_inception.py > gpurun_out/$1.log 2>&1
echo "$1 rc=$?"
ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1_raw.csv 2>/dev/null
ncu -i gpurun_out/$1.ncu-rep --page source --csv > gpurun_out/$1_source.csv 2>/dev/null
rm -f gpurun_out/$1.ncu-rep
