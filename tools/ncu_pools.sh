# --set full over the max-pool launches of one InceptionI3d forward (2a, 3a, 3b.b3a, 3c.b3a, 4a, 4b.b3a ...): raw CSV only
set -u
TAG=${1:-pools}
ncu --set full --clock-control none -k regex:'maxpool' -c 7 -o gpurun_out/$TAG -f python tools/ncu_target_inception.py > gpurun_out/$TAG.log 2>&1
echo "$TAG rc=$?"
ncu -i gpurun_out/$TAG.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw.csv 2>/dev/null
rm -f gpurun_out/$TAG.ncu-rep
