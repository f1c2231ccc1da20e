# one --set full capture of the preprocessing kernel (K1) alone: raw metrics + per-line stall samples
set -u
TAG=${1:-k1}
ncu --set full --clock-control none --import-source on -k regex:preprocess -s 3 -c 1 -o gpurun_out/$TAG -f python tools/k1_time.py > gpurun_out/$TAG.log 2>&1
echo "$TAG rc=$?"
ncu -i gpurun_out/$TAG.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw.csv 2>/dev/null
ncu -i gpurun_out/$TAG.ncu-rep --page source --csv > gpurun_out/${TAG}_source.csv 2>/dev/null
rm -f gpurun_out/$TAG.ncu-rep
