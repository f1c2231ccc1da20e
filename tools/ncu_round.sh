#!/bin/bash
# ncu evidence for one build, written under gpurun_out/ (kept small: raw CSV pages, reports deleted unless tiny):
#  (1) launch list of one bench run, (2) --set full of the stem kernel, (3) --set full of selected conv launches.
# Run under gpurun; one GPU.
set -u
TAG=${1:-r01}
export VAD_BENCH_NO_INCEPTION=1 VAD_BENCH_NO_SMOOTH=1
python bench.py --steps 1 --warmup 3 --sustain-seconds 0 --no-cpu-baseline > gpurun_out/${TAG}_bench_plain.log 2>&1 || { echo "plain bench failed"; exit 1; }
python tools/ncu_target.py > gpurun_out/${TAG}_target_plain.log 2>&1 || { echo "plain target failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 3 --sustain-seconds 0 --no-cpu-baseline > gpurun_out/${TAG}_ncu_bench.log 2>&1
echo "launch list rc=$?"
full() {  # name, kernel regex, skip, count
  ncu --set full --clock-control none --import-source on -k regex:"$2" -s $3 -c $4 -o gpurun_out/${TAG}_$1 -f \
      python tools/ncu_target.py > gpurun_out/${TAG}_ncu_$1.log 2>&1
  echo "full $1 rc=$?"
  ncu -i gpurun_out/${TAG}_$1.ncu-rep --page raw --csv > gpurun_out/${TAG}_$1_raw.csv 2>/dev/null
  ncu -i gpurun_out/${TAG}_$1.ncu-rep --page source --csv > gpurun_out/${TAG}_$1_source.csv 2>/dev/null
  sz=$(stat -c %s gpurun_out/${TAG}_$1.ncu-rep)
  if [ "$sz" -gt 12000000 ]; then rm -f gpurun_out/${TAG}_$1.ncu-rep; fi
}
full stem 'stem_umma' 0 1
full conv_l1 'conv_umma|conv_thalo|conv_s3x3|conv_pair|conv_tail' 0 8
full conv_l3 'conv_umma|conv_thalo|conv_s3x3|conv_pair|conv_tail' 23 4
full aux 'maxpool|preprocess|avgpool|segment' 0 5
du -sh gpurun_out
