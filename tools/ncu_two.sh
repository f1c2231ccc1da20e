set -u
export VAD_PAIR_EPI_MIN_KB=0
cap() { # tag regex skip
  ncu --set full --clock-control none --import-source on -k regex:"$2" -s $3 -c 1 -o gpurun_out/$1 -f python tools/ncu_target.py > gpurun_out/$1.log 2>&1
  echo "$1 rc=$?"
  ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1_raw.csv 2>/dev/null
  ncu -i gpurun_out/$1.ncu-rep --page source --csv > gpurun_out/$1_source.csv 2>/dev/null
  rm -f gpurun_out/$1.ncu-rep
}
cap r01v17_l31c3_epi 'conv_umma' 13
cap r01v17_l31c2_pair 'conv_pair' 9
