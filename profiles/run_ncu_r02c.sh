#!/bin/bash
# Round-2 ncu captures of the two rec-side evaluation kernels that had none this round (run under gpurun from the repo
# root; bounded: the round's last GPU-minute): ST-Gumbel L2 evaluation on augmented rows (k_eval_tiled<KIND_GUMBEL_L2>)
# and the soft-preference evaluation (k_eval_soft) at d=128 against 1M items.
set -u
mkdir -p gpurun_out/ncu
cap() {   # name, kernel regex, target, reps, skip
  timeout 45 ncu --set full --clock-control none --import-source on -k "regex:$2" -s "$5" -c 1 -f -o gpurun_out/ncu/$1 \
      python profiles/prof_targets.py $3 $4 > gpurun_out/ncu/$1.log 2>&1
  ncu -i gpurun_out/ncu/$1.ncu-rep --page raw --csv > gpurun_out/ncu/$1_raw.csv 2>/dev/null
  ncu -i gpurun_out/ncu/$1.ncu-rep --page source --csv > gpurun_out/ncu/$1_source.csv 2>/dev/null
  rm -f gpurun_out/ncu/$1.ncu-rep
}
cap r02_eval_soft_d128 k_eval_soft soft_eval_d128 2 1
cap r02_eval_gumbel_l2 k_eval_tiled gumbel_eval 2 1
ls -la gpurun_out/ncu | tail
