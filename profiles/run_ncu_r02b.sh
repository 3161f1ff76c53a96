#!/bin/bash
# ncu captures of the kernels added / changed late in round 2 (run under gpurun from the repo root), raw + source pages
# to CSV, and the launch list of the final bench command.
set -u
mkdir -p gpurun_out/ncu
cap() {   # name, kernel regex, target, skip
  ncu --set full --clock-control none --import-source on -k "regex:$2" -s "$4" -c 1 -f -o gpurun_out/ncu/$1 \
      python profiles/prof_targets.py $3 > gpurun_out/ncu/$1.log 2>&1
  ncu -i gpurun_out/ncu/$1.ncu-rep --page raw --csv > gpurun_out/ncu/$1_raw.csv 2>/dev/null
  ncu -i gpurun_out/ncu/$1.ncu-rep --page source --csv > gpurun_out/ncu/$1_source.csv 2>/dev/null
  rm -f gpurun_out/ncu/$1.ncu-rep
}
cap r02_run_r k_run_step_r transr_step 1
cap r02_gumbel_pairs k_gumbel_pairs tup_gumbel_opt 1
cap r02_soft_pairs k_soft_pairs tup_soft_opt 1
cap r02_eval_d128_rot k_eval_tiled eval_d128 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/ncu/launches_r02_final.csv \
    python bench.py --steps 2 --warmup 3 --launches-per-step 4 --no-cpu-baseline > gpurun_out/ncu/bench_under_ncu.log 2>&1
ls -la gpurun_out/ncu | tail -20
