#!/bin/sh
# A/B of kernel builds on the bench workload (run on the GPU box): tools/ab_libs.sh name1 name2 ...  ("tree" = the in-tree library)
for v in "$@"; do
  if [ "$v" = tree ]; then unset BLANGPT_LIB; else export BLANGPT_LIB="$PWD/scratch/libs/libblangpt_$v.so"; fi
  printf "%-12s " "$v"
  timeout 300 python bench.py --steps ${STEPS:-10} --warmup 3 --no-cpu --no-e2e --no-nrpt ${EXTRA} 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('evals/s %.0f  ms/step %.3f  ms/launch %.3f  evals/launch %.0f  clocks %s' % (d['value'], d['ms_per_step'], d['roofline']['ms_per_launch'], d['roofline']['evals_per_launch'], d['clocks']['sm_mhz']))"
done
