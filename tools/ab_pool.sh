#!/bin/sh
# A/B of the two exploration kernels on the bench workload (run on the GPU box)
for v in ${1:-0 1}; do
  echo "POOL=$v"
  BLANGPT_POOL=$v timeout 300 python bench.py --steps ${STEPS:-10} --warmup 3 --no-cpu --no-e2e ${EXTRA} 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('evals/s %.0f  ms/step %.3f  ms/launch %.3f  evals/launch %.0f  clocks %s' % (d['value'], d['ms_per_step'], d['roofline']['ms_per_launch'], d['roofline']['evals_per_launch'], d['clocks']))"
done
