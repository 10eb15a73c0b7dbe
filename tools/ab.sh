#!/bin/bash
# A/B of builds of the library on the SAME box: tools/ab.sh [reps] lib1.so lib2.so ...   (run under gpurun)
# e.g. build/libdesman_b200_head.so (nvcc on `git archive HEAD`) against desman_b200/libdesman_b200.so.
# bench.py at C3, 200 sweeps, no CPU leg; the per-sweep time is reproducible to ~0.1 us between repetitions on one box.
R=${1:-2}; shift
for rep in $(seq $R); do
  for v in "$@"; do
    DESMAN_B200_LIB=$v timeout 100 python bench.py --no-cpu --no-nmft --steps 200 2>/dev/null | python -c "
import json, sys
d = json.loads(sys.stdin.read()); k = d['kernel_ms_per_sweep']
print('%-36s %7.2f us/sweep  tau_update %5.1f  tau_group %5.1f tau_sample %5.1f mu %5.1f finalize %5.1f draw %5.1f maintain %4.1f  e2e %6.0f work %d' % (
    '$v', d['ms_per_step'] * 1e3, k.get('tau_update', 0) * 1e3, k['tau_group'] * 1e3, k['tau_sample'] * 1e3, k['mu_stats'] * 1e3, k['finalize'] * 1e3,
    k['draw_gamma_eta'] * 1e3, k['maintain'] * 1e3, d['e2e']['value'], d['tau_groups']['work']))"
  done
done
