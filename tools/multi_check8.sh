# tools/multi_check8.sh -- the multi-GPU measurements DESIGN.md quotes (run under `gpurun --gpus 8`)
N=${1:-8}
run() { # n config extra
  n=$1; cfg=$2; shift 2
  if [ $n -eq 1 ]; then python bench.py --gpus 1 --config $cfg --steps ${STEPS:-100} --warmup 5 --no-cpu --no-nmft "$@" 2> gpurun_out/r2_mg.err
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) bench.py --gpus $n --config $cfg --steps ${STEPS:-100} --warmup 5 --no-cpu --no-nmft "$@" 2> gpurun_out/r2_mg.err; fi
}
show() { python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$1', 'N=%d' % d['n_gpus'], d['scaling'], 'V_per_gpu=%d' % d['config']['V_per_gpu'], 'value %.1f sweeps/s' % d['value'], '%.1f us/sweep' % (d['ms_per_step']*1e3), d['config'].get('collective'), 'tau_update %.1f us frac %.3f' % (d['kernel_ms_per_sweep']['tau_update']*1e3, d['roofline']['frac']), 'exchange %.1f' % (d['kernel_ms_per_sweep']['other']*1e3), d.get('rank_consistency'))
"; }
for n in 1 2 4 8; do [ $n -le $N ] && run $n c3 | tee gpurun_out/r2_c3_n$n.json | show "C3 weak"; done
STEPS=20; for n in 1 2 4; do [ $n -le $N ] && run $n c4 --strong | tee gpurun_out/r2_c4_strong_n$n.json | show "C4 strong"; done
[ 8 -le $N ] && run 8 c5 | tee gpurun_out/r2_c5_n8.json | show "C5 8 GPUs"
