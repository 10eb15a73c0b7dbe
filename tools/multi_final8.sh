# tools/multi_final8.sh -- C3 weak scaling 1/2/4/8 and C5 on 8 GPUs (run under `gpurun --gpus 8`); outputs gpurun_out/r2f_*
run() { n=$1; cfg=$2; st=$3; shift 3
  if [ $n -eq 1 ]; then python bench.py --gpus 1 --config $cfg --steps $st --warmup 5 --no-cpu --no-nmft "$@" 2> gpurun_out/r2f_mg.err
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) bench.py --gpus $n --config $cfg --steps $st --warmup 5 --no-cpu --no-nmft "$@" 2> gpurun_out/r2f_mg.err; fi
}
show() { python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$1', 'N=%d' % d['n_gpus'], d['scaling'], 'V_per_gpu=%d' % d['config']['V_per_gpu'], 'value %.1f sweeps/s' % d['value'], '%.1f us/sweep' % (d['ms_per_step']*1e3), 'e2e %.1f' % d['e2e']['value'], d['config'].get('collective'), 'tau_update %.1f us' % (d['kernel_ms_per_sweep']['tau_update']*1e3), d.get('rank_consistency'))
"; }
cd $GRAFT_REPO_ROOT
for n in 1 2 4 8; do run $n c3 200 | tee gpurun_out/r2f_c3_n$n.json | show "C3 weak"; done
run 8 c5 20 | tee gpurun_out/r2f_c5_n8.json | show "C5 8 GPUs"
