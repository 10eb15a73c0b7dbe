# tools/multi_check.sh -- multi-GPU tests + the C3 weak-scaling lines (run under `gpurun --gpus 4`)
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for n in 1 2 4; do
  if [ $n -eq 1 ]; then python bench.py --gpus 1 --steps 200 --warmup 5 --no-cpu --no-nmft > gpurun_out/r2_scale_$n.json 2> gpurun_out/r2_scale_$n.err
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2950$n bench.py --gpus $n --steps 200 --warmup 5 --no-cpu --no-nmft > gpurun_out/r2_scale_$n.json 2> gpurun_out/r2_scale_$n.err; fi
  python -c "
import json;d=json.load(open('gpurun_out/r2_scale_$n.json'));print($n, round(d['value']), round(d['ms_per_step']*1e3,1), d['config'].get('collective'), d.get('rank_consistency'), {k:round(v*1e3,1) for k,v in d['kernel_ms_per_sweep'].items() if v>0})"
done
