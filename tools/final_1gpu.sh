#!/bin/bash
# tools/final_1gpu.sh TAG -- what is recorded per round on ONE GPU box (run under gpurun): GPU tests, bench line, reference arm,
# ncu launch list and full capture of the sweep's kernels (-> tools/summarize_profiles.py TAG)
TAG=${1:-r2f}
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/${TAG}_pytest.txt
timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -c 600 gpurun_out/${TAG}_bench.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/${TAG}_ref.json 2> gpurun_out/${TAG}_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 50 --warmup 10 --no-cpu --no-nmft > gpurun_out/${TAG}_ncu1.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"tau_group_tc|tau_open|tau_sample|mu_binomial|mu_class|ll_table|finalize_sweep|draw_gamma|table_maintain" -s 135 -c 9 -f -o gpurun_out/prof_${TAG}_final python bench.py --steps 40 --warmup 10 --no-cpu --no-nmft > gpurun_out/${TAG}_ncu2.log 2>&1
python - <<P
import json
for f in ("bench","ref"):
    try:
        d=json.loads([l for l in open("gpurun_out/${TAG}_%s.json"%f) if l.startswith("{")][-1])
        print(f, d["value"], d.get("ms_per_step"), (d.get("e2e") or {}).get("value"), (d.get("roofline") or {}).get("frac"), (d.get("cpu_baseline") or {}).get("value"), (d.get("nmft") or {}).get("ms_per_iter") if d.get("nmft") else None)
    except Exception as e: print(f, "ERR", e)
P
ls -la gpurun_out/prof_${TAG}_final.ncu-rep gpurun_out/launches_${TAG}.csv
