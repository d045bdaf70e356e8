#!/bin/bash
# Round-end evidence run on one B200: all GPU tests, smoke, default bench, reference arm, ncu, other configs.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/t_gpu_all.log 2>&1; echo "gpu tests rc=$?"; tail -2 gpurun_out/t_gpu_all.log | cut -c1-200
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log | cut -c1-300
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/clocks.csv &
SMI=$!
timeout 400 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_default.err | cut -c1-300
kill $SMI
cp gpurun_out/bench_ops_latest.json gpurun_out/bench_ops_final.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "ref rc=$?"
timeout 200 python bench.py --ncu > gpurun_out/plain_ncu_cmd.log 2>&1 &&
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv python bench.py --ncu > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"dense_wgrad_adam_mma|gemm_kernel" -c 14 -f -o gpurun_out/prof_full python bench.py --ncu > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
timeout 200 python tools/infer_sweep.py > gpurun_out/infer_sweep.log 2>&1; echo "infer rc=$?"; tail -2 gpurun_out/infer_sweep.log | cut -c1-200
timeout 200 python tools/dcnf_bench.py > gpurun_out/dcnf_bench.log 2>&1; echo "dcnf rc=$?"; tail -2 gpurun_out/dcnf_bench.log | cut -c1-200
python - <<'P'
import json
l=json.loads(open('gpurun_out/bench_default.json').read().strip().splitlines()[-1])
print(json.dumps({k: l[k] for k in ('value','ms_per_step','steps','clocks','gpu_launches')}))
print(l['e2e']); print(l['e2e_u8']); print(l['roofline']); print(l['cpu_baseline'])
P
