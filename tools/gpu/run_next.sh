#!/bin/bash
# First GPU call of the next round: bring up the two kernels written after round 1's GPU budget was spent.
#   1. tc_mcast.cuh  (weight tile multicast across a cluster of 2/4 CTAs) -- never run on hardware: tests first,
#      each under its own timeout (a protocol bug traps through the bounded mbarrier wait; a hang is cut off here);
#   2. tc_persist.cuh (persistent tile loop) -- verified (45 cases), measured neutral on the step;
#   3. the step with each of them as tuner candidates, per-candidate times in the .err files.
mkdir -p gpurun_out
A3D_TEST_MCAST=1 A3D_MCAST=1 timeout 120 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "multicast" > gpurun_out/t_mcast.log 2>&1; echo "mcast tests rc=$?"; tail -6 gpurun_out/t_mcast.log | cut -c1-250
A3D_TEST_PERSIST=1 A3D_PERSIST=1 timeout 120 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "persistent" > gpurun_out/t_persist.log 2>&1; echo "persist tests rc=$?"; tail -3 gpurun_out/t_persist.log | cut -c1-250
for cfg in "0 0" "1 0" "0 1" "1 1" "0 0"; do
set -- $cfg
A3D_MCAST=$1 A3D_PERSIST=$2 A3D_BENCH_U8=0 A3D_AUTOTUNE_VERBOSE=2 timeout 200 python bench.py --steps 100 --warmup 5 --no-cpu-baseline 2> gpurun_out/bench_next_$1$2.err | python -c "
import sys, json
l = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('mcast=$1 persist=$2', round(l['ms_per_step'], 4), round(l['value']), round(l['roofline']['conv_tensor_tflops'], 1))
"
cp gpurun_out/bench_ops_latest.json gpurun_out/bench_ops_next_$1$2.json
done
grep "autotune: kind 2 \[\|autotune: kind 5 \[" gpurun_out/bench_next_10.err | cut -c1-140
