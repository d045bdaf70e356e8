#!/bin/bash
# ncu passes with the tuner's choices of a plain run (A3D_TUNE_CACHE): launch list, then the full set on the GEMMs
mkdir -p gpurun_out
rm -f gpurun_out/tune_cache.txt
export A3D_TUNE_CACHE=$PWD/gpurun_out/tune_cache.txt
timeout 200 python bench.py --ncu > gpurun_out/plain_ncu_cmd.log 2>&1; echo "plain rc=$? cache lines: $(wc -l < gpurun_out/tune_cache.txt)"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv python bench.py --ncu > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"dense_wgrad_adam_mma|gemm_kernel" -c 14 -f -o gpurun_out/prof_full python bench.py --ncu > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
