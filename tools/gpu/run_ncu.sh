#!/bin/bash
# ncu passes with the tuner's choices of a plain run (A3D_TUNE_CACHE): launch list, then the full set on the GEMMs and the
# dense update.  Round tag as $1 (default r02).  Budget ~10 GPU-minutes: the full-set pass replays 44 kernels ~40 times each
# under a Python process that ncu slows down considerably (the DCNF variant, run_ncu_dcnf.sh, takes ~3.5 minutes).
mkdir -p gpurun_out
TAG=${1:-r02}
rm -f gpurun_out/tune_cache.txt
export A3D_TUNE_CACHE=$PWD/gpurun_out/tune_cache.txt
timeout 200 python bench.py --ncu > gpurun_out/plain_ncu_cmd.log 2>&1; echo "plain rc=$? cache lines: $(wc -l < gpurun_out/tune_cache.txt)"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/ncu_launches_$TAG.csv python bench.py --ncu > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"dense_wgrad_adam_mma|gemm_kernel|gemm_pair_kernel" -c 44 -f -o gpurun_out/prof_full_$TAG python bench.py --ncu > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
ncu -i gpurun_out/prof_full_$TAG.ncu-rep --page raw --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__occupancy_limit_registers,launch__occupancy_limit_shared_mem,lts__t_sectors_srcunit_tex_op_read.sum,l1tex__m_xbar2l1tex_read_bytes.sum > gpurun_out/ncu_full_${TAG}_summary.csv 2> gpurun_out/ncu_summary.err; echo "summary rc=$? lines $(wc -l < gpurun_out/ncu_full_${TAG}_summary.csv)"
cp gpurun_out/tune_cache.txt gpurun_out/tune_cache_$TAG.txt
# the report itself can exceed gpurun's 64 MiB return limit: keep the summaries, drop the report when it is too large
sz=$(stat -c %s gpurun_out/prof_full_$TAG.ncu-rep 2>/dev/null || echo 0)
echo "report bytes: $sz"
if [ "$sz" -gt 40000000 ]; then rm -f gpurun_out/prof_full_$TAG.ncu-rep; fi
