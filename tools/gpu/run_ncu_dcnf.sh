#!/bin/bash
# ncu on one DCNF step (batch 16): launch list, then the full set on the tcgen05 GEMMs, with the tuner's choices of a plain run
mkdir -p gpurun_out
TAG=${1:-r02}
rm -f gpurun_out/tune_cache_dcnf.txt
export A3D_TUNE_CACHE=$PWD/gpurun_out/tune_cache_dcnf.txt
timeout 200 python tools/dcnf_ncu.py > gpurun_out/dcnf_plain.log 2>&1; echo "plain rc=$? cache lines: $(wc -l < gpurun_out/tune_cache_dcnf.txt)"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/ncu_launches_dcnf_$TAG.csv python tools/dcnf_ncu.py > gpurun_out/ncu_launches_dcnf.log 2>&1; echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"gemm_kernel|gemm_pair_kernel" -c 24 -f -o gpurun_out/prof_dcnf_$TAG python tools/dcnf_ncu.py > gpurun_out/ncu_full_dcnf.log 2>&1; echo "ncu full rc=$?"
ncu -i gpurun_out/prof_dcnf_$TAG.ncu-rep --page raw --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,l1tex__m_xbar2l1tex_read_bytes.sum > gpurun_out/ncu_full_dcnf_${TAG}_summary.csv 2> gpurun_out/ncu_summary_dcnf.err; echo "summary rc=$? lines $(wc -l < gpurun_out/ncu_full_dcnf_${TAG}_summary.csv)"
sz=$(stat -c %s gpurun_out/prof_dcnf_$TAG.ncu-rep 2>/dev/null || echo 0)
echo "report bytes: $sz"
if [ "$sz" -gt 40000000 ]; then rm -f gpurun_out/prof_dcnf_$TAG.ncu-rep; fi
