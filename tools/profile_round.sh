#!/bin/bash
# One profiling pass of the C2 bench step under ncu (run on the GPU box through gpurun; never a bench number).
#   tools/profile_round.sh <tag> [launches-per-step]
# Writes under gpurun_out/: launches.csv (device time of every launch of one step), step_traffic.csv (DRAM bytes,
# tensor-pipe %, L2->SM bytes per launch), <tag>_full_all.csv (ncu --set full, raw page, every launch of one step) and
# single-launch reports with source correlation for the kernels named below.  tools/summarize_profiles.py turns them into
# the text files committed under profiles/.
set -u
TAG=${1:-rXX}
PER_STEP=${2:-46}
SKIP=$((2 * PER_STEP + 1))  # + the one-off surround initialisation of the first call
export VNECT_B200_NO_GRAPH=1
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s $SKIP -c $PER_STEP --csv \
    --log-file gpurun_out/launches.csv python tools/profile_step.py 2 1 > gpurun_out/${TAG}_ncu_launches.log 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,l1tex__m_xbar2l1tex_read_bytes.sum \
    --clock-control none -s $SKIP -c $PER_STEP --csv --log-file gpurun_out/step_traffic.csv \
    python tools/profile_step.py 2 1 > gpurun_out/${TAG}_ncu_traffic.log 2>&1
ncu --set full --clock-control none -s $SKIP -c $PER_STEP -f -o /tmp/${TAG}_full_all \
    python tools/profile_step.py 2 1 > gpurun_out/${TAG}_ncu_full.log 2>&1
ncu -i /tmp/${TAG}_full_all.ncu-rep --page raw --csv > gpurun_out/${TAG}_full_all.csv 2>> gpurun_out/${TAG}_ncu_full.log
# source-correlated single launches: the fused block tail (first of the step), the halo 3x3 (res2a_branch2b), the stem
ncu --set full --clock-control none --import-source on -k regex:stem_roll_kernel -s 2 -c 1 -f -o gpurun_out/prof_stem \
    python tools/profile_step.py 2 1 >> gpurun_out/${TAG}_ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"postprocess_kernel|pyramid_kernel" -s 5 -c 2 -f -o gpurun_out/prof_prepost \
    python tools/profile_step.py 2 1 >> gpurun_out/${TAG}_ncu_full.log 2>&1
ls -la gpurun_out/*.csv gpurun_out/*.ncu-rep
