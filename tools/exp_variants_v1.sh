#!/bin/bash
# v1 counterpart of exp_variants.sh: tools/exp_variants_v1.sh "N n_envs" variant...   (NCU=1 adds the counters)
M=gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed,sm__icc_request_hit_rate.pct,sm__icc_requests.sum.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio
cfg="$1"; shift
for v in "$@"; do
  if [ "$v" = default ]; then unset FUTBOL_B200_LIB; else export FUTBOL_B200_LIB=libfutbol_b200_$v.so; fi
  python tools/time_rollout_v1.py $cfg 64 5 2>&1 | tail -1 | sed "s/^/[$v] /"
  if [ -n "$NCU" ]; then
    ncu --metrics $M --clock-control none -k regex:v1_rollout -s 2 -c 1 python tools/time_rollout_v1.py $cfg 64 1 2>&1 | grep -E "gcc__|icc_|issue_active|inst_executed|warps_active" | awk -v v="$v" '{print "[" v "] ncu", $1, $NF}'
  fi
done
