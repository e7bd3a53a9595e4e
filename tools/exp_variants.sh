#!/bin/bash
# Times the fused rollout for build variants (tools/build_variant.py) and, with NCU=1, reads the instruction-cache
# counters of each: tools/exp_variants.sh default rp rpa ...
M=gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed,sm__icc_request_hit_rate.pct,sm__icc_requests.sum.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum
for v in "$@"; do
  if [ "$v" = default ]; then unset FUTBOL_B200_LIB; else export FUTBOL_B200_LIB=libfutbol_b200_$v.so; fi
  python tools/time_rollout.py ${N:-1048576} 64 10 2>&1 | tail -2 | sed "s/^/[$v] /"
  if [ -n "$NCU" ]; then
    ncu --metrics $M --clock-control none -k regex:rollout -s 3 -c 1 python tools/time_rollout.py ${N:-1048576} 64 1 2>&1 | grep -E "gcc__|icc_|issue_active|inst_executed|warps_active|time_duration" | awk -v v="$v" '{print "[" v "] ncu", $1, $NF}'
  fi
done
