#!/bin/bash
# ncu evidence for the v1 rollout kernels: tag = file prefix under gpurun_out/, then "N:envs" pairs.
#   bash tools/capture_v1.sh r2a 5:262144 10:65536 2:1048576
set -e
tag=$1; shift
mkdir -p gpurun_out
for spec in "$@"; do
  N=${spec%%:*}; E=${spec##*:}
  python tools/time_rollout_v1.py $N $E > gpurun_out/${tag}_v1_time_${N}v${N}.log 2>&1 || true
  PROF_N=$N PROF_ENVS=$E PROF_LAUNCHES=3 python tools/profile_rollout_v1.py > gpurun_out/${tag}_v1_plain_${N}v${N}.log 2>&1
  PROF_N=$N PROF_ENVS=$E PROF_LAUNCHES=3 ncu --set full --clock-control none --import-source on -k regex:v1_rollout -s 2 -c 1 \
      -o gpurun_out/${tag}_v1_${N}v${N} -f python tools/profile_rollout_v1.py > gpurun_out/${tag}_v1_ncu_${N}v${N}.log 2>&1
done
echo captures done
