#!/bin/bash
# Round-end ncu evidence for the v0 rollout: full captures at the four per-rank sizes of the 2^20 job and the launch list of
# the bench command.  Each capture follows a clean run of the same command.   bash tools/capture_v0.sh r2
set -e
tag=${1:-r2}
mkdir -p gpurun_out
for e in 1048576 524288 262144 131072; do
  PROF_ENVS=$e PROF_LAUNCHES=3 python tools/profile_rollout.py > gpurun_out/${tag}_v0_plain_$e.log 2>&1
  PROF_ENVS=$e PROF_LAUNCHES=3 ncu --set full --clock-control none --import-source on -k regex:v0_rollout -s 2 -c 1 \
      -o gpurun_out/${tag}_v0_e$e -f python tools/profile_rollout.py > gpurun_out/${tag}_v0_ncu_$e.log 2>&1
done
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/${tag}_bench_plain.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/${tag}_bench_ncu.log 2>&1
echo captures done
