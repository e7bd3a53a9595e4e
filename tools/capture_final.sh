#!/bin/bash
# Round-end ncu evidence for profiles/: full captures of the v0 rollout at the four per-rank sizes of the 2^20 job, the
# launch list of the bench command, and the v1 2v2 / 5v5 rollouts.  Each capture follows a clean run of the same command.
set -e
mkdir -p gpurun_out
for e in 1048576 524288 262144 131072; do
  PROF_ENVS=$e PROF_LAUNCHES=3 python tools/profile_rollout.py > gpurun_out/final_plain_$e.log 2>&1
  PROF_ENVS=$e PROF_LAUNCHES=3 ncu --set full --clock-control none --import-source on -k regex:v0_rollout -s 2 -c 1 \
      -o gpurun_out/prof_final_e$e -f python tools/profile_rollout.py > gpurun_out/final_ncu_$e.log 2>&1
done
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/final_bench_plain.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/final_bench_ncu.log 2>&1
for N in 2 5; do
  PROF_N=$N python tools/profile_rollout_v1.py > gpurun_out/final_v1_plain_$N.log 2>&1
  PROF_N=$N ncu --set full --clock-control none --import-source on -k regex:v1_rollout -s 2 -c 1 \
      -o gpurun_out/prof_final_v1_${N}v${N} -f python tools/profile_rollout_v1.py > gpurun_out/final_v1_ncu_$N.log 2>&1
done
echo captures done
