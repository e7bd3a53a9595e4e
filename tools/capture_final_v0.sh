#!/bin/bash
# End-of-round evidence for the final v0 build: GPU tests, ncu captures at the per-rank sizes of the 2^20 job (each after a clean
# run of the same command), the launch list of the bench command, the bench itself.   bash tools/capture_final_v0.sh "sizes"
set -e
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2f_tests.log 2>&1 || { tail -20 gpurun_out/r2f_tests.log; exit 1; }
tail -2 gpurun_out/r2f_tests.log
for e in ${1:-1048576 524288 262144 131072}; do
  PROF_ENVS=$e PROF_LAUNCHES=3 python tools/profile_rollout.py > gpurun_out/r2f_v0_plain_$e.log 2>&1
  PROF_ENVS=$e PROF_LAUNCHES=3 ncu --set full --clock-control none --import-source on -k regex:v0_rollout -s 2 -c 1 \
      -o gpurun_out/r2f_v0_e$e -f python tools/profile_rollout.py > gpurun_out/r2f_v0_ncu_$e.log 2>&1
done
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/r2f_bench_plain.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2f_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/r2f_bench_ncu.log 2>&1
python bench.py > gpurun_out/r2_bench_f.json 2> gpurun_out/r2_bench_f.err
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')"
echo captures done
