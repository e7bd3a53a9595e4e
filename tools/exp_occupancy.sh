#!/bin/bash
# occupancy experiment: rollout without the observation store, 5 / 6 / 7 resident blocks per SM
for v in "$@"; do
  FUTBOL_B200_LIB=libfutbol_b200_$v.so python tools/time_rollout.py 1048576 64 10 2>&1 | tail -2
done
