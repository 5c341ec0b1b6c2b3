#!/bin/bash
# development sweep of the traversal tunables: thresholds (0 = simple kernel) for bounce 0 / later bounces / connect, steps per vote
for wl in c3 c4; do
 for cfg in "0 0 0 1" "0 16 0 1" "1 16 16 4" "0 16 16 4" "16 16 16 4" "1 12 12 4" "1 20 20 4"; do
  set -- $cfg
  echo -n "$wl ext0=$1 ext=$2 con=$3 spv=$4 : "
  XRT_TUNING="thr_ext0=$1,thr_ext=$2,thr_con=$3,steps_per_vote=$4" python scripts/profile_step.py $wl 8 | tail -1
 done
done
python scripts/profile_step.py c5 8 | tail -1
