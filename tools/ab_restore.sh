#!/bin/bash
# A/B of the un-paint forms of k_step_lane: SNK_RESTORE_THR = 0 (walk) against several thresholds.
for thr in 0 3 5 8 12; do
  echo "== SNK_RESTORE_THR=$thr"; SNK_RESTORE_THR=$thr python tools/ab.py short long 2>&1 | grep -v "^lib"
done
