#!/bin/bash
# A/B of the un-paint forms of k_step_lane: SNK_DEBUG=restore_thr=0 (walk) against several thresholds.
for thr in 0 3 5 8 12; do
  echo "== restore_thr=$thr"; SNK_DEBUG=restore_thr=$thr python tools/ab.py short long 2>&1 | grep -v "^lib"
done
