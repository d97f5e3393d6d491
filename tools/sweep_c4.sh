#!/bin/bash
# forward-kernel shape sweep at the C4 geometry (512x512, 720 angles, 16-image slice)
export TIME_FWD_SHAPES=16x512x720
for ka in 1 2 4; do for st in 2 3; do for r in 6 9 12; do
  CTR_FWD_KA=$ka CTR_FWD_STAGES=$st CTR_FWD_R=$r timeout 120 python tools/time_fwd.py 2>&1 | grep "B="
done; done; done
