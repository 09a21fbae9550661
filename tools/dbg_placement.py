"""WSG_SWEEP_DEBUG=1 python tools/dbg_placement.py [impl]: per-band SM, start and end times of the LAST sweep of a frame."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from wass_b200 import capi, synth
from oracle import sgbm
r, l, _ = synth.make_pair(2448, 2048, 256, seed=0)
i1, i2 = synth.pad_for_sgbm(r, l, 256)
h = capi.Handle(0)
h.sgbm_set_impl(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
p = sgbm.wass_params(256, mode=1)
for _ in range(3):
    h.sgbm_compute(i1, i2, p)
print(h.sgbm_stats())
