import sys; sys.path.insert(0, '/root/repo')
import numpy as np
from wass_b200 import capi, synth
from oracle import sgbm
r, l, _ = synth.make_pair(2448, 2048, 256, seed=0)
i1, i2 = synth.pad_for_sgbm(r, l, 256)
h = capi.Handle(0)
p = sgbm.wass_params(256, mode=1)
h.sgbm_compute(i1, i2, p)
print(h.sgbm_stats())
