"""Per-stage times of one 2448x2048x256 frame in both matcher modes (device-resident inputs)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from wass_b200 import capi, synth
from oracle import sgbm
r, l, _ = synth.make_pair(2448, 2048, 256, seed=0)
i1, i2 = synth.pad_for_sgbm(r, l, 256)
H, Wp = i1.shape
h = capi.Handle(0)
st = torch.cuda.Stream(); torch.cuda.set_stream(st); h.set_stream(st.cuda_stream)
d1, d2 = torch.from_numpy(i1).cuda(), torch.from_numpy(i2).cuda()
dd = torch.empty((H, Wp), dtype=torch.int16, device="cuda")
for mode, name in ((0, "MODE_SGBM (5 paths, reference default)"), (1, "MODE_HH (8 paths)")):
    p = sgbm.wass_params(256, mode=mode)
    for _ in range(3):
        h.sgbm_compute_device(d1.data_ptr(), d2.data_ptr(), H, Wp, Wp, p, dd.data_ptr())
    torch.cuda.synchronize()
    h.profile_enable(True); h.profile_reset()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(5):
        h.sgbm_compute_device(d1.data_ptr(), d2.data_ptr(), H, Wp, Wp, p, dd.data_ptr())
    e1.record(st); torch.cuda.synchronize()
    prof = h.profile_get(); h.profile_enable(False)
    print(name, "%.2f ms/frame" % (e0.elapsed_time(e1) / 5), {k: round(v[0] / 5, 3) for k, v in prof.items() if v[1]})
