import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from autonomous_quadrotor_environment_b200 import BatchedQuad, _lib as L
DEV = "cuda:0"
N = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
sensor = len(sys.argv) > 2 and sys.argv[2] == "1"
mk = lambda ld: BatchedQuad(N, 0.01, 25, T=3, precision="f32", async_reset=True, sensor_noise=sensor, seed=11, device=DEV).set_step_loader(ld)
a, b = mk(2), mk(1)
a.reset(); b.reset()
g = torch.Generator(device=DEV); g.manual_seed(5)
names = {L.QS_FIELD_OBS: "obs", L.QS_FIELD_ANG: "ang", L.QS_FIELD_REWARD: "reward", L.QS_FIELD_DONE: "done", L.QS_FIELD_SOLVED: "solved",
         L.QS_FIELD_I: "i", L.QS_FIELD_ABS_SUM: "abs_sum", L.QS_FIELD_PREV_SHAPING: "shaping", L.QS_FIELD_EP_RETURN: "ep_ret",
         L.QS_FIELD_EPISODE: "episode", L.QS_FIELD_FLAGS: "flags"}
if sensor:
    names[L.QS_FIELD_SENSED_OBS] = "sensed"; names[L.QS_FIELD_SENSOR_STATE] = "sstate"
for t in range(40):
    act = (torch.rand(4, N, device=DEV, generator=g) * 2 - 1).contiguous()
    a.step_soa(act); b.step_soa(act)
    for f, nm in names.items():
        fa, fb = a._field(f), b._field(f)
        if not torch.equal(fa, fb):
            d = (fa.double() - fb.double()).abs()
            bad = (fa != fb)
            rows = bad.any(dim=1).nonzero().flatten().tolist()
            cols = bad.any(dim=0).nonzero().flatten()
            print("t=%d %s: %d mismatches, max abs diff %.3g, rows %s, first cols %s" % (t, nm, int(bad.sum()), float(d.max()), rows, cols[:8].tolist()))
            if nm == "obs" and t == 0:
                c = int(cols[0]); print(" a:", fa[:, c].tolist()); print(" b:", fb[:, c].tolist())
print("done")
