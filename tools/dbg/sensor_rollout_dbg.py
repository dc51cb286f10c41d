"""Developer tool: one sensor step through rollout (pair kernel), qs_step loader 3 and qs_step loader 1 from the same workspace."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from autonomous_quadrotor_environment_b200 import BatchedQuad, _lib as L
DEV = "cuda:0"
for N in (4096, 4098, 4100):
    mk = lambda: BatchedQuad(N, 0.01, 30, T=3, precision="f32", async_reset=True, sensor_noise=True, seed=21, device=DEV, params={"gps_blend": 20.0})
    a, b3, b1 = mk(), mk(), mk().set_step_loader(1)
    a.reset()
    g = torch.Generator(device=DEV); g.manual_seed(9)
    for it in range(3):
        b3._ws.copy_(a._ws); b1._ws.copy_(a._ws)
        acts = (torch.rand(1, 4, N, device=DEV, generator=g) * 2 - 1)
        rec = a.rollout(1, actions=acts, record_sensed=True, record_obs=True)
        b3.step_soa(acts[0].contiguous()); b1.step_soa(acts[0].contiguous())
        torch.cuda.synchronize()
        for name, x in (("rollout.rec", rec["sensed_obs"][0]), ("rollout.handle", a._field(L.QS_FIELD_SENSED_OBS)), ("step3", b3._field(L.QS_FIELD_SENSED_OBS))):
            d = (x - b1._field(L.QS_FIELD_SENSED_OBS)).abs()
            bad = d > 1e-3
            print("N=%d it=%d %-15s vs step1: bad per row %s  bad envs %s" % (N, it, name, bad.sum(dim=1).tolist(), torch.nonzero(bad.any(dim=0)).flatten()[:8].tolist()))
        d = (a._field(L.QS_FIELD_SENSOR_STATE) - b1._field(L.QS_FIELD_SENSOR_STATE)).abs() > 1e-3
        print("      sensor_state rollout vs step1 bad per row", d.sum(dim=1).tolist())
        d = (rec["obs"][0] - b1._field(L.QS_FIELD_OBS)).abs() > 1e-3
        print("      true obs rollout vs step1 bad per row", d.sum(dim=1).tolist(), "warm envs", int(b1.warmup.sum()), "done", int((b1.done_flags & 1).sum()))
