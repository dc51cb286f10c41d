"""Are the one-env-per-lane step kernel (loader 2, scalar FP32 phases) and the pair kernel (loader 3, packed phases) bit-identical?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from autonomous_quadrotor_environment_b200 import BatchedQuad, _lib as L
DEV = "cuda:0"
for sensor in (False, True):
    for direct in (1, 0):
        N, K = 5000, 120
        mk = lambda ld: BatchedQuad(N, 0.01, 40, T=3, precision="f32", direct_control=direct, async_reset=True, sensor_noise=sensor,
                                    seed=11, device=DEV).set_step_loader(ld)
        envs = [mk(ld) for ld in (1, 2, 3)]
        for e in envs: e.reset()
        g = torch.Generator(device=DEV); g.manual_seed(5)
        bad = {}
        for t in range(K):
            if direct:
                act = (torch.rand(4, N, device=DEV, generator=g) * 2 - 1).contiguous()
            else:
                act = torch.stack([torch.rand(N, device=DEV, generator=g) * 20, *(torch.rand(3, N, device=DEV, generator=g) - 0.5)]).contiguous()
            for e in envs: e.step_soa(act)
            fields = [L.QS_FIELD_OBS, L.QS_FIELD_ANG, L.QS_FIELD_REWARD, L.QS_FIELD_ABS_SUM, L.QS_FIELD_PREV_SHAPING, L.QS_FIELD_DONE, L.QS_FIELD_FLAGS, L.QS_FIELD_EPISODE]
            if sensor: fields += [L.QS_FIELD_SENSED_OBS, L.QS_FIELD_SENSOR_STATE]
            for f in fields:
                for name, o in (("1v3", envs[0]), ("2v3", envs[1])):
                    a, b = o._field(f), envs[2]._field(f)
                    if a.dtype.is_floating_point:
                        neq = int(((a != b) & ~(torch.isnan(a) & torch.isnan(b))).sum())
                    else:
                        neq = int((a != b).sum())
                    if neq: bad[(name, f)] = bad.get((name, f), 0) + neq
        print("sensor=%d direct=%d free-running %d steps: mismatching elements per (pair, field): %s" % (sensor, direct, K, bad or "none"), flush=True)
