import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from autonomous_quadrotor_environment_b200 import BatchedQuad, _lib as L
DEV = "cuda:0"
for N, seed, T, nmax in ((4160, 11, 3, 25), (4097, 11, 3, 25), (600, 11, 3, 25), (640, 11, 3, 25), (8200, 11, 3, 25), (64 * 9 * 8, 11, 3, 25), (64 * 9 * 8 - 64, 11, 3, 25)):
    mk = lambda ld: BatchedQuad(N, 0.01, nmax, T=T, precision="f32", direct_control=1, async_reset=True, sensor_noise=True, seed=seed, device=DEV).set_step_loader(ld)
    a, b = mk(3), mk(1)
    a.reset(); b.reset()
    g = torch.Generator(device=DEV); g.manual_seed(5)
    for t in range(3):
        b._ws.copy_(a._ws)
        act = (torch.rand(4, N, device=DEV, generator=g) * 2 - 1).contiguous()
        a.step_soa(act); b.step_soa(act)
        d = (a._field(L.QS_FIELD_SENSED_OBS) - b._field(L.QS_FIELD_SENSED_OBS)).abs().max(dim=0).values
        bad = (d > 1e-4).nonzero().flatten()
        print("N=%d seed=%d t=%d  bad envs: %d  first %s  last %s" % (N, seed, t, bad.numel(), bad[:8].tolist(), bad[-4:].tolist()), flush=True)
