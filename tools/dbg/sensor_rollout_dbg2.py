import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from autonomous_quadrotor_environment_b200 import BatchedQuad, _lib as L
DEV = "cuda:0"
torch.set_printoptions(precision=4, linewidth=250)
N = 4098
for sensor in (True, False):
    mk = lambda: BatchedQuad(N, 0.01, 30, T=3, precision="f32", async_reset=True, sensor_noise=sensor, seed=21, device=DEV)
    a, b1 = mk(), mk().set_step_loader(1)
    a.reset()
    b1._ws.copy_(a._ws)
    acts = (torch.rand(1, 4, N, device=DEV) * 2 - 1)
    rec = a.rollout(1, actions=acts, record_obs=True)
    b1.step_soa(acts[0].contiguous())
    torch.cuda.synchronize()
    d = L.qs_field_desc(); import ctypes as C
    L.check(a.lib.qs_field_info(a._h, L.QS_FIELD_OBS, C.byref(d)))
    full = a._ws[d.ws_offset:d.ws_offset + 17 * d.ld * 4].view(torch.float32).view(17, d.ld)
    print("sensor", sensor, "ld", d.ld)
    print(" rec[1,:6]      ", rec["obs"][0][1, :6].tolist())
    print(" handle[1,:6]   ", full[1, :6].tolist())
    print(" step1[1,:6]    ", b1._field(L.QS_FIELD_OBS)[1, :6].tolist())
    print(" handle pad[0,N:N+6]", full[0, N:N + 6].tolist())
    print(" handle pad[1,N:N+6]", full[1, N:N + 6].tolist())
    print(" rec[0,N-4:]    ", rec["obs"][0][0, N - 4:].tolist(), " handle[0,N-4:N]", full[0, N - 4:N].tolist())
    bad = (rec["obs"][0] - b1._field(L.QS_FIELD_OBS)).abs() > 1e-3
    print(" bad per row", bad.sum(dim=1).tolist())
