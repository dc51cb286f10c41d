"""Developer tool: time qs_step_host (pinned host buffers, H2D + step + D2H per call) at 1M envs with the sensor model."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from autonomous_quadrotor_environment_b200 import BatchedQuad, _lib as L
N = 1 << 20
dev = torch.device("cuda", 0)
env = BatchedQuad(N, 0.01, 1000, training=True, direct_control=1, T=5, precision="f32", async_reset=True, sensor_noise=True, seed=0, device=dev)
env.reset()
a = (torch.rand(4, N) * 2 - 1).pin_memory()
o = torch.empty(14, N).pin_memory(); r = torch.empty(N).pin_memory(); d = torch.empty(N, dtype=torch.uint8).pin_memory()
st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
def step():
    L.check(env.lib.qs_step_host(env._h, C.c_void_p(a.data_ptr()), C.c_void_p(o.data_ptr()), C.c_void_p(r.data_ptr()), C.c_void_p(d.data_ptr()), st))
for _ in range(20): step()
torch.cuda.synchronize()
t0 = time.perf_counter()
K = 200
for _ in range(K): step()
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / K
print("QS_HOST_SLICES=%s  %.1f us/step  %.3e env-steps/s  D2H %.1f GB/s" % (os.environ.get("QS_HOST_SLICES", "8"), dt * 1e6, N / dt, 61 * N / dt / 1e9))
