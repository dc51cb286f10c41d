"""PPO training from scratch on the batched simulator (developer tool / demo).
    python tools/train_ppo.py [--envs 65536] [--horizon 128] [--iters 20]
Prints per-iteration mean reward, solved fraction, and the device time of each phase (CUDA events)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from autonomous_quadrotor_environment_b200 import BatchedQuad
from autonomous_quadrotor_environment_b200.ppo import BatchedPPO

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, default=65536)
ap.add_argument("--horizon", type=int, default=128)
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--epochs", type=int, default=10)
ap.add_argument("--lr", type=float, default=5e-4)
a = ap.parse_args()
dev = torch.device("cuda", 0)
env = BatchedQuad(a.envs, 0.01, 1000, training=True, direct_control=1, T=5, precision="f32", async_reset=True, seed=0, device=dev)
env.reset()
ppo = BatchedPPO(env, hidden=128, K_epochs=a.epochs, lr=a.lr, chunk_envs=16384, seed=0)


def timed(fn):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); r = fn(); e1.record(); torch.cuda.synchronize()
    return r, e0.elapsed_time(e1)


for it in range(a.iters):
    batch, t_collect = timed(lambda: ppo.collect(a.horizon))
    losses, t_update = timed(lambda: ppo.update(batch))
    s = env.stats(reset=True)
    mr = float((batch["reward"] * batch["weight"]).sum() / max(1.0, batch["count"]))
    print("iter %3d  mean reward/step %+.4f  episodes %8d  solved %.3f  mean len %6.1f  loss %.4f -> %.4f   collect %.1f ms  update %.1f ms"
          % (it, mr, int(s["n_episodes"]), s["solved_frac"], s["mean_length"], losses[0], losses[-1], t_collect, t_update), flush=True)
