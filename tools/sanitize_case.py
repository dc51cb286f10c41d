"""Developer tool: a small pass over every kernel family for compute-sanitizer (tools/sanitize.sh).  Ragged env counts so
that partial tiles / chunks and the scalar tails of the vector paths execute; few steps (the sanitizer is 10-100x slower)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from autonomous_quadrotor_environment_b200 import BatchedQuad
from autonomous_quadrotor_environment_b200 import controllers

DEV = "cuda:0"
which = sys.argv[1:] or ["step", "rollout", "control", "policy", "f64", "ppo"]
g = torch.Generator(device=DEV); g.manual_seed(3)

if "step" in which:                                   # qs_step: the four loaders, sensor rows, async / strict auto-reset
    for N in (4099, 2 * 64 * 148 + 6):
        for loader in (0, 1, 2, 3):
            for sensor in (False, True):
                env = BatchedQuad(N, 0.01, 12, T=3, precision="f32", async_reset=True, sensor_noise=sensor, seed=5, device=DEV)
                env.set_step_loader(loader)
                env.reset()
                for t in range(16):
                    env.step_soa((torch.rand(4, N, device=DEV, generator=g) * 2 - 1).contiguous())
                torch.cuda.synchronize()
                print("step N=%d loader=%d sensor=%d ok  episodes=%d" % (N, loader, sensor, env.stats()["n_episodes"]), flush=True)
    env = BatchedQuad(1000, 0.01, 10, T=2, precision="f32", auto_reset=True, seed=5, device=DEV)
    env.reset()
    for t in range(14):
        env.step_soa((torch.rand(4, 1000, device=DEV, generator=g) * 2 - 1).contiguous())
    torch.cuda.synchronize(); print("strict auto-reset ok", flush=True)

if "rollout" in which:                                # qs_rollout: pair kernel (FP32) with both action sources
    for N in (4097, 1002):
        env = BatchedQuad(N, 0.01, 12, T=3, precision="f32", async_reset=True, seed=5, device=DEV)
        env.reset()
        env.rollout(20, record_obs=True, record_actions=True, record_reward=True, record_done=True)
        acts = (torch.rand(20, 4, N, device=DEV, generator=g) * 2 - 1).contiguous()
        env.rollout(20, actions=acts, record_obs=True, record_reward=True, record_done=True)
        torch.cuda.synchronize(); print("rollout N=%d ok" % N, flush=True)
        # sensor model inside the fused rollout: sensor rows in shared memory, warp-wide re-sampler hand-off, action prefetch stages
        env = BatchedQuad(N, 0.01, 12, T=3, precision="f32", async_reset=True, sensor_noise=True, seed=5, device=DEV)
        env.reset()
        env.rollout(20, record_sensed=True, record_reward=True, record_done=True)
        env.rollout(20, actions=acts, record_sensed=True, record_obs=True, record_reward=True, record_done=True)
        torch.cuda.synchronize(); print("sensor rollout N=%d ok  episodes=%d" % (N, env.stats()["n_episodes"]), flush=True)

if "control" in which:                                # qs_control_rollout: LQR and PID laws
    for ctl in (controllers.lqr_controller(), controllers.pid_controller(target_vel=(1.0, 0.0, 0.0))):
        env = BatchedQuad(777, 0.01, 1000, direct_control=0, T=1, precision="f32", seed=5, device=DEV)
        env.reset()
        env.control_rollout(ctl, 20, record_obs=True, record_actions=True, record_aux=True)
        torch.cuda.synchronize()
    print("control rollout ok", flush=True)

if "policy" in which:                                 # qs_policy_rollout: tcgen05 actor MLP + dynamics
    w = dict(np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "actor_128.npz")))
    env = BatchedQuad(128 * 5 + 37, 0.01, 1000, T=5, precision="f32", async_reset=True, seed=5, device=DEV)
    env.reset()
    env.load_actor(w, action_std=0.1)
    env.policy_rollout(8, record_obs=True, record_actions=True, record_logprob=True, record_reward=True, record_done=True)
    torch.cuda.synchronize(); print("policy rollout ok", flush=True)

if "f64" in which:                                    # parity mode: FP64 RK45 replica, AUX rows
    env = BatchedQuad(515, 0.01, 12, T=1, precision="f64", integrator="rk45", aux=True, async_reset=True, seed=5, device=DEV)
    env.reset()
    for t in range(6):
        env.step_soa((torch.rand(4, 515, device=DEV, generator=g, dtype=torch.float64) * 2 - 1).contiguous())
    torch.cuda.synchronize(); print("f64 rk45 ok", flush=True)


if "ppo" in which:                                    # round 2: critic head, sensed-observation policy rollout, qs_ppo_grad / qs_adam_step
    from autonomous_quadrotor_environment_b200.ppo import BatchedPPO
    for sensor in (False, True):
        env = BatchedQuad(128 * 3 + 21, 0.01, 1000, T=5, precision="f32", async_reset=True, sensor_noise=sensor, seed=5, device=DEV)
        env.reset()
        ppo = BatchedPPO(env, hidden=128, K_epochs=2, seed=1)
        out = ppo.iterate(40)                         # 40 steps: the gradient launch splits them into chunks
        torch.cuda.synchronize(); print("ppo iteration sensor=%d ok  losses %s" % (sensor, out["losses"]), flush=True)
