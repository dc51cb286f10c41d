"""Kernel micro-benchmarks (developer tool; prints one JSON line per case).  Usage on the GPU box:
    python tools/kbench.py [case ...]      cases: step, rollout, all
Times with CUDA events on the launching stream after warm-up."""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from autonomous_quadrotor_environment_b200 import BatchedQuad, _lib as L

DEV = torch.device("cuda", 0)


def emit(d):
    """one compact line per case (full JSON with KBENCH_JSON=1)"""
    if os.environ.get("KBENCH_JSON"):
        print(json.dumps(d), flush=True)
        return
    tag = "%s N=%d %s/%s S=%d T=%d%s%s%s%s" % (d["case"], d["N"], d["precision"], d["integrator"], d["substeps"], d["T"],
                                              " strict-reset" if d.get("auto_reset") else "", " async-reset" if d.get("async_reset") else "",
                                              " sensor" if d.get("sensor_noise") else "", " K=%d" % d["K"] if "K" in d else "")
    print("%-62s %8.2f us/step  %.3e env-steps/s" % (tag, d["ms"] * 1e3 / d.get("K", 1), d["steps_per_s"]), flush=True)


def time_ms(fn, iters, warm=20):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def case_ppo_grad(N=1 << 16, K=128, iters=5, warm=2):
    """qs_ppo_grad on synthetic rollout buffers: ms per full-batch gradient of one network, tensor-core rate in the FLOPs of
    the forward + backward products (2 x (80+128+16) x 128 forward incl. the bias block, 2 x (128+16+16+128+80) x 128 backward)."""
    from autonomous_quadrotor_environment_b200.ppo import BatchedPPO
    ppo = BatchedPPO(None, hidden=128, seed=0, device=DEV)
    g = torch.Generator(device=DEV); g.manual_seed(1)
    r = lambda *s: torch.randn(*s, device=DEV, generator=g)
    b = dict(hist0=r(75, N).bfloat16().float(), entries=r(K, 15, N).bfloat16().float(), actions=r(K, 4, N) * 0.3,
             logprob=r(K, 4, N) * 0.1 + 1.38, adv=r(K, N), returns=r(K, N), weight=torch.ones(K, N, device=DEV), count=float(K * N))
    t = ppo._batch_tensors(b)
    loss = torch.zeros(2, dtype=torch.float64, device=DEV)
    st = C.c_void_p(torch.cuda.current_stream(DEV).cuda_stream)
    flops = 2.0 * 128 * ((80 + 128 + 16) + (128 + 16 + 16 + 128 + 80)) * K * N
    for which, name in ((L.QS_PPO_ACTOR, "actor"), (L.QS_PPO_CRITIC, "critic")):
        bt = L.qs_ppo_batch(N, K, 0, t["hist0"].data_ptr(), t["entries"].data_ptr(), None, t["actions"].data_ptr(), t["logprob"].data_ptr(),
                            t["adv"].data_ptr(), t["returns"].data_ptr(), t["weight"].data_ptr())
        net, grad = ppo._net_ptrs(ppo._flat, which == L.QS_PPO_CRITIC), ppo._net_ptrs(ppo._grad, which == L.QS_PPO_CRITIC)

        def fn():
            L.check(ppo.lib.qs_ppo_grad(C.byref(bt), C.byref(net), C.byref(grad), which, 0.1, 0.2, float(K * N), loss[which].data_ptr(), st))

        ms = time_ms(fn, iters, warm=warm)
        print("ppo_grad %-6s N=%d K=%d   %9.3f ms  %.3e samples/s  %.1f TFLOP/s (BF16 tensor, algorithmic)  %.2f us per 128-sample tile-step per SM"
              % (name, N, K, ms, N * K / ms * 1e3, flops / ms * 1e-9, ms * 1e3 / (N / 128 * K / 148)), flush=True)


def case_step(N=1 << 20, iters=300, **kw):
    cfg = dict(T=5, auto_reset=False, async_reset=False, precision="f32", integrator="rk4", substeps=1, n=1000, act_scale=1.0)
    cfg.update(kw)
    env = BatchedQuad(N, 0.01, cfg["n"], training=True, direct_control=1, T=cfg["T"], precision=cfg["precision"],
                      integrator=cfg["integrator"], substeps=cfg["substeps"], auto_reset=cfg["auto_reset"],
                      async_reset=cfg["async_reset"], sensor_noise=cfg.get("sensor_noise", False), seed=0, device=DEV)
    env.reset()
    acts = [((torch.rand(4, N, device=DEV, dtype=env.dtype) * 2 - 1) * cfg["act_scale"]).contiguous() for _ in range(8)]
    st = C.c_void_p(torch.cuda.current_stream(DEV).cuda_stream)
    k = [0]

    def fn():
        L.check(env.lib.qs_step(env._h, C.c_void_p(acts[k[0] % 8].data_ptr()), None, None, None, None, st))
        k[0] += 1

    ms = time_ms(fn, iters, warm=60)
    s = env.stats()
    emit({"case": "step", "N": N, **cfg, "ms": ms, "steps_per_s": N / ms * 1e3,
          "mean_len": s["mean_length"], "episodes_per_step": s["n_episodes"] / max(1, s["n_steps"] / N) / N})


def case_rollout(N=1 << 20, K=32, iters=10, **kw):
    cfg = dict(T=5, auto_reset=False, async_reset=False, precision="f32", integrator="rk4", substeps=1, n=1000)
    cfg.update(kw)
    env = BatchedQuad(N, 0.01, cfg["n"], training=True, direct_control=1, T=cfg["T"], precision=cfg["precision"],
                      integrator=cfg["integrator"], substeps=cfg["substeps"], auto_reset=cfg["auto_reset"],
                      async_reset=cfg["async_reset"], sensor_noise=cfg.get("sensor_noise", False), seed=0, device=DEV)
    env.reset()
    rec = cfg.get("record", False)                 # record = the stream a trainer reads: sensed (or true) observation, reward, done
    sens = bool(cfg.get("sensor_noise", False))
    acts = None
    if cfg.get("actions", False):                  # the API path: actions read from a (K,4,N) device tensor instead of drawn in-kernel
        import torch
        acts = (torch.rand(K, 4, N, device=DEV) * 2 - 1).contiguous()
    fn = ((lambda: env.rollout(K, actions=acts, record_sensed=sens, record_obs=not sens, record_reward=True, record_done=True)) if rec
          else (lambda: env.rollout(K, actions=acts)))
    for _ in range(int(cfg.get("preroll", 0))):    # untimed launches: past the first episode turnover
        env.rollout(K, actions=acts)
    ms = time_ms(fn, iters, warm=3)
    emit({"case": "rollout" + ("+rec" if rec else "") + ("+actbuf" if acts is not None else ""), "N": N, "K": K, **cfg, "ms": ms, "steps_per_s": N * K / ms * 1e3})


def case_policy(N=1 << 20, K=128, iters=3, sigma=0.1, record=True, critic=False):
    import numpy as np
    g = dict(np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "actor_128.npz")))
    env = BatchedQuad(N, 0.01, 1000, training=True, direct_control=1, T=5, precision="f32", async_reset=True, seed=0, device=DEV)
    env.reset()
    crit = None
    if critic:                                          # critic of the same shape; the actor's hidden layers stand in for trained weights
        crit = {"critic_0_weight": g["actor_0_weight"], "critic_0_bias": g["actor_0_bias"], "critic_2_weight": g["actor_2_weight"],
                "critic_2_bias": g["actor_2_bias"], "critic_4_weight": g["actor_4_weight"][:1], "critic_4_bias": g["actor_4_bias"][:1]}
    env.load_actor(g, action_std=sigma, critic=crit)
    kw = dict(record_obs=record, record_actions=record, record_logprob=record, record_reward=record, record_done=record, record_values=critic)
    ms = time_ms(lambda: env.policy_rollout(K, **kw), iters, warm=2)
    s = env.stats()
    print("policy_rollout%s N=%d K=%d sigma=%.2f record=%d   %8.2f us/step  %.3e env-steps/s   (%.1f ms per %d-step rollout; solved %.3f, mean len %.0f)"
          % ("+critic" if critic else "", N, K, sigma, record, ms * 1e3 / K, N * K / ms * 1e3, ms, K, s["solved_frac"], s["mean_length"]), flush=True)


if __name__ == "__main__":
    which = sys.argv[1:] or ["all"]
    if "prof" in which:                                              # short run for ncu
        case_step(async_reset=True, T=5, iters=20)
    if "profsensor" in which:
        case_step(async_reset=True, T=5, iters=20, sensor_noise=True)
    if "profrollout" in which:
        case_rollout(n=10 ** 9, K=32, iters=2)
    if "ppograd" in which:
        case_ppo_grad()
        case_ppo_grad(N=1 << 20, K=128, iters=2)
    if "profppograd" in which:
        case_ppo_grad(N=1 << 15, K=64, iters=1, warm=1)
    if "f64" in which:
        case_step(N=1 << 16, iters=100, precision="f64", integrator="rk45", async_reset=True, T=5)
        case_step(N=1 << 18, iters=50, precision="f64", integrator="rk45", async_reset=True, T=5)
        case_step(N=4096, iters=200, precision="f64", integrator="rk45", async_reset=True, T=5)
        case_rollout(N=1 << 16, K=32, iters=4, precision="f64", integrator="rk45", async_reset=True, T=5)
    if "rolloutab" in which:                                         # fused rollouts with resets: re-sampler and action-fetch A/B
        case_rollout(K=32, iters=8, async_reset=True, T=5, preroll=6)
        case_rollout(K=32, iters=8, async_reset=True, T=5, preroll=6, actions=True, record=True)
        case_rollout(K=32, iters=8, async_reset=True, T=5, preroll=6, sensor_noise=True)
        case_rollout(K=32, iters=8, async_reset=True, T=5, preroll=6, sensor_noise=True, record=True)
        case_rollout(K=32, iters=8, async_reset=True, T=5, preroll=6, sensor_noise=True, actions=True, record=True)
    if "strict" in which:                                            # handles that cannot use the per-warp kernels: strict resets, AUX rows
        case_step(auto_reset=True, T=5)
        case_step(N=1 << 18, auto_reset=True, T=5)
    if "proff64" in which:
        case_step(N=1 << 16, iters=5, precision="f64", integrator="rk45", async_reset=True, T=5)
    if "profpolicy" in which:
        case_policy(N=1 << 18, K=32, iters=1)
    if "policycritic" in which:
        case_policy(critic=True)
        case_policy()
    if "profpolicycritic" in which:
        case_policy(N=1 << 18, K=32, iters=1, critic=True)
    if "policy" in which:
        case_policy()
        case_policy(record=False)
        case_policy(N=1 << 18)
    if "sensor" in which:
        case_step(async_reset=True, T=5)
        case_step(async_reset=True, T=5, sensor_noise=True)
        case_step(n=10 ** 9, act_scale=0.05, sensor_noise=True)
    if "step" in which or "all" in which:
        case_step(n=10 ** 9, act_scale=0.05)                         # pure step, nobody finishes
        case_step(async_reset=True, T=5)
        case_step(async_reset=True, T=1)
        case_step(auto_reset=True, T=5)
        case_step(async_reset=True, T=5, substeps=4)
        case_step(N=1 << 21, async_reset=True, T=5)
        case_step(N=1 << 18, async_reset=True, T=5)
        case_step(N=1 << 16, iters=100, precision="f64", integrator="rk45", async_reset=True, T=5)
    if "sensorrollout" in which:
        case_rollout(async_reset=True, T=5, sensor_noise=True)
        case_rollout(async_reset=True, T=5, sensor_noise=True, record=True)
        case_rollout(async_reset=True, T=5, sensor_noise=True, K=128, iters=4)
        case_rollout(async_reset=True, T=5)
        case_rollout(async_reset=True, T=5, record=True)
    if "profrollouts8" in which:                                     # the fused rollout at 8 RK4 sub-intervals, no resets: the FMA-roofline point
        case_rollout(n=10 ** 9, K=8, iters=2, substeps=8)
    if "profsensorrollout" in which:
        case_rollout(async_reset=True, T=5, sensor_noise=True, K=32, iters=2, preroll=6)
    if "rollout" in which or "all" in which:
        case_rollout(n=10 ** 9)
        case_rollout(async_reset=True, T=5)
        case_rollout(async_reset=True, T=5, K=128)
