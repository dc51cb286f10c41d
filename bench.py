#!/usr/bin/env python
"""bench.py — env-steps/sec of the batched quadrotor hot path (BASELINE.json metric) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W]          # our arm (CUDA)
    python bench.py --impl reference [--gpus N] ...              # the reference's algorithm on the host cores

A "step" is one lock-step env step (quad.step) of every environment of the workload:
  N=1  : BASELINE.json configs[2] — 1,048,576 envs, FP32 RK4, auto-reset, T=5, random actions read from HBM
  N>1  : BASELINE.json configs[3] — 2,097,152 envs per GPU sharded by global env id (16,777,216 at N=8),
         with an NCCL all-reduce of the episode statistics every 128 steps.
`value` = env-steps/s with the actions already resident in HBM; `e2e` = the same through the C-ABI entry
point qs_step_host with pinned HOST buffers (H2D of actions, D2H of obs/reward/done inside the timed region).
One JSON line is printed by rank 0.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "env-steps/sec (RK4, FP32)"
UNIT = "env-steps/s"
ALGO_BYTES_PER_ENV_STEP = 181          # state 13 in + 13 out, action 4 in, obs 14 + reward 1 out (fp32), done 1 B
FLOPS_PER_ENV_STEP = lambda S: 701 * S + 180


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20000)
    ap.add_argument("--warmup", type=int, default=200)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--envs-per-gpu", type=int, default=0, help="0 = 1,048,576 at N=1, 2,097,152 at N>1")
    ap.add_argument("--substeps", type=int, default=1)
    ap.add_argument("--T", type=int, default=5)
    ap.add_argument("--e2e-steps", type=int, default=100)
    ap.add_argument("--cpu-seconds", type=float, default=10.0, help="budget of the cpu_baseline leg (N=1, rank 0)")
    ap.add_argument("--ref-seconds", type=float, default=90.0, help="--impl reference: CPU time the bounded sample is sized for")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--sensor-noise", type=int, default=1)
    ap.add_argument("--variant-steps", type=int, default=2000, help="timed steps of the sensor_noise=0 variant (0 = skip)")
    ap.add_argument("--workload", default="step", choices=["step", "policy"],
                    help="step = BASELINE.json configs[2]/[3] (the headline, default); policy = configs[4]: PPO rollout, "
                         "1M envs x 128-step horizon per launch with the actor MLP fused in on tcgen05 (one 'step' = one rollout)")
    ap.add_argument("--horizon", type=int, default=128)
    return ap.parse_args()


# ----------------------------------------------------------------------------------------------------------
# CPU legs (the only places bench.py may execute oracle/)
# ----------------------------------------------------------------------------------------------------------
def cpu_reference_run(n_envs, steps, warmup, T, threads=0):
    """The reference's algorithm (SciPy-RK45 replica + drone_eq + done/reward, oracle/quad_oracle.c) on the host."""
    import numpy as np
    from oracle.c_oracle import COracle
    from oracle import quad_oracle as qo
    if threads == 0:                                   # all host cores (torchrun pins OMP_NUM_THREADS=1: ask explicitly)
        threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    env = COracle(n_envs, 0.01, 10 ** 9, training=False, direct_control=1, T=T, integrator="rk45", threads=threads)
    cores = threads
    init, _ = qo.sample_reset_state(0, np.arange(n_envs), 0)
    env.reset(init)
    rng = np.random.default_rng(0)
    acts = [rng.uniform(-1, 1, (n_envs, 4)) for _ in range(4)]
    for w in range(warmup):
        env.step(acts[w % 4])
    t0 = time.perf_counter()
    for k in range(steps):
        env.step(acts[k % 4])
        if k % 16 == 15:                               # keep envs alive: re-seed the ones that left the box
            bad = env.done.astype(bool) | ~np.isfinite(env.state).all(axis=1)
            if bad.any():
                env.state[bad] = init[bad]; env.flags[bad] = 0; env.step_i[bad] = 0
    dt = time.perf_counter() - t0
    return n_envs * steps / dt, dt, cores


def python_reference_run(seconds, T, steps=None, warmup=None):
    """The reference's OWN quad.step (unmodified; bytecode build oracle/_ref) on all host cores: P worker processes x m envs each,
    U(-1,1) actions, reset() on done.  Either a time budget (seconds) or an explicit (steps, warmup) sweep count."""
    from oracle import cpu_reference as cr
    probe = cr.run_isolated(2, 20, 2, T)                         # ~ 0.1 s per worker: per-core rate incl. resets
    per_core = probe["rate"] / probe["cores"]
    if steps is None:
        m, warmup = 4, 2
        steps = max(10, int(per_core * seconds / m))
    else:
        m = int(max(1, min(4096, per_core * seconds / max(1, steps + warmup))))
    r = cr.run_isolated(m, steps, warmup, T)
    r["m"], r["steps"] = m, steps
    return r


def cpu_baseline(seconds, T):
    """cpu_baseline of the default line: the reference itself when its bytecode build travelled with the tree (kind "reference"),
    else the C port; the C port and the vectorised NumPy oracle are always reported as labelled extras."""
    import numpy as np
    from oracle import cpu_reference as cr
    from oracle import quad_oracle as qo
    n = 8192
    rate, _, cores = cpu_reference_run(n, 4, 1, T)
    steps = max(4, int(rate * min(seconds, 5.0) / n))
    rate, dt, cores = cpu_reference_run(n, steps, 2, T)
    port = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "%d envs x %d steps (%.1f s) of the FP64 SciPy-RK45-replica step (oracle/quad_oracle.c, OpenMP)" % (n, steps, dt)}
    nv = 4096                                                    # "best NumPy": the vectorised FP64 oracle, one process
    ora = qo.BatchQuadOracle(nv, 0.01, 10 ** 9, training=False, direct_control=1, T=T, integrator="rk45")
    init, _ = qo.sample_reset_state(0, np.arange(nv), 0)
    ora.reset(init)
    rng = np.random.default_rng(0)
    t0 = time.perf_counter()
    k = 0
    while time.perf_counter() - t0 < 3.0:
        ora.step(rng.uniform(-0.05, 0.05, (nv, 4)))
        k += 1
    numpy_vec = {"value": nv * k / (time.perf_counter() - t0), "unit": UNIT, "cores": 1, "kind": "port",
                 "sample": "%d envs x %d steps of oracle/quad_oracle.py (vectorised NumPy FP64 RK45 replica), one process" % (nv, k)}
    if cr.available():
        r = python_reference_run(seconds, T)
        return {"value": r["rate"], "unit": UNIT, "cores": r["cores"], "kind": "reference",
                "sample": "the reference's own quad.step (unmodified, bytecode build oracle/_ref): %d worker processes x %d envs x %d "
                          "steps (%.1f s), direct control, T=%d, U(-1,1) actions, reset() on done (%d resets)"
                          % (r["cores"], r["m"], r["steps"], r["seconds"], T, r["resets"]),
                "extra": {"c_port_all_cores": port, "numpy_vectorised_one_process": numpy_vec}}
    port["extra"] = {"numpy_vectorised_one_process": numpy_vec}
    port["sample"] += "; oracle/_ref (the Python reference itself) was not shipped with this tree"
    return port


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import cpu_reference as cr
    n_gpu_envs = args.envs_per_gpu or (1 << 20 if args.gpus == 1 else 1 << 21)
    if cr.available():
        r = python_reference_run(args.ref_seconds, args.T, steps=args.steps, warmup=args.warmup)
        rate, dt, cores, n = r["rate"], r["seconds"], r["cores"], r["n_envs"]
        kind, dtype = "reference", "f64"
        how = ("the reference's own quad.step (unmodified Python/NumPy/SciPy, bytecode build oracle/_ref), %d worker processes x %d "
               "envs, U(-1,1) actions, reset() on done" % (cores, r["m"]))
    else:
        probe_n = 4096
        rate, _, cores = cpu_reference_run(probe_n, 4, 1, args.T)
        n = int(max(256, min(65536, rate * args.ref_seconds / max(1, args.steps + args.warmup))))
        rate, dt, cores = cpu_reference_run(n, args.steps, args.warmup, args.T)
        kind, dtype = "port", "f64"
        how = "reference algorithm (FP64 RK45 rtol 1e-3), C port oracle/quad_oracle.c with OpenMP (oracle/_ref not shipped with this tree)"
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": dtype, "data": "synthetic",
        "config": {"workload": "bounded sample (%d envs per step) of the %d-env lock-step workload on the host cores: %s"
                               % (n, n_gpu_envs * args.gpus, how), "T": args.T},
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": "%d envs x %d steps (%.1f s)" % (n, args.steps, dt)},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit_line(line)


# ----------------------------------------------------------------------------------------------------------
# clocks sampler (B200_PROFILING.md "clocks DURING the timed region")
# ----------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.rows.append((time.perf_counter(), ln.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, ln in self.rows:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                if t0 <= ts <= t1 + 0.1:
                    sm.append(float(f[0]))
                smax = float(f[1])
            except ValueError:
                continue
            if t0 <= ts <= t1 + 0.1:
                for nm, v in zip(names, f[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------------
def policy_measure(args, dev, rank, world, steps, warm):
    """BASELINE.json configs[4]: fused actor-MLP rollout; one bench 'step' is one K-step launch.  Returns the measurement dict
    (value = env-steps/s of this rank's shard x world when the caller all-reduces the time)."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from autonomous_quadrotor_environment_b200 import BatchedQuad
    N = args.envs_per_gpu or (1 << 20)
    K = args.horizon
    env = BatchedQuad(N, 0.01, 1000, training=True, direct_control=1, T=args.T, precision="f32", async_reset=True, seed=0,
                      env_id_offset=rank * N, device=dev)
    env.reset()
    env.load_actor(dict(np.load(os.path.join(ROOT, "tests", "golden", "actor_128.npz"))), action_std=0.1)
    for _ in range(warm):
        rec = env.policy_rollout(K)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        rec = env.policy_rollout(K)                        # records actions, log-probs, rewards, dones: (K,*,N) buffers in HBM
    e1.record()
    torch.cuda.synchronize(dev)
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    value = N * world * K * steps / (ms * 1e-3)
    # e2e: the per-iteration statistic a PPO driver reads back (mean reward of the rollout) is reduced on the device and read
    mean_r = float(rec["reward"].mean().item())            # untimed: loads torch's reduction kernel (lazy module loading)
    torch.cuda.synchronize(dev)
    e2e_iters = 3
    te0 = time.perf_counter()
    for _ in range(e2e_iters):
        rec = env.policy_rollout(K)
        mean_r = float(rec["reward"].mean().item())
    torch.cuda.synchronize(dev)
    e2e = N * world * K * e2e_iters / (time.perf_counter() - te0)
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.isfile(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    tf_peak = float(peaks.get("bf16_tflops_sustained", 1367.3))
    mlp_flops = 2 * (80 * 128 + 128 * 128 + 128 * 16)          # per env-step as issued (K padded 75->80, N 4->16)
    ach = mlp_flops * N * K * steps / (ms * 1e-3) / 1e12
    return {"value": value, "unit": UNIT, "steps": steps, "warmup": warm, "ms_per_step": ms / steps, "envs_per_gpu": N, "horizon": K,
            "workload": "PPO rollout: %d envs/GPU x %d-step horizon per launch, actor MLP 75-128-128-4 (BF16 tcgen05, FP32 accumulate) + "
                        "Normal(sigma=0.1) sampling + FP32 RK4 quad.step fused, async auto-reset, T=%d; records actions/log-probs/"
                        "rewards/dones" % (N, K, args.T),
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 4,
                    "api": "BatchedQuad.policy_rollout + mean reward read back (actions are produced on the device by the fused actor: no per-step host input exists)"},
            "roofline": {"bound": "tensor", "achieved": ach, "peak": tf_peak, "unit": "TFLOP/s", "frac": ach / tf_peak,
                         "traffic": None, "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained", "kernel": "policy_rollout_kernel",
                         "note": "the kernel is bound by MUFU (256 tanh per env-step) and its serial MMA->epilogue->dynamics chain, not by the tensor pipe"},
            "stats": env.stats(all_reduce=False), "mean_reward_last_rollout": mean_r}


def ppo_iteration_measure(args, dev, iters=2):
    """One PPO iteration (ppo.py:125-209) on 1M envs x 128 steps, entirely on the device: device time of collect and update, and the
    tensor-core rate of the update's gradient kernel in the FLOPs of its products as issued."""
    import torch
    from autonomous_quadrotor_environment_b200 import BatchedQuad
    from autonomous_quadrotor_environment_b200.ppo import BatchedPPO
    N, K = args.envs_per_gpu or (1 << 20), args.horizon
    env = BatchedQuad(N, 0.01, 1000, training=True, direct_control=1, T=args.T, precision="f32", async_reset=True, seed=0, device=dev)
    env.reset()
    ppo = BatchedPPO(env, hidden=128, K_epochs=10, seed=0)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    t_collect, t_update, losses = [], [], None
    for it in range(iters + 1):                                 # the first iteration is the warm-up
        ev[0].record()
        batch = ppo.collect(K)
        ev[1].record()
        losses = ppo.update(batch)
        ev[2].record()
        torch.cuda.synchronize(dev)
        if it > 0:
            t_collect.append(ev[0].elapsed_time(ev[1])); t_update.append(ev[1].elapsed_time(ev[2]))
        del batch
    mc, mu = sum(t_collect) / len(t_collect), sum(t_update) / len(t_update)
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.isfile(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    tf_peak = float(peaks.get("bf16_tflops_sustained", 1367.3))
    # per sample and network as issued: forward K = 80 + 144 against N = 128; backward dW3 (N 16) + dH1 (N 128) + [dW2|db2] (N 144) + dW1 (N 80), K = 128
    flops = 2.0 * 128 * ((80 + 144) + (16 + 128 + 144 + 80)) * 2 * ppo.K_epochs * float(N) * K
    ach = flops / (mu * 1e-3) / 1e12
    return {"envs_per_gpu": N, "horizon": K, "K_epochs": ppo.K_epochs, "collect_ms": mc, "update_ms": mu,
            "value": N * K / ((mc + mu) * 1e-3), "unit": "env-steps/s collected AND trained on (one PPO iteration)",
            "losses_last_iteration": [losses[0], losses[-1]],
            "roofline": {"bound": "tensor", "achieved": ach, "peak": tf_peak, "unit": "TFLOP/s", "frac": ach / tf_peak, "traffic": None,
                         "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained", "kernel": "ppo_grad_kernel<actor|critic>",
                         "note": "update_ms covers 2 x K_epochs gradient launches + Adam; the kernel is bound by its four MMA -> epilogue "
                                 "round trips per (128-sample tile, step), profiles/r02_prof_ppo_grad.txt"}}


def run_policy_workload(args):
    """--workload policy: configs[4] as the headline line.  Same metric (env-steps/s)."""
    import torch
    from autonomous_quadrotor_environment_b200.sharding import init_distributed
    import torch.distributed as dist
    rank, world, local = init_distributed()
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    steps, warm = min(args.steps, 20), max(3, min(args.warmup, 3))
    sampler = ClockSampler(local) if rank == 0 else None
    t0 = time.perf_counter()
    m = policy_measure(args, dev, rank, world, steps, warm)
    t1 = time.perf_counter()
    if rank == 0:
        line = {"metric": METRIC, "value": m["value"], "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warm,
                "ms_per_step": m["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic",
                "config": {"workload": m["workload"], "envs_per_gpu": m["envs_per_gpu"], "horizon": m["horizon"],
                           "l2": "rollout buffers %.1f GB/launch >> 126 MB L2" % (m["envs_per_gpu"] * m["horizon"] * 37 / 1e9)},
                "e2e": m["e2e"], "gpu_launches": steps, "clocks": sampler.stop(t0, t1) if sampler else None,
                "roofline": m["roofline"], "stats": m["stats"], "mean_reward_last_rollout": m["mean_reward_last_rollout"]}
        emit_line(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


_REAL_STDOUT = None


def protect_stdout():
    """The contract is ONE JSON line on stdout.  Native libraries write there too (NCCL prints its version banner on stdout at
    NCCL_DEBUG=VERSION and WARN), so file descriptor 1 is pointed at stderr for the duration of the run and the JSON line is
    written to the saved descriptor."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit_line(line):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    args = parse()
    protect_stdout()
    # NCCL prints its version banner on STDOUT when its debug level is VERSION (the level may also come from /etc/nccl.conf, which
    # the environment variable overrides): keep stdout to the one JSON line unless the caller asked for NCCL's own logging
    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
        os.environ["NCCL_DEBUG"] = "WARN"
    if args.impl == "reference":
        run_reference_arm(args)
        return
    if args.workload == "policy":
        run_policy_workload(args)
        return
    import torch
    from autonomous_quadrotor_environment_b200 import BatchedQuad, _lib as L
    from autonomous_quadrotor_environment_b200.sharding import init_distributed, allreduce_stats
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    rank, world, local = init_distributed()
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, world))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    N = args.envs_per_gpu or (1 << 20 if world == 1 else 1 << 21)
    env = BatchedQuad(N, 0.01, 1000, training=True, direct_control=1, T=args.T, precision="f32", integrator="rk4",
                      substeps=args.substeps, async_reset=True, sensor_noise=bool(args.sensor_noise), seed=0,
                      env_id_offset=rank * N, device=dev)
    env.reset()
    P = 16                                               # action pool: 16 x (4,N) fp32 = 256 MB at 1M envs (> 126 MB L2)
    g = torch.Generator(device=dev); g.manual_seed(1234 + rank)
    acts = [(torch.rand(4, N, device=dev, generator=g) * 2 - 1).contiguous() for _ in range(P)]
    ptrs = [C.c_void_p(a.data_ptr()) for a in acts]
    lib, h = env.lib, env._h
    stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    stats_buf = torch.zeros(8, dtype=torch.float64, device=dev)
    main_stream = torch.cuda.current_stream(dev)
    side = torch.cuda.Stream(dev)                        # the statistics exchange never sits between two step kernels
    period = max(1, min(128, args.steps))                # BASELINE.json configs[3]: one all-reduce of the episode statistics per 128 steps
    n_exchanges = [0]

    def exchange_stats():
        """The path's only collective: snapshot the 8 accumulators and sum them over the ranks (NCCL), on the side stream."""
        ev = torch.cuda.Event()
        ev.record(main_stream)
        side.wait_event(ev)
        with torch.cuda.stream(side):
            stats_buf.copy_(env.stats_tensor())
            allreduce_stats(stats_buf)
        n_exchanges[0] += 1

    def step(k, exchange=True):
        rc = lib.qs_step(h, ptrs[k % P], None, None, None, None, stream)
        if rc != 0:
            L.check(rc)
        if exchange and (k + 1) % period == 0:
            exchange_stats()

    # the clock sampler starts here: nvidia-smi needs ~0.1-0.3 s to print its first row, and a 20-step timed region lasts 2 ms —
    # the GPU is under the same load from the pre-roll on, so the clocks reported are the samples from here to the end of the
    # timed region
    sampler = ClockSampler(torch.cuda.current_device() if "CUDA_VISIBLE_DEVICES" not in os.environ else
                           os.environ["CUDA_VISIBLE_DEVICES"].split(",")[local]) if rank == 0 else None
    t_load0 = time.perf_counter()
    # untimed pre-roll: right after reset() every env is at step 0 of its first episode; the reset rate (and with it the
    # kernel's re-sampling work) only becomes stationary once the first episodes have turned over (~50-step episodes under
    # random actions): roll until the episode count of a 32-step window stops changing by more than 2 %
    preroll, last = 0, None
    while preroll < 2000:
        before = float(env.stats_tensor()[2].item())
        for k in range(32):
            step(preroll + k, exchange=False)
        preroll += 32
        rate = float(env.stats_tensor()[2].item()) - before
        if last is not None and last > 0 and abs(rate - last) <= 0.02 * last and preroll >= 256:
            break
        last = rate
    for w in range(args.warmup):
        step(w, exchange=False)
    if sampler is not None:                              # keep the load up until the sampler has seen it at least twice (<= 1.5 s)
        w = 0
        while len(sampler.rows) < 2 and time.perf_counter() - t_load0 < 1.5:
            for k in range(64):
                step(w + k, exchange=False)
            w += 64
            torch.cuda.synchronize(dev)
    env.stats_tensor().zero_()                           # the statistics reported below are those of the timed region
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    ev0.record()
    for k in range(args.steps):
        step(k)
    if args.steps % period != 0:
        exchange_stats()                                 # the last iteration's exchange
    main_stream.wait_stream(side)                        # the timed region ends when the last collective has
    ev1.record()
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    t1 = time.perf_counter()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop(t_load0, t1) if sampler else None
    tmax = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms_max = float(tmax.item())
    total_envs = N * world
    value = total_envs * args.steps / (ms_max * 1e-3)
    from autonomous_quadrotor_environment_b200.sharding import stats_dict
    stats_all = stats_dict(stats_buf)                    # ALL-REDUCED statistics of the timed region (last exchange)
    stats_all["exchanges_in_timed_region"] = n_exchanges[0]
    stats_all["preroll_steps"] = preroll
    if stats_all["n_steps"] != float(total_envs) * args.steps:
        raise SystemExit("all-reduced n_steps %r != envs x ranks x steps %r" % (stats_all["n_steps"], float(total_envs) * args.steps))

    # ---- e2e: qs_step_host with pinned host buffers (H2D actions, D2H obs/reward/done inside the timed region)
    e2e_steps = max(1, args.e2e_steps)
    a_host = [torch.empty(4, N).uniform_(-1, 1).pin_memory() for _ in range(2)]
    obs_h = torch.empty(14, N).pin_memory(); rew_h = torch.empty(N).pin_memory()
    done_h = torch.empty(N, dtype=torch.uint8).pin_memory()
    for w in range(3):
        L.check(lib.qs_step_host(h, a_host[w % 2].data_ptr(), obs_h.data_ptr(), rew_h.data_ptr(), done_h.data_ptr(), stream))
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    te0 = time.perf_counter()
    for k in range(e2e_steps):
        L.check(lib.qs_step_host(h, a_host[k % 2].data_ptr(), obs_h.data_ptr(), rew_h.data_ptr(), done_h.data_ptr(), stream))
    torch.cuda.synchronize(dev)
    te = torch.tensor([time.perf_counter() - te0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = total_envs * e2e_steps / float(te.item())
    h2d = 4 * N * 4
    d2h = 14 * N * 4 + N * 4 + N

    # ---- roofs
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(peaks_path):
        hbm_peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    else:
        hbm_peak, peak_src = 6650.0, "B200_PROFILING.md fallback (of fallback)"
    kernel_ms = ms / args.steps                          # one launch of step_kernel per step, back to back on one stream
    sensor = int(bool(args.sensor_noise))
    kernel_name = {0: "step_kernel_direct<float,RK4,direct,sensor=%d>", 1: "step_kernel_tma<float,RK4,direct,sensor=%d>",
                   2: "step_kernel_warp<direct,sensor=%d>", 3: "step_kernel_pair<direct,sensor=%d>"}[env.step_loader] % sensor
    # algorithmic bytes per env-step: 181 B; the sensor model adds the 17 rows of its state that it reads AND
    # writes back (biases, drift rates, velocity_t0, position_t0, quaternion_t0, third column of R: 2 x 68 B) and the 14-float
    # sensed observation it writes (56 B) = 192 B.  The three acceleration_t0 rows of the 20-row QS_FIELD_SENSOR_STATE are a
    # write-only diagnostic (sensor.acceleration_t0, never read back) and are NOT counted, although the kernel moves them:
    # the DRAM traffic (`traffic`, 402 MB per launch) therefore sits above the 391 MB counted here
    algo_bytes = ALGO_BYTES_PER_ENV_STEP + (192 if sensor else 0)
    achieved_gbs = algo_bytes * N / (kernel_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.isfile(tp) and N == (1 << 20):
        try:
            traffic = json.load(open(tp)).get(kernel_name)
        except Exception:
            traffic = None
    probe_ms = C.c_float(0)
    blocks, threads, iters = 148 * 8, 256, 1 << 14
    L.check(lib.qs_fp32_peak_probe(blocks, threads, iters, C.byref(probe_ms), stream))
    fp32_peak = 2.0 * 8 * blocks * threads * iters / (probe_ms.value * 1e-3) / 1e12
    flops = FLOPS_PER_ENV_STEP(args.substeps) * N / (kernel_ms * 1e-3) / 1e12

    # ---- the same workload without the sensor model (the kernel every non-sensor user of qs_step runs), reported
    #      beside the headline so that both step kernels are measured by the same command (N=1 only)
    variants = {}
    if world == 1 and sensor and args.variant_steps > 0:
        env2 = BatchedQuad(N, 0.01, 1000, training=True, direct_control=1, T=args.T, precision="f32", integrator="rk4",
                           substeps=args.substeps, async_reset=True, sensor_noise=False, seed=0, env_id_offset=rank * N, device=dev)
        env2.reset()
        for w in range(max(3, args.warmup)):
            L.check(lib.qs_step(env2._h, ptrs[w % P], None, None, None, None, stream))
        torch.cuda.synchronize(dev)
        v0, v1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        v0.record()
        for k in range(args.variant_steps):
            L.check(lib.qs_step(env2._h, ptrs[k % P], None, None, None, None, stream))
        v1.record()
        torch.cuda.synchronize(dev)
        vms = v0.elapsed_time(v1) / args.variant_steps
        vname = {0: "step_kernel_direct<float,RK4,direct,sensor=0>", 1: "step_kernel_tma<float,RK4,direct,sensor=0>",
                 2: "step_kernel_warp<direct,sensor=0>", 3: "step_kernel_pair<direct,sensor=0>"}[env2.step_loader]
        vgbs = ALGO_BYTES_PER_ENV_STEP * N / (vms * 1e-3) / 1e9
        vtraffic = None
        try:
            vtraffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(vname) if N == (1 << 20) else None
        except Exception:
            pass
        variants["sensor_noise=0"] = {
            "value": N / (vms * 1e-3), "unit": UNIT, "steps": args.variant_steps, "kernel_ms": vms,
            "roofline": {"bound": "hbm", "achieved": vgbs, "peak": hbm_peak, "unit": "GB/s", "frac": vgbs / hbm_peak,
                         "traffic": vtraffic, "kernel": vname, "algorithmic_bytes_per_env_step": ALGO_BYTES_PER_ENV_STEP},
            "fp32_frac": FLOPS_PER_ENV_STEP(args.substeps) * N / (vms * 1e-3) / 1e12 / fp32_peak}
        # BASELINE.json configs[2], "actions from Philox in-kernel, so no HBM read" variant: K fused steps per launch with the env
        # state in registers (qs_rollout) — the kernel the metric's "% of FP32 FMA roofline" is about
        KR = 32
        for w in range(3):
            env2.rollout(KR)
        torch.cuda.synchronize(dev)
        reps = max(1, args.variant_steps // (KR * 4))
        v0.record()
        for k in range(reps):
            env2.rollout(KR)
        v1.record()
        torch.cuda.synchronize(dev)
        rms = v0.elapsed_time(v1) / (reps * KR)
        variants["rollout_philox_actions"] = {
            "value": N / (rms * 1e-3), "unit": UNIT, "steps": reps * KR, "ms_per_env_step_of_all_envs": rms,
            "kernel": "rollout_pair_kernel<direct>" if env2.step_loader == 3 else "rollout_kernel<float,RK4,direct>",
            "fp32": {"achieved_tflops": FLOPS_PER_ENV_STEP(args.substeps) * N / (rms * 1e-3) / 1e12, "peak_tflops_probe": fp32_peak,
                     "frac": FLOPS_PER_ENV_STEP(args.substeps) * N / (rms * 1e-3) / 1e12 / fp32_peak},
            "note": "K=%d steps per launch, async auto-reset, state in registers, no state/action traffic" % KR}
        del env2
        # with S >= 3 RK4 sub-intervals per env step the path is compute-bound whatever the launch shape (the
        # north star's "fused multi-substep RK4 kernel"); the same fused rollout at S = 4 and S = 8, without resets, shows how close
        # the integrator itself gets to the FP32 FMA roofline (701 S + 180 algorithmic FLOPs per env step)
        for S_ in (4, 8):
            env3 = BatchedQuad(N, 0.01, 10 ** 9, training=True, direct_control=1, T=args.T, precision="f32", integrator="rk4",
                               substeps=S_, sensor_noise=False, seed=0, env_id_offset=rank * N, device=dev)
            env3.reset()
            for w in range(2):
                env3.rollout(KR)
            torch.cuda.synchronize(dev)
            reps = max(1, args.variant_steps // (KR * 4 * S_))
            v0.record()
            for k in range(reps):
                env3.rollout(KR)
            v1.record()
            torch.cuda.synchronize(dev)
            rms4 = v0.elapsed_time(v1) / (reps * KR)
            tf = FLOPS_PER_ENV_STEP(S_) * N / (rms4 * 1e-3) / 1e12
            variants["rollout_philox_actions_substeps%d" % S_] = {
                "value": N / (rms4 * 1e-3), "unit": UNIT, "steps": reps * KR, "ms_per_env_step_of_all_envs": rms4,
                "kernel": "rollout_pair_kernel<direct>" if env3.step_loader == 3 else "rollout_kernel<float,RK4,direct>",
                "fp32": {"achieved_tflops": tf, "peak_tflops_probe": fp32_peak, "frac": tf / fp32_peak,
                         "flops_per_env_step": FLOPS_PER_ENV_STEP(S_)},
                "note": "K=%d steps per launch, %d RK4 sub-intervals per env step (h = %.3g ms), no resets" % (KR, S_, 10.0 / S_)}
            del env3
        # BASELINE.json configs[2] as ONE fused launch per K steps: the headline workload (sensor model, async auto-reset, actions read
        # from a device tensor, the sensed observation / reward / done of every step recorded) with the env AND sensor state held
        # on chip for the horizon (rollout_pair_kernel<direct,sensor>): 77 B per env-step instead of 373
        try:
            env4 = BatchedQuad(N, 0.01, 1000, training=True, direct_control=1, T=args.T, precision="f32", integrator="rk4",
                               substeps=args.substeps, async_reset=True, sensor_noise=True, seed=0, env_id_offset=rank * N, device=dev)
            env4.reset()
            KS = 32
            act4 = (torch.rand(KS, 4, N, device=dev, generator=g) * 2 - 1).contiguous()
            for w in range(8):                                         # 256 untimed steps: past the first episode turnover
                env4.rollout(KS, actions=act4)
            rec = None
            for w in range(3):                                         # the recorded buffers come out of torch's caching allocator:
                rec = env4.rollout(KS, actions=act4, record_sensed=True, record_reward=True, record_done=True)   # no cudaMalloc inside the timed region
            torch.cuda.synchronize(dev)
            reps = max(2, args.variant_steps // (KS * 8))
            v0.record()
            for k in range(reps):
                rec = env4.rollout(KS, actions=act4, record_sensed=True, record_reward=True, record_done=True)
            v1.record()
            torch.cuda.synchronize(dev)
            sms = v0.elapsed_time(v1) / (reps * KS)
            sb = 16 + 56 + 4 + 1
            variants["rollout_sensor_recorded"] = {
                "value": N / (sms * 1e-3), "unit": UNIT, "steps": reps * KS, "ms_per_env_step_of_all_envs": sms,
                "kernel": "rollout_pair_kernel<direct,sensor>",
                "roofline": {"bound": "hbm", "achieved": sb * N / (sms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                             "frac": sb * N / (sms * 1e-3) / 1e9 / hbm_peak, "traffic": None, "algorithmic_bytes_per_env_step": sb,
                             "note": "with the state on chip the stream is 77 B per env-step and HBM is no longer the limiter: the kernel is bound "
                                     "by instruction issue and dependency latency (2,600 warp-instructions per 64 env-steps at 8 warps per SM, "
                                     "FMA pipe 45 % busy; profiles/r02_prof_rollout_pair_sensor.txt)"},
                "fp32_frac": FLOPS_PER_ENV_STEP(args.substeps) * N / (sms * 1e-3) / 1e12 / fp32_peak,
                "done_frac_last_launch": float(((rec["done"] & 1) != 0).float().mean().item()),
                "note": "K=%d steps per launch of the headline workload (sensor model + async auto-reset), actions from a (K,4,N) device "
                        "tensor, sensed observation + reward + done recorded per step; env and sensor state stay on chip" % KS}
            del env4, act4, rec
        except Exception as ex:                                      # a variant must not take the headline line down with it
            variants["rollout_sensor_recorded"] = {"error": repr(ex)}

    if world == 1 and args.variant_steps > 0:
        # (a) the per-GPU shard of BASELINE.json configs[3] (2,097,152 envs) on ONE GPU: the like-for-like base point of the
        #     weak-scaling series the N > 1 runs of this script produce
        N2 = 1 << 21
        envb = BatchedQuad(N2, 0.01, 1000, training=True, direct_control=1, T=args.T, precision="f32", integrator="rk4",
                           substeps=args.substeps, async_reset=True, sensor_noise=bool(args.sensor_noise), seed=0, device=dev)
        envb.reset()
        actsb = [(torch.rand(4, N2, device=dev, generator=g) * 2 - 1).contiguous() for _ in range(8)]
        for w in range(400):
            L.check(lib.qs_step(envb._h, C.c_void_p(actsb[w % 8].data_ptr()), None, None, None, None, stream))
        torch.cuda.synchronize(dev)
        b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        nb = max(100, args.variant_steps // 4)
        b0.record()
        for k in range(nb):
            L.check(lib.qs_step(envb._h, C.c_void_p(actsb[k % 8].data_ptr()), None, None, None, None, stream))
        b1.record()
        torch.cuda.synchronize(dev)
        bms = b0.elapsed_time(b1) / nb
        variants["envs_per_gpu=2097152"] = {
            "value": N2 / (bms * 1e-3), "unit": UNIT, "steps": nb, "kernel_ms": bms,
            "note": "same workload as the headline at the per-GPU shard size of the N > 1 runs (400 untimed pre-roll steps)",
            "roofline_frac": algo_bytes * N2 / (bms * 1e-3) / 1e9 / hbm_peak}
        del envb, actsb
        # (b) BASELINE.json configs[4]: PPO rollout, 1M envs x 128-step horizon per launch, actor MLP fused in on tcgen05
        try:
            variants["policy_rollout_1Mx128"] = policy_measure(args, dev, 0, 1, steps=4, warm=2)
        except Exception as ex:                                  # a variant must not take the headline line down with it
            variants["policy_rollout_1Mx128"] = {"error": repr(ex)}

        # (b') BASELINE.json configs[0]: ONE env through the reference's own class API — the drop-in quad (N = 1 handle, FP64 + SciPy-RK45
        #      replica; host action in, host observation / attributes out on every call): wall time per quad.step
        try:
            import numpy as np
            from autonomous_quadrotor_environment_b200.quadrotor_env import quad as dropin_quad
            q1 = dropin_quad(0.01, 10 ** 6, training=False, euler=0, direct_control=1, T=1, clipped=True, verbose=False)
            q1.seed(1); q1.reset()
            for k in range(50):
                q1.step(np.zeros(4))
            ts = time.perf_counter()
            K1 = 500
            for k in range(K1):
                q1.step(0.02 * np.sin(0.1 * k + np.arange(4)))
            us = (time.perf_counter() - ts) / K1 * 1e6
            variants["single_env_dropin"] = {
                "value": 1e6 / us, "unit": UNIT, "us_per_step": us, "steps": K1,
                "note": "compat quad.step of one env (FP64, RK45 replica; kernel launch + host round trip of the whole attribute surface "
                        "per call); the reference's own quad.step takes 1.48 ms on one core (BASELINE.md)"}
            del q1
        except Exception as ex:
            variants["single_env_dropin"] = {"error": repr(ex)}

        # (c) SURVEY.md 8(f)1: one PPO iteration on the same 1M x 128 rollout — collect (fused actor + critic rollout, GAE) and the
        #     K_epochs = 10 network update on qs_ppo_grad / qs_adam_step (forward + backward on tcgen05)
        try:
            variants["ppo_iteration_1Mx128"] = ppo_iteration_measure(args, dev)
        except Exception as ex:
            variants["ppo_iteration_1Mx128"] = {"error": repr(ex)}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "%d envs/GPU x %d GPU lock-step quad.step, FP32 RK4 x%d substeps, auto-reset (async warm-up), T=%d, "
                                   "sensor_noise=%d, U(-1,1) actions read from a %d-buffer HBM pool"
                                   % (N, world, args.substeps, args.T, int(bool(args.sensor_noise)), P),
                       "envs_per_gpu": N, "substeps": args.substeps, "T": args.T,
                       "l2": "working set %.0f MB/step > 126 MB L2 (state+obs rows %d B/env + rotating action pool)"
                             % (N * 200 / 1e6, 200)},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "api": "qs_step_host (pinned host buffers)"},
            "gpu_launches": args.steps,
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s",
                         "frac": achieved_gbs / hbm_peak, "traffic": traffic, "peak_source": peak_src,
                         "kernel": kernel_name, "algorithmic_bytes_per_env_step": algo_bytes,
                         "kernel_ms": kernel_ms,
                         "note": "HBM is the roof of a one-step launch (SURVEY 8(d)); the kernel itself is bound by instruction issue and "
                                 "the latency of its dependency chains: 2,660 warp-instructions per 64 env-steps with the sensor model "
                                 "(two envs per lane on FFMA2/FMUL2), 8 warps per SM at 210 registers "
                                 "(profiles/r02_prof_step_pair_sensor.txt; DESIGN.md lists the variants measured against it)"},
            "fp32": {"achieved_tflops": flops, "peak_tflops_probe": fp32_peak, "frac": flops / fp32_peak,
                     "flops_per_env_step": FLOPS_PER_ENV_STEP(args.substeps),
                     "note": "algorithmic FLOPs (SURVEY.md 8(d)) vs an in-run dependent-FFMA probe"},
            "stats": stats_all,
        }
        if variants:
            line["variants"] = variants
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args.cpu_seconds, args.T)
        emit_line(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
