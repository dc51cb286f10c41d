"""TEST / BENCH INFRASTRUCTURE ONLY — the reference's OWN quad.step timed on the host cores (bench.py's CPU legs).

The unmodified reference (bytecode build oracle/_ref, oracle/build_ref.py — it travels to the GPU box) is run the way
BASELINE.md section 4 plans it: P = all host cores worker processes, each stepping its own `quad` objects (direct control,
T = 5, actions U(-1,1)^4, `reset()` when an episode ends: the episode mix of the GPU workload's auto-reset), env-steps/s
aggregated over the workers.  Nothing in the product package imports this module."""
import multiprocessing as mp
import os
import time


def host_cores() -> int:
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


def _worker(rank, m_envs, steps, warmup, T, start_evt, q):
    import contextlib
    import io
    import sys
    import numpy as np
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if root not in sys.path:
        sys.path.insert(0, root)
    os.environ["OMP_NUM_THREADS"] = "1"
    try:
        from oracle import ref_runtime as rr
        import warnings
        warnings.simplefilter("ignore")
        ref = rr.import_reference_env()
        np.random.seed(1000 + rank)
        rng = np.random.default_rng(rank)
        with contextlib.redirect_stdout(io.StringIO()):
            envs = [ref.quad(0.01, 1000, training=True, euler=0, direct_control=1, T=T) for _ in range(m_envs)]
            for e in envs:
                e.reset()

            def sweep():
                n_reset = 0
                for e in envs:
                    _, _, done = e.step(rng.uniform(-1, 1, 4))
                    if done:
                        e.reset()
                        n_reset += 1
                return n_reset

            for _ in range(warmup):
                sweep()
            q.put(("ready", rank))
            start_evt.wait()
            t0 = time.perf_counter()
            resets = 0
            for _ in range(steps):
                resets += sweep()
            dt = time.perf_counter() - t0
        q.put(("done", rank, dt, resets))
    except Exception as ex:                                   # surface the failure instead of hanging the parent
        q.put(("error", rank, repr(ex)))


def run(m_envs_per_worker, steps, warmup, T=5, workers=0):
    """Each of P workers steps m envs `steps` times (after `warmup` untimed sweeps).  Returns dict(rate, seconds, cores,
    n_envs, resets): rate = P*m*steps / (slowest worker's time)."""
    P = workers or host_cores()
    ctx = mp.get_context("fork")
    q, start = ctx.Queue(), ctx.Event()
    procs = [ctx.Process(target=_worker, args=(r, m_envs_per_worker, steps, warmup, T, start, q), daemon=True) for r in range(P)]
    for p in procs:
        p.start()
    ready = 0
    results = []
    while ready < P:
        msg = q.get(timeout=900)
        if msg[0] == "error":
            raise RuntimeError("reference worker %d failed: %s" % (msg[1], msg[2]))
        ready += 1
    start.set()
    while len(results) < P:
        msg = q.get(timeout=1800)
        if msg[0] == "error":
            raise RuntimeError("reference worker %d failed: %s" % (msg[1], msg[2]))
        results.append(msg)
    for p in procs:
        p.join(timeout=30)
    dt = max(r[2] for r in results)
    n = P * m_envs_per_worker
    return dict(rate=n * steps / dt, seconds=dt, cores=P, n_envs=n, resets=sum(r[3] for r in results))


def available() -> bool:
    from oracle import ref_runtime as rr
    return rr.available()


def run_isolated(m_envs_per_worker, steps, warmup, T=5, workers=0, timeout=1200):
    """run() in a FRESH interpreter: bench.py calls this after CUDA / NCCL are initialised in its own process, where forking
    worker processes is not safe (threads of the CUDA runtime hold locks a forked child would inherit)."""
    import json
    import subprocess
    import sys
    cmd = [sys.executable, os.path.abspath(__file__), str(m_envs_per_worker), str(steps), str(warmup), str(T), str(workers)]
    env = dict(os.environ)
    env.pop("OMP_NUM_THREADS", None)
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env)
    if r.returncode != 0:
        raise RuntimeError("reference run failed: " + r.stderr[-2000:])
    return json.loads(r.stdout.strip().splitlines()[-1])


if __name__ == "__main__":
    import json
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    a = [int(x) for x in sys.argv[1:6]]
    print(json.dumps(run(a[0], a[1], a[2], a[3], a[4])))
