"""CPU tests of the N>1 host logic: contiguous global-id sharding and the statistics all-reduce over a
world_size-2 gloo group (the same code path bench.py runs over NCCL)."""
import os
import socket
import subprocess
import sys
import textwrap

import numpy as np

from conftest import ROOT
from autonomous_quadrotor_environment_b200.sharding import shard_range
from oracle import quad_oracle as qo


def test_shard_range_partitions_exactly():
    for n_total in (1, 7, 4096, 16_777_216, 1_000_003):
        for world in (1, 2, 3, 8):
            if n_total < world:
                continue
            got, nxt = 0, 0
            for r in range(world):
                n, off = shard_range(n_total, r, world)
                assert off == nxt and n >= n_total // world
                nxt += n; got += n
            assert got == n_total
    assert shard_range(16_777_216, 3, 8) == (2_097_152, 3 * 2_097_152)


def test_sharded_oracle_equals_whole():
    """Philox keyed by global env id: two shards sample exactly the states the whole batch samples."""
    whole, _ = qo.sample_reset_state(9, np.arange(1000), 4)
    parts = []
    for r in range(2):
        n, off = shard_range(1000, r, 2)
        parts.append(qo.sample_reset_state(9, np.arange(n) + off, 4)[0])
    assert np.array_equal(whole, np.concatenate(parts))


WORKER = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, %r)
    import torch, torch.distributed as dist
    from autonomous_quadrotor_environment_b200.sharding import init_distributed, shard_range, allreduce_stats, stats_dict
    rank, world, local = init_distributed("gloo")
    assert world == 2
    n, off = shard_range(1001, rank, world)
    # each rank's local statistics vector (what the step kernels accumulate on the device)
    s = torch.tensor([10.0 * (rank + 1), 100.0 + rank, n, rank, 1 - rank, 0.0, 0.5, float(n * 3)], dtype=torch.float64)
    allreduce_stats(s)
    d = stats_dict(s)
    assert d["n_episodes"] == 1001 and d["n_steps"] == 3003 and d["sum_return"] == 30.0, d
    assert d["n_solved"] == 1 and d["n_broken"] == 1
    w = allreduce_stats(s.clone(), async_op=True); w.wait()
    dist.barrier()
    if rank == 0:
        print("GLOO_OK", off, n)
    dist.destroy_process_group()
""")


def test_stats_allreduce_world2_gloo(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(WORKER % ROOT)
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", LOCAL_RANK=str(r), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port), CUDA_VISIBLE_DEVICES="")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT))
    outs = [p.communicate(timeout=240)[0].decode() for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert "GLOO_OK 0 501" in outs[0]
