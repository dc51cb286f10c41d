"""Multi-GPU check of the on-device PPO iteration (developer tool; run under torchrun on >= 2 GPUs):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/ppo_ddp_check.py
Every rank owns a contiguous shard of the envs (env_id_offset), the flat gradient is all-reduced over NCCL once per epoch
(qs_ppo_grad accumulates, qs_adam_step applies).  Checks: (1) the ranks stay in lock-step (identical parameters);
(2) the sharded run follows the single-GPU run of all envs (rank 0 repeats it alone): same loss trajectory and parameters up to the
FP32 summation order of the gradient."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from autonomous_quadrotor_environment_b200 import BatchedQuad
from autonomous_quadrotor_environment_b200.ppo import BatchedPPO
from autonomous_quadrotor_environment_b200.sharding import init_distributed

rank, world, local = init_distributed()
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
N_total, K, iters = 16384, 32, 3
n = N_total // world


def run(n_envs, offset):
    env = BatchedQuad(n_envs, 0.01, 1000, training=True, direct_control=1, T=5, precision="f32", async_reset=True, seed=3,
                      env_id_offset=offset, device=dev)
    env.reset()
    ppo = BatchedPPO(env, hidden=128, K_epochs=3, seed=7)
    losses = [ppo.iterate(K)["losses"] for _ in range(iters)]
    return ppo._flat.clone(), losses


flat, losses = run(n, rank * n)
gathered = [torch.empty_like(flat) for _ in range(world)]
dist.all_gather(gathered, flat)
same = all(torch.equal(gathered[0], g) for g in gathered)
if rank == 0:
    print("ranks in lock-step:", same, " losses per iteration:", [[round(x, 5) for x in l] for l in losses], flush=True)
dist.barrier()
dist.destroy_process_group()
if rank == 0:                                           # the same job on one GPU (no process group: BatchedPPO sees world = 1)
    env = BatchedQuad(N_total, 0.01, 1000, training=True, direct_control=1, T=5, precision="f32", async_reset=True, seed=3, device=dev)
    env.reset()
    ppo = BatchedPPO(env, hidden=128, K_epochs=3, seed=7)
    l1 = [ppo.iterate(K)["losses"] for _ in range(iters)]
    d = float((ppo._flat - flat).abs().max()); rel = float((ppo._flat - flat).norm() / ppo._flat.norm())
    print("single-GPU losses:", [[round(x, 5) for x in l] for l in l1], flush=True)
    print("max |param diff| sharded vs single: %.3e  (relative L2 %.3e)" % (d, rel), flush=True)
    ok = same and rel < 5e-3 and all(abs(a - b) < 5e-3 * max(1.0, abs(b)) for la, lb in zip(losses, l1) for a, b in zip(la, lb))
    print("PPO DDP CHECK", "OK" if ok else "FAILED", flush=True)
