"""On-device PPO iteration around the fused rollout — host-side mirror of the reference trainer
`environment/controller/ppo.py` (class PPO :80-209, model.py ActorCritic) for the batched simulator:

    rollout   BatchedQuad.policy_rollout   history -> actor (tcgen05) -> Normal sample -> quad.step, K steps per launch
    values    critic head of the same kernel (second pair of tcgen05 GEMMs on the same history tile; fused_critic=False: torch GEMMs)
    GAE       qs_gae / qs_adv_normalize    hand-written backward scan + masked normalisation on the [K][N] buffers
    update    K_epochs full-batch clipped-surrogate steps (ppo.py:164-206): qs_ppo_grad (forward AND backward of each
              75-128-128-{4,1} network on tcgen05, weight-gradient accumulators resident in tensor memory) for the actor and
              the critic, one all-reduce of the flat gradient across ranks (the path's second collective: ~0.2 MB per step),
              qs_adam_step on the flat FP32 master parameters

Nothing of the 1M x 128 rollout leaves the GPU, and on a CUDA device no part of an iteration runs in PyTorch: torch supplies
the device buffers and the NCCL plumbing.  update_impl="torch" (autograd + torch.optim.Adam over the same flat parameters) is
what runs on a CPU device — the gloo tests of the host logic — and is the FP32 comparison of the kernel in tests/test_ppo.py.
"""
from __future__ import annotations

import ctypes as C
import math

import torch
import torch.distributed as dist
import torch.nn as nn

from . import _lib as L


class ActorCritic(nn.Module):
    """model.py:19-47 — actor Linear-Tanh-Linear-Tanh-Linear-Tanh, critic Linear-Tanh-Linear-Tanh-Linear, fixed std."""

    def __init__(self, hidden: int = 128, state_dim: int = 75, action_dim: int = 4, action_std: float = 0.1):
        super().__init__()
        self.actor = nn.Sequential(nn.Linear(state_dim, hidden), nn.Tanh(), nn.Linear(hidden, hidden), nn.Tanh(),
                                   nn.Linear(hidden, action_dim), nn.Tanh())
        self.critic = nn.Sequential(nn.Linear(state_dim, hidden), nn.Tanh(), nn.Linear(hidden, hidden), nn.Tanh(),
                                    nn.Linear(hidden, 1))
        # model.py:42 keeps std as a float32 Parameter and Normal(mean, ones(4)*std) keeps a float32 scale: its square and
        # its log are rounded to float32 before they meet the (double) means — reproduced so that losses agree to 1e-12
        s32 = torch.tensor(action_std, dtype=torch.float32)
        self.std, self._var, self._log_std = float(s32), float(s32 * s32), float(torch.log(s32))
        self._entropy = float(0.5 + 0.5 * math.log(2 * math.pi) + torch.log(s32))     # Normal.entropy() stays float32

    def logprob(self, state, action):
        """Per-dimension log-prob of `action` under Normal(actor(state), std) (the actor half of evaluate)."""
        mean = self.actor(state)
        return -((action - mean) ** 2) / (2 * self._var) - self._log_std - math.log(math.sqrt(2 * math.pi))

    def evaluate(self, state, action):
        """model.py:74-88: per-dimension log-prob of `action` under Normal(actor(state), std), state value, entropy."""
        mean = self.actor(state)
        logp = -((action - mean) ** 2) / (2 * self._var) - self._log_std - math.log(math.sqrt(2 * math.pi))
        entropy = torch.full_like(mean, self._entropy)
        return logp, self.critic(state).squeeze(-1), entropy


def ppo_loss(policy: ActorCritic, states, actions, old_logprobs, advantages, returns, weight=None, eps_clip: float = 0.2):
    """The loss of PPO.update (ppo.py:183-201) on one batch; `weight` (0/1) drops warm-up transitions.  Returns the SUM over
    the batch of the per-sample loss (divide by the global sample count: the reference takes loss.mean(), :203; its scalar
    0.5*MSELoss term is the mean of 0.5*(V - R)^2, so it distributes over the samples)."""
    logp, values, entropy = policy.evaluate(states, actions)
    ratios = torch.exp(logp.sum(-1) - old_logprobs.sum(-1))                          # :187
    surr1 = ratios * advantages                                                      # :192
    surr2 = torch.clamp(ratios, 1 - eps_clip, 1 + eps_clip) * advantages             # :193
    sq = (values - returns) ** 2
    per = -torch.min(surr1, surr2) + 0.5 * sq - 0.006 * entropy.sum(-1)              # :194-200 (0.5*MSE enters every sample's loss)
    if weight is not None:
        per = per * weight
    return per.sum()


class BatchedPPO:
    """class PPO (ppo.py:80-209) driving a BatchedQuad.  Hyper-parameters default to the reference's (:296-318)."""

    def __init__(self, env, hidden: int = 128, action_std: float = 0.1, lr: float = 5e-4, betas=(0.9, 0.999), gamma: float = 0.99,
                 lmbda: float = 0.99, K_epochs: int = 10, eps_clip: float = 0.2, chunk_envs: int = 16384, seed: int = 0,
                 tf32: bool = True, fused_critic: bool = True, update_impl: str | None = None, device=None):
        self.env = env
        self.fused_critic = bool(fused_critic)   # state values from the critic head of the fused rollout kernel (BF16 tensor cores)
        self.dev = env.device if env is not None else torch.device(device or "cpu")     # env=None: update()-only use (tests)
        torch.manual_seed(seed)
        self.policy = ActorCritic(hidden, 75, 4, action_std).to(self.dev)
        # one flat FP32 master vector [actor w1 b1 w2 b2 w3 b3 | critic ...]; the module's parameters are views of it
        params = list(self.policy.actor.parameters()) + list(self.policy.critic.parameters())
        self._flat = torch.cat([p.detach().reshape(-1) for p in params]).contiguous()
        o = 0
        for p in params:
            p.data = self._flat[o:o + p.numel()].view_as(p); o += p.numel()
        self._n_actor = sum(p.numel() for p in self.policy.actor.parameters())
        self.optimizer = torch.optim.Adam(self.policy.parameters(), lr=lr, betas=betas)
        self.update_impl = update_impl or ("kernel" if self.dev.type == "cuda" else "torch")
        if self.update_impl not in ("kernel", "torch"):
            raise ValueError("update_impl must be 'kernel' or 'torch'")
        if self.update_impl == "kernel":
            if self.dev.type != "cuda" or hidden != 128:
                raise RuntimeError("update_impl='kernel' is the sm_100a path: CUDA device and hidden=128 (model.py:24) required")
            self.lib = env.lib if env is not None else L.load_library()
            self._grad = torch.zeros_like(self._flat)
            self._exp_avg, self._exp_avg_sq = torch.zeros_like(self._flat), torch.zeros_like(self._flat)
            self._adam_step = 0
            self.lr, self.betas = float(lr), (float(betas[0]), float(betas[1]))
        self.gamma, self.lmbda, self.K_epochs, self.eps_clip = gamma, lmbda, K_epochs, eps_clip
        self.chunk = int(chunk_envs)
        self.tf32 = bool(tf32)            # TF32 tensor-core GEMMs for the update's autograd (FP32 accumulate); False = IEEE FP32
        self.moments = torch.zeros(3, dtype=torch.float64, device=self.dev)
        self._sync_actor()

    def _sync_actor(self):
        """policy_old.load_state_dict(policy.state_dict()) (:206): the rollout kernel reads the actor weights."""
        if self.env is None:
            return
        self.env.load_actor({"actor_%d_%s" % (i, k): getattr(self.policy.actor[i], k).detach() for i in (0, 2, 4) for k in ("weight", "bias")},
                            action_std=self.policy.std,
                            critic={"critic_%d_%s" % (i, k): getattr(self.policy.critic[i], k).detach() for i in (0, 2, 4) for k in ("weight", "bias")}
                            if self.fused_critic else None)

    # ---------------------------------------------------------------------------------------------------------
    @staticmethod
    def history_entries(rec):
        """(K,15,N) dl_in_gen entries [action(4), v(3), q(4), dq(4)] (dl_auxiliary.py:27-30) of a recorded rollout, rounded
        to BF16 like the A operand the fused actor reads."""
        obs, act = rec["obs"], rec["actions"]
        e = torch.cat([act, obs[:, (1, 3, 5)], obs[:, 6:14]], dim=1)
        return e.bfloat16().float()

    def network_inputs(self, hist0, entries, n0, n1):
        """Inputs the actor saw at every step for envs [n0,n1): (K, n, 75), oldest entry first (dl_auxiliary.py:25-32).
        hist0 (75,N) = history at rollout start; entries (K,15,N)."""
        K = entries.shape[0]
        h = hist0[:, n0:n1].reshape(5, 15, n1 - n0)
        seq = torch.cat([h, entries[:, :, n0:n1]], dim=0)                 # (5+K, 15, n): input of step t = seq[t:t+5]
        x = seq.unfold(0, 5, 1)[:K + 1]                                   # (K+1, 15, n, 5)
        return x.permute(0, 2, 3, 1).reshape(K + 1, n1 - n0, 75)          # [..., 15*slot + e]

    @torch.no_grad()
    def collect(self, horizon: int):
        """One rollout + values + GAE.  Returns the batch dict (all tensors on the device, time-major)."""
        env = self.env
        hist0 = env.history.t().contiguous().clone()                      # (75, N)
        sensed = bool(env._cfg.flags & L.QS_FLAG_SENSOR_NOISE)             # the policy's observation is the sensed one there
        rec = env.policy_rollout(horizon, record_obs=not sensed, record_actions=True, record_logprob=True, record_reward=True,
                                 record_done=True, record_values=self.fused_critic, record_sensed=sensed)
        if sensed:
            rec["obs"] = rec["sensed_obs"]
        kernel_update = self.update_impl == "kernel"
        # update_impl="kernel": qs_ppo_grad builds the history entries from the recorded observations and actions on the fly
        entries = None if kernel_update else self.history_entries(rec)
        K, N = horizon, env.N
        value = rec["value"] if self.fused_critic else torch.empty(K + 1, N, dtype=torch.float32, device=self.dev)
        # old_logprobs of the update's ratio (ppo.py:187) must come from the SAME evaluation path as the new ones: the reference's
        # policy_old is an exact copy of policy, so the ratio is exactly 1 at the first epoch.  The fused kernel's own log-probs
        # (BF16 operands, tanh.approx) differ from the update's FP32/TF32 re-evaluation by O(0.1-1) in the tails, which would
        # clip or over-weight those samples: they are kept as a diagnostic (batch["logprob_kernel"]) and the ratio's
        # denominator is re-evaluated here, on the recorded actions, by the network the update differentiates.
        # update_impl="kernel": qs_ppo_grad's first epoch records them from its own forward pass (QS_PPO_RECORD_LOGP).
        logprob = torch.empty(K, 4, N, dtype=torch.float32, device=self.dev)
        if kernel_update and not self.fused_critic:
            raise RuntimeError("update_impl='kernel' takes the state values from the fused critic head (fused_critic=True)")
        prev_tf32 = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = self.tf32
        for n0 in range(0, N if not kernel_update else 0, self.chunk):
            n1 = min(N, n0 + self.chunk)
            x = self.network_inputs(hist0, entries, n0, n1)
            if not self.fused_critic:
                value[:, n0:n1] = self.policy.critic(x).squeeze(-1)
            lp = self.policy.logprob(x[:K], rec["actions"][:, :, n0:n1].permute(0, 2, 1))
            logprob[:, :, n0:n1] = lp.permute(0, 2, 1)
        torch.backends.cuda.matmul.allow_tf32 = prev_tf32
        ret = torch.empty(K, N, dtype=torch.float32, device=self.dev)
        adv = torch.empty(K, N, dtype=torch.float32, device=self.dev)
        weight = torch.empty(K, N, dtype=torch.float32, device=self.dev)
        self.moments.zero_()
        st = C.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream)
        L.check(env.lib.qs_gae(N, K, self.gamma, self.lmbda, rec["reward"].data_ptr(), value.data_ptr(), rec["done"].data_ptr(),
                               ret.data_ptr(), adv.data_ptr(), self.moments.data_ptr(), st))
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.moments)                                 # normalise over the GLOBAL batch
        L.check(env.lib.qs_adv_normalize(K * N, rec["done"].data_ptr(), self.moments.data_ptr(), adv.data_ptr(), weight.data_ptr(), st))
        return dict(hist0=hist0, entries=entries, obs=rec["obs"], actions=rec["actions"], logprob=logprob, logprob_kernel=rec["logprob"], reward=rec["reward"],
                    done=rec["done"], value=value, returns=ret, adv=adv, weight=weight, count=float(self.moments[0].item()),
                    logprob_pending=kernel_update)

    # ---------------------------------------------------------------------------------------------------------
    def _net_ptrs(self, flat, critic: bool):
        """qs_ppo_net over one network's slice of a flat vector (PyTorch Linear layout, [out][in])."""
        out = 1 if critic else 4
        o = self._n_actor if critic else 0
        ptrs = []
        for n in (128 * 75, 128, 128 * 128, 128, out * 128, out):
            ptrs.append(flat.data_ptr() + 4 * o); o += n
        return L.qs_ppo_net(*ptrs)

    def _batch_tensors(self, batch):
        t = {k: batch[k].contiguous().float() for k in ("hist0", "actions", "adv", "returns", "weight")}
        for k in ("entries", "obs"):                               # entries, or the recorded observations to build them from
            t[k] = batch[k].contiguous().float() if batch.get(k) is not None else None
        if t["entries"] is None and t["obs"] is None:
            raise ValueError("the batch needs 'entries' (K,15,N) or 'obs' (K,14,N)")
        logp = batch["logprob"]
        if not (logp.is_contiguous() and logp.dtype == torch.float32):
            raise ValueError("batch['logprob'] must be a contiguous float32 (K,4,N) tensor (QS_PPO_RECORD_LOGP writes it in place)")
        t["logprob"] = logp
        return t

    def _grad_pass(self, t, K, N, count, record: bool, loss_row):
        """self._grad <- d(loss)/d(parameters) of the whole local batch: zero + qs_ppo_grad(actor) + qs_ppo_grad(critic)."""
        st = C.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream)
        self._grad.zero_()
        ptr = lambda x: x.data_ptr() if x is not None else None
        bt = L.qs_ppo_batch(N, K, L.QS_PPO_RECORD_LOGP if record else 0, t["hist0"].data_ptr(), ptr(t["entries"]),
                            None if t["entries"] is not None else ptr(t["obs"]), t["actions"].data_ptr(), t["logprob"].data_ptr(), t["adv"].data_ptr(), t["returns"].data_ptr(), t["weight"].data_ptr())
        for which in (L.QS_PPO_ACTOR, L.QS_PPO_CRITIC):
            net, grad = self._net_ptrs(self._flat, which == L.QS_PPO_CRITIC), self._net_ptrs(self._grad, which == L.QS_PPO_CRITIC)
            L.check(self.lib.qs_ppo_grad(C.byref(bt), C.byref(net), C.byref(grad), which, self.policy.std, self.eps_clip, float(count),
                                         loss_row[which].data_ptr(), st))

    def gradients(self, batch, record_logprob: bool = False):
        """One full-batch gradient of the local shard (no all-reduce, no optimizer step): (flat gradient, [actor, critic] loss)."""
        K, N = batch["adv"].shape
        loss = torch.zeros(2, dtype=torch.float64, device=self.dev)
        with torch.cuda.device(self.dev):
            self._grad_pass(self._batch_tensors(batch), K, N, batch["count"], record_logprob, loss)
        return self._grad.clone(), loss

    def _update_kernel(self, batch):
        """The K_epochs steps on qs_ppo_grad / qs_adam_step.  No host synchronisation inside the loop."""
        K, N = batch["adv"].shape
        world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
        t = self._batch_tensors(batch)
        pending = bool(batch.get("logprob_pending", False))
        st = C.c_void_p(torch.cuda.current_stream(self.dev).cuda_stream)
        loss = torch.zeros(self.K_epochs, 2, dtype=torch.float64, device=self.dev)
        with torch.cuda.device(self.dev):
            for e in range(self.K_epochs):
                self._grad_pass(t, K, N, batch["count"], pending and e == 0, loss[e])
                if world > 1:
                    dist.all_reduce(self._grad)                           # ~0.2 MB: sum of the per-rank partial gradients
                self._adam_step += 1
                L.check(self.lib.qs_adam_step(self._flat.numel(), self._flat.data_ptr(), self._grad.data_ptr(), self._exp_avg.data_ptr(),
                                              self._exp_avg_sq.data_ptr(), self._adam_step, self.lr, self.betas[0], self.betas[1], 1e-8, st))
        batch["logprob_pending"] = False
        if world > 1:
            dist.all_reduce(loss)
        # the constant entropy bonus of ppo.py:200 (fixed std: no gradient) enters the reported loss only
        return [float(x) - 0.006 * 4 * self.policy._entropy for x in loss.sum(dim=1).tolist()]

    def update(self, batch):
        """PPO.update (:143-206): K_epochs full-batch steps (the reference's randperm does not change a full-batch mean)."""
        if self.update_impl == "kernel":
            losses = self._update_kernel(batch)
            self._sync_actor()
            return losses
        K, N = batch["adv"].shape
        world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
        count = batch["count"]                                            # global number of valid transitions
        losses = []
        prev_tf32 = torch.backends.cuda.matmul.allow_tf32
        if self.dev.type == "cuda":
            torch.backends.cuda.matmul.allow_tf32 = self.tf32
        for _ in range(self.K_epochs):
            self.optimizer.zero_grad(set_to_none=False)
            total = torch.zeros((), dtype=torch.float64, device=self.dev)
            for n0 in range(0, N, self.chunk):
                n1 = min(N, n0 + self.chunk)
                x = self.network_inputs(batch["hist0"], batch["entries"], n0, n1)[:K]
                loss = ppo_loss(self.policy, x, batch["actions"][:, :, n0:n1].permute(0, 2, 1), batch["logprob"][:, :, n0:n1].permute(0, 2, 1),
                                batch["adv"][:, n0:n1], batch["returns"][:, n0:n1], batch["weight"][:, n0:n1], self.eps_clip)
                (loss / count).backward()                                 # loss.mean().backward() over the global batch (:203)
                total += loss.detach().double()
            if world > 1:
                flat = torch.cat([p.grad.reshape(-1) for p in self.policy.parameters() if p.grad is not None])
                dist.all_reduce(flat)                                     # ~0.2 MB: sum of the per-rank partial gradients
                o = 0
                for p in self.policy.parameters():
                    if p.grad is not None:
                        n = p.grad.numel(); p.grad.copy_(flat[o:o + n].view_as(p.grad)); o += n
                dist.all_reduce(total)
            self.optimizer.step()
            losses.append(float(total.item()) / count)
        torch.backends.cuda.matmul.allow_tf32 = prev_tf32
        self._sync_actor()
        return losses

    def iterate(self, horizon: int = 128):
        batch = self.collect(horizon)
        losses = self.update(batch)
        rsum = (batch["reward"] * batch["weight"]).sum().double()
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(rsum)                                         # batch["count"] is the GLOBAL number of valid transitions
        return dict(mean_reward=float(rsum.item()) / max(1.0, batch["count"]), losses=losses, stats=self.env.stats(reset=True))
