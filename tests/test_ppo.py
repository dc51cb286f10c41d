"""SURVEY.md §8(f)1 — PPO iteration around the fused rollout: GAE scan, advantage normalisation, clipped-surrogate update.
CPU tests pin the restatements to the reference's OWN code (tests/golden/ppo_vectors.npz is produced by executing
PPO.get_advantages and ActorCritic.evaluate lifted unmodified from the reference, oracle/gen_golden.py:gen_ppo_vectors)."""
import ctypes as C
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from conftest import load_golden
from oracle import quad_oracle as qo
from autonomous_quadrotor_environment_b200 import ppo as P

HAS_CUDA = torch.cuda.is_available()


def test_gae_oracle_vs_reference_function():
    g = load_golden("ppo_vectors.npz")
    for i in (0, 1):
        ret, adv = qo.gae_advantages(g["values%d" % i], np.logical_not(g["terminals%d" % i]), g["rewards%d" % i])
        assert np.abs(ret - g["returns%d" % i]).max() < 1e-12
        assert np.abs(adv - g["adv%d" % i]).max() < 1e-12
        # the batched time-major form (what the CUDA kernel computes) on the same sequence as one env
        done = g["terminals%d" % i].astype(np.uint8)[:, None]
        r2, a2, valid, n2 = qo.gae_batched(g["rewards%d" % i][:, None], g["values%d" % i][:, None], done)
        assert valid.all() and np.abs(r2[:, 0] - ret).max() < 1e-12 and np.abs(n2[:, 0] - adv).max() < 1e-12


def _golden_policy(g):
    pol = P.ActorCritic(32, 75, 4, 0.1).double()
    sd = {k[2:].replace("actor_", "actor.").replace("critic_", "critic.").replace("_weight", ".weight").replace("_bias", ".bias"): torch.tensor(g[k])
          for k in g if k.startswith("w_actor") or k.startswith("w_critic")}
    pol.load_state_dict(sd)
    return pol


def test_ppo_loss_and_gradients_vs_reference_model():
    """Loss value and every parameter gradient of one update step equal the reference's (model.py evaluate + ppo.py:183-203)."""
    g = load_golden("ppo_vectors.npz")
    pol = _golden_policy(g)
    T = lambda k: torch.tensor(g[k])
    B = g["loss_adv"].shape[0]
    loss = P.ppo_loss(pol, T("loss_states"), T("loss_actions"), T("loss_old_logprobs"), T("loss_adv"), T("loss_returns")) / B
    # the (constant, gradient-free) entropy term of the reference is evaluated in float32 (Normal.entropy() of a float32 scale):
    # the loss VALUE agrees to that rounding, the gradients below to 1e-12
    assert abs(float(loss.detach()) - float(g["loss_value"])) < 1e-8
    loss.backward()
    for name, p in pol.named_parameters():
        ref = g["grad_" + name.replace(".", "_")]
        assert np.abs(p.grad.numpy() - ref).max() < 1e-12, name


def _fake_batch(K, N, seed, dev="cpu"):
    gen = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=gen)
    w = (torch.rand(K, N, generator=gen) > 0.1).float()
    return dict(hist0=r(75, N), entries=r(K, 15, N), actions=r(K, 4, N) * 0.3, logprob=r(K, 4, N) * 0.1 + 1.0, reward=r(K, N),
                returns=r(K, N), adv=r(K, N), weight=w, count=float(w.sum()))


def test_update_is_chunk_invariant_and_moves_parameters():
    b = _fake_batch(6, 40, 1)
    outs = []
    for chunk in (40, 7):
        ppo = P.BatchedPPO(None, hidden=16, chunk_envs=chunk, K_epochs=3, seed=3)
        losses = ppo.update(b)
        outs.append((losses, torch.cat([p.detach().reshape(-1) for p in ppo.policy.parameters()])))
    assert np.allclose(outs[0][0], outs[1][0], rtol=1e-5)
    assert torch.allclose(outs[0][1], outs[1][1], rtol=1e-4, atol=1e-6)
    fresh = P.BatchedPPO(None, hidden=16, seed=3)
    assert not torch.allclose(outs[0][1], torch.cat([p.detach().reshape(-1) for p in fresh.policy.parameters()]))


def _ppo_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    full = _fake_batch(5, 32, 9)
    n0, n1 = rank * 16, (rank + 1) * 16
    part = {k: (v[..., n0:n1].contiguous() if torch.is_tensor(v) else v) for k, v in full.items()}
    part["count"] = full["count"]                       # the GLOBAL number of valid transitions
    ppo = P.BatchedPPO(None, hidden=16, chunk_envs=16, K_epochs=2, seed=5)
    losses = ppo.update(part)
    q.put((rank, losses, torch.cat([p.detach().reshape(-1) for p in ppo.policy.parameters()]).numpy()))
    dist.destroy_process_group()


def test_update_world2_gloo_equals_single_process():
    """Env-sharded update: two ranks with half of the envs each + gradient all-reduce == one process with all envs."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 300
    ps = [ctx.Process(target=_ppo_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in ps]
    res = sorted([q.get(timeout=120) for _ in ps], key=lambda t: t[0])
    [p.join(60) for p in ps]
    single = P.BatchedPPO(None, hidden=16, chunk_envs=32, K_epochs=2, seed=5)
    l1 = single.update(_fake_batch(5, 32, 9))
    w1 = torch.cat([p.detach().reshape(-1) for p in single.policy.parameters()]).numpy()
    assert np.allclose(res[0][2], res[1][2], rtol=0, atol=0)            # ranks stay in lock-step
    assert np.allclose(res[0][2], w1, rtol=1e-4, atol=1e-6) and np.allclose(res[0][1], l1, rtol=1e-5)


# ------------------------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
def test_gae_kernel_vs_oracle():
    from autonomous_quadrotor_environment_b200 import _lib as L
    lib = L.load_library()
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(3)
    for K, N in ((128, 5000), (37, 1), (8, 333)):
        reward = rng.normal(0, 1, (K, N)).astype(np.float32); value = rng.normal(0, 2, (K + 1, N)).astype(np.float32)
        done = ((rng.random((K, N)) < 0.04).astype(np.uint8)) | ((rng.random((K, N)) < 0.1).astype(np.uint8) << 1)
        ret_ref, adv_ref, valid, norm_ref = qo.gae_batched(reward.astype(np.float64), value.astype(np.float64), done)
        tr, tv, td = (torch.as_tensor(x, device=dev) for x in (reward, value, done))
        ret = torch.empty(K, N, device=dev); adv = torch.empty(K, N, device=dev); w = torch.empty(K, N, device=dev)
        mom = torch.zeros(3, dtype=torch.float64, device=dev)
        L.check(lib.qs_gae(N, K, 0.99, 0.99, tr.data_ptr(), tv.data_ptr(), td.data_ptr(), ret.data_ptr(), adv.data_ptr(), mom.data_ptr(), None))
        assert np.allclose(ret.cpu().numpy(), ret_ref, rtol=2e-5, atol=2e-4) and np.allclose(adv.cpu().numpy(), adv_ref, rtol=2e-5, atol=2e-4)
        assert float(mom[0]) == valid.sum() and abs(float(mom[1]) - adv_ref[valid].sum()) < 1e-3 * max(1.0, np.abs(adv_ref[valid]).sum())
        L.check(lib.qs_adv_normalize(K * N, td.data_ptr(), mom.data_ptr(), adv.data_ptr(), w.data_ptr(), None))
        assert np.allclose(adv.cpu().numpy(), norm_ref, rtol=1e-4, atol=1e-4)
        assert np.array_equal(w.cpu().numpy() > 0, valid)


@pytest.mark.gpu
def test_ppo_iteration_on_device():
    """collect(): the reconstructed network inputs are the ones the fused actor saw (recorded log-probs follow from the torch
    actor's means on them); iterate(): losses finite, parameters move, the rollout kernel picks up the new actor."""
    from autonomous_quadrotor_environment_b200 import BatchedQuad
    dev = torch.device("cuda", 0)
    N, K = 4096, 32
    env = BatchedQuad(N, 0.01, 1000, training=True, direct_control=1, T=5, precision="f32", async_reset=True, seed=2, device=dev)
    env.reset()
    ppo = P.BatchedPPO(env, hidden=128, K_epochs=2, chunk_envs=1024, seed=1, tf32=False, update_impl="torch")
    b = ppo.collect(K)
    assert b["value"].shape == (K + 1, N) and b["adv"].shape == (K, N) and 0 < b["count"] <= K * N
    x = ppo.network_inputs(b["hist0"], b["entries"], 0, 512)[:K]
    with torch.no_grad():
        mean = ppo.policy.actor(x)                                         # (K, 512, 4)
    a = b["actions"][:, :, :512].permute(0, 2, 1); lp = b["logprob_kernel"][:, :, :512].permute(0, 2, 1)
    sigma = ppo.policy.std
    lp_torch = -((a - mean) ** 2) / (2 * sigma * sigma) - np.log(sigma) - 0.9189385
    ok = (b["weight"][:, :512] > 0).unsqueeze(-1).expand_as(lp)
    # BF16 operands in the kernel vs FP32 here: |mean error| ~ 1e-2 -> log-prob error ~ |z| * 0.1 + small
    err = (lp_torch - lp)[ok].abs()
    assert float(err.median()) < 0.05 and float(err.quantile(0.99)) < 1.0, (float(err.median()), float(err.quantile(0.99)))
    # the ratio's denominator is the update network's OWN evaluation of the recorded actions (policy_old == policy at epoch 0,
    # ppo.py:187,:206): the first epoch's ratio is exactly 1 for every valid sample
    with torch.no_grad():
        lp_new, _, _ = ppo.policy.evaluate(x, a)
    assert torch.allclose(lp_new, b["logprob"][:, :, :512].permute(0, 2, 1), rtol=0, atol=2e-3)    # FP32 GEMMs of two batch shapes
    ratio = torch.exp(lp_new.sum(-1) - b["logprob"][:, :, :512].permute(0, 2, 1).sum(-1))
    assert float((ratio - 1).abs().max()) < 1e-2
    before = torch.cat([p.detach().reshape(-1) for p in ppo.policy.parameters()]).clone()
    out = ppo.iterate(K)
    after = torch.cat([p.detach().reshape(-1) for p in ppo.policy.parameters()])
    assert all(np.isfinite(out["losses"])) and np.isfinite(out["mean_reward"]) and not torch.equal(before, after)
    assert torch.equal(env._actor[1]["w1"], ppo.policy.actor[0].weight.detach())


def _bf16_st(x):
    """Round to BF16 in the forward pass, identity in the backward pass (the kernel's operand rounding, seen by autograd)."""
    return x + (x.bfloat16().float() - x).detach()


def _torch_grads(ppo, b, bf16_forward):
    """Gradient of the PPO loss (ppo.py:183-203) w.r.t. the flat parameter vector by torch autograd in FP32 (TF32 off);
    bf16_forward: W1, W2, the inputs and the first hidden activation are rounded to BF16 like the kernel's MMA operands."""
    K, N = b["adv"].shape
    x = ppo.network_inputs(b["hist0"], b["entries"], 0, N)[:K]
    rnd = _bf16_st if bf16_forward else (lambda v: v)

    def mlp(seq, out_tanh):
        h1 = torch.tanh(torch.nn.functional.linear(rnd(x), rnd(seq[0].weight), seq[0].bias))
        h2 = torch.tanh(torch.nn.functional.linear(rnd(h1), rnd(seq[2].weight), seq[2].bias))
        o = torch.nn.functional.linear(h2, seq[4].weight, seq[4].bias)
        return torch.tanh(o) if out_tanh else o

    pol = ppo.policy
    mean, value = mlp(pol.actor, True), mlp(pol.critic, False).squeeze(-1)
    a = b["actions"].permute(0, 2, 1)
    logp = -((a - mean) ** 2) / (2 * pol._var) - pol._log_std - 0.5 * np.log(2 * np.pi)
    ratio = torch.exp(logp.sum(-1) - b["logprob"].permute(0, 2, 1).sum(-1))
    surr = torch.min(ratio * b["adv"], torch.clamp(ratio, 1 - ppo.eps_clip, 1 + ppo.eps_clip) * b["adv"])
    la = (-(surr) * b["weight"]).sum() / b["count"]
    lc = (0.5 * (value - b["returns"]) ** 2 * b["weight"]).sum() / b["count"]
    params = list(pol.actor.parameters()) + list(pol.critic.parameters())
    g = torch.autograd.grad(la + lc, params)
    return torch.cat([t.reshape(-1) for t in g]), float(la.detach()), float(lc.detach()), logp.detach()


def _bf(v):
    return v.bfloat16().float()


def _kernel_arithmetic_grads(ppo, b, rnd=_bf):
    """The arithmetic of qs_ppo_grad restated in torch, operand roundings included (csrc/ppo_update.cuh): BF16 MMA operands (X, W1, W2,
    H1, H2, dZ3, dZ2, dZ1 and the parked 1 - H1^2), FP32 accumulation, FP32 output layer and loss, the gradient of torch.min/clamp as
    a branch select.  rnd = identity turns it into a plain manual backward pass (checked against autograd on the CPU).
    Returns (flat gradient, actor loss, critic loss, per-dimension log-probs (K, N, 4))."""
    K, N = b["adv"].shape
    x = ppo.network_inputs(b["hist0"], b["entries"], 0, N)[:K].reshape(K * N, 75)
    wgt = (b["weight"] / b["count"]).reshape(-1)
    pol, eps = ppo.policy, ppo.eps_clip
    grads, losses, logp = [], [], None
    for net, seq in ((0, pol.actor), (1, pol.critic)):
        W1, b1, W2, b2, W3, b3 = [p.detach() for p in seq.parameters()]
        h1 = torch.tanh(rnd(x) @ rnd(W1).t() + b1)
        h2 = torch.tanh(rnd(h1) @ rnd(W2).t() + b2)
        out = h2 @ W3.t() + b3
        if net == 0:
            mean = torch.tanh(out)
            a = b["actions"].permute(0, 2, 1).reshape(K * N, 4)
            d = a - mean
            logp = -0.5 * d * d / pol._var - pol._log_std - 0.5 * np.log(2 * np.pi)
            ratio = torch.exp(logp.sum(-1) - b["logprob"].permute(0, 2, 1).reshape(K * N, 4).sum(-1))
            adv = b["adv"].reshape(-1)
            surr1, surr2 = ratio * adv, torch.clamp(ratio, 1 - eps, 1 + eps) * adv
            through = (surr1 <= surr2) | ((ratio >= 1 - eps) & (ratio <= 1 + eps))
            losses.append(float((wgt * -torch.minimum(surr1, surr2)).sum()))
            dl = torch.where(through, -adv * ratio, torch.zeros_like(adv))
            dz3 = (wgt * dl)[:, None] * d / pol._var * (1 - mean * mean)
        else:
            dv = out.squeeze(-1) - b["returns"].reshape(-1)
            losses.append(float((wgt * 0.5 * dv * dv).sum()))
            dz3 = (wgt * dv)[:, None]
        dW3, db3 = rnd(dz3).t() @ rnd(h2), dz3.sum(0)
        dz2 = rnd((dz3 @ W3) * (1 - h2 * h2))
        dW2, db2 = dz2.t() @ rnd(h1), dz2.sum(0)
        dz1 = rnd((dz2 @ rnd(W2)) * rnd(1 - h1 * h1))
        dW1, db1 = dz1.t() @ rnd(x), dz1.sum(0)
        grads += [dW1, db1, dW2, db2, dW3, db3]
    return torch.cat([t.reshape(-1) for t in grads]), losses[0], losses[1], logp.reshape(K, N, 4)


def _kernel_batch(K, N, seed, dev, ppo, spread):
    """Synthetic rollout buffers: BF16-representable history, actions = mean + sigma z.  spread > 0: old log-probs such that the ratio of
    the BF16-operand forward pass is drawn from [0.9, 1.1], [0.3, 0.7] and [1.35, 2] (both clip branches, and no sample within 5 % of
    a clip boundary: MUFU.TANH vs tanh moves a ratio by ~1 %, and a sample that changes branch changes the gradient by its whole term)."""
    gen = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=gen).to(dev)
    b = dict(hist0=(r(75, N) * 0.7).bfloat16().float(), entries=(r(K, 15, N) * 0.7).bfloat16().float(), adv=r(K, N), returns=r(K, N) * 2)
    b["weight"] = (torch.rand(K, N, generator=gen) > 0.15).float().to(dev)
    b["count"] = float(b["weight"].sum())
    with torch.no_grad():
        x = ppo.network_inputs(b["hist0"], b["entries"], 0, N)[:K]
        a = ppo.policy.actor(x) + ppo.policy.std * r(K, N, 4)
    b["actions"] = a.permute(0, 2, 1).contiguous()
    b["logprob"] = torch.zeros(K, 4, N, device=dev)
    lp = _kernel_arithmetic_grads(ppo, b)[3]
    if spread > 0:
        u = torch.rand(K, N, generator=gen).to(dev); sel = torch.randint(0, 3, (K, N), generator=gen).to(dev)
        ratio = torch.where(sel == 0, 0.9 + 0.2 * u, torch.where(sel == 1, 0.3 + 0.4 * u, 1.35 + 0.65 * u))
        lp = lp - (torch.log(ratio) / 4).unsqueeze(-1)
    b["logprob"] = lp.permute(0, 2, 1).contiguous()
    return b


def test_manual_backward_equals_autograd_cpu():
    """The restatement the kernel is compared with, rounding switched off, IS the autograd gradient of the PPO loss (ppo.py:183-203)."""
    ppo = P.BatchedPPO(None, hidden=128, seed=11)
    b = _kernel_batch(4, 96, 5, torch.device("cpu"), ppo, spread=1.0)
    ident = lambda v: v
    g_m, la_m, lc_m, _ = _kernel_arithmetic_grads(ppo, b, ident)
    g_a, la_a, lc_a, _ = _torch_grads(ppo, b, False)
    assert abs(la_m - la_a) < 1e-5 and abs(lc_m - lc_a) < 1e-5
    assert float((g_m - g_a).norm() / g_a.norm()) < 1e-4
    frac_low = float((torch.exp(_kernel_arithmetic_grads(ppo, b)[3].sum(-1) - b["logprob"].permute(0, 2, 1).sum(-1)) < 0.8).float().mean())
    assert 0.2 < frac_low < 0.5                                 # the clipped branches are populated


@pytest.mark.gpu
@pytest.mark.parametrize("K,N,sigma", [(6, 300, 0.1), (3, 128, 0.5), (17, 1000, 0.1), (70, 200, 0.1)])
def test_ppo_grad_kernel_vs_autograd(K, N, sigma):
    """qs_ppo_grad (tcgen05 forward + backward, BF16 operands) on synthetic rollout buffers, including a horizon that the launch splits
    into step chunks (70 steps): every parameter tensor's gradient within 5e-3 (relative L2) of the same arithmetic restated in torch
    (BF16 operand roundings included; what is left is MUFU.TANH vs tanh and the summation order), losses within 1e-3; against plain
    FP32 autograd only what BF16 operands allow (sigma = 0.1 turns a 3e-3 error of the mean into 3 % of (a - mean), and bias gradients
    are sums with heavy cancellation)."""
    dev = torch.device("cuda", 0)
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        ppo = P.BatchedPPO(None, hidden=128, action_std=sigma, seed=11, device=dev)
        assert ppo.update_impl == "kernel"
        b = _kernel_batch(K, N, 5, dev, ppo, spread=1.0)
        g_k, loss_k = ppo.gradients(b)
        g_e, la_e, lc_e, _ = _kernel_arithmetic_grads(ppo, b)
        g_f, la_f, lc_f, _ = _torch_grads(ppo, b, False)
        assert abs(float(loss_k[0]) - la_e) < 1e-3 * max(1.0, abs(la_e)) and abs(float(loss_k[1]) - lc_e) < 1e-3 * max(1.0, abs(lc_e)), (loss_k, la_e, lc_e)
        o = 0
        names = ["actor." + n for n, _ in ppo.policy.actor.named_parameters()] + ["critic." + n for n, _ in ppo.policy.critic.named_parameters()]
        params = list(ppo.policy.actor.parameters()) + list(ppo.policy.critic.parameters())
        for name, p in zip(names, params):
            n = p.numel()
            k, e, f = g_k[o:o + n], g_e[o:o + n], g_f[o:o + n]
            rel_e = float((k - e).norm() / e.norm()); rel_f = float((k - f).norm() / f.norm())
            assert rel_e < 5e-3, (name, rel_e, rel_f)
            assert rel_f < (0.3 if name.startswith("actor") else 5e-2), (name, rel_e, rel_f)
            o += n
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev


@pytest.mark.gpu
def test_ppo_grad_records_its_own_logprobs():
    """QS_PPO_RECORD_LOGP: memory.logprobs come from the update network's own forward pass (policy_old == policy, ppo.py:206), the
    first epoch's ratio is exactly 1 -> the surrogate's value is -mean(adv) and a second pass without the flag reproduces the gradient."""
    dev = torch.device("cuda", 0)
    ppo = P.BatchedPPO(None, hidden=128, seed=4, device=dev)
    b = _kernel_batch(5, 700, 8, dev, ppo, spread=0.0)
    lp_f32 = b["logprob"].clone()
    b["logprob"] = torch.full_like(lp_f32, float("nan"))
    g1, loss1 = ppo.gradients(b, record_logprob=True)
    assert torch.isfinite(b["logprob"]).all()
    expect = -float((b["adv"] * b["weight"]).sum() / b["count"])
    assert abs(float(loss1[0]) - expect) < 1e-5 * max(1.0, abs(expect))
    # BF16 operands vs the FP32 network: |mean error| ~ 3e-3 -> log-prob error ~ |z| * 0.03 (sigma = 0.1)
    err = (b["logprob"] - lp_f32).abs()
    assert float(err.median()) < 0.05 and float(err.quantile(0.99)) < 1.0
    g2, loss2 = ppo.gradients(b, record_logprob=False)
    assert float((g1 - g2).norm() / g1.norm()) < 1e-5 and abs(float(loss1[0] - loss2[0])) < 1e-6


@pytest.mark.gpu
def test_adam_kernel_vs_torch_optimizer():
    from autonomous_quadrotor_environment_b200 import _lib as L
    lib = L.load_library()
    dev = torch.device("cuda", 0)
    gen = torch.Generator().manual_seed(2)
    n = 100_003
    p0 = torch.randn(n, generator=gen).to(dev)
    ref = torch.nn.Parameter(p0.clone()); opt = torch.optim.Adam([ref], lr=5e-4, betas=(0.9, 0.999))
    p = p0.clone(); m = torch.zeros_like(p); v = torch.zeros_like(p)
    for step in range(1, 6):
        g = torch.randn(n, generator=gen).to(dev) * (10.0 ** (step - 3))
        ref.grad = g.clone(); opt.step()
        L.check(lib.qs_adam_step(n, p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), step, 5e-4, 0.9, 0.999, 1e-8,
                                 C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
        assert torch.allclose(p, ref.detach(), rtol=0, atol=5e-7), float((p - ref.detach()).abs().max())


@pytest.mark.gpu
def test_ppo_kernel_update_tracks_the_torch_update():
    """Two epochs of the hand-written update (qs_ppo_grad + qs_adam_step) vs the autograd + torch.optim.Adam update from the same
    parameters on the same batch: the parameter DISPLACEMENTS agree (Adam's normalised steps amplify gradient noise where the gradient
    is near zero, so the comparison is the cosine of the two displacement vectors and the losses)."""
    dev = torch.device("cuda", 0)
    a = P.BatchedPPO(None, hidden=128, seed=6, K_epochs=2, device=dev)
    t = P.BatchedPPO(None, hidden=128, seed=6, K_epochs=2, device=dev, update_impl="torch", tf32=False, chunk_envs=512)
    assert torch.equal(a._flat, t._flat)
    start = a._flat.clone()
    b = _kernel_batch(8, 2048, 12, dev, a, spread=1.0)
    la = a.update(dict(b)); lt = t.update(dict(b))
    da, dt_ = a._flat - start, t._flat - start
    cos = float((da * dt_).sum() / (da.norm() * dt_.norm()))
    assert cos > 0.97, cos
    assert abs(la[0] - lt[0]) < 5e-3 * max(1.0, abs(lt[0])) and abs(la[1] - lt[1]) < 2e-2 * max(1.0, abs(lt[1])), (la, lt)


@pytest.mark.gpu
def test_ppo_iteration_kernel_path():
    """The default on a CUDA device: collect (fused actor + critic rollout, GAE) -> update on qs_ppo_grad/qs_adam_step; the rollout
    kernel picks up the new weights; the loss of the second epoch is computed with the recorded log-probs."""
    from autonomous_quadrotor_environment_b200 import BatchedQuad
    dev = torch.device("cuda", 0)
    N, K = 4096, 32
    env = BatchedQuad(N, 0.01, 1000, training=True, direct_control=1, T=5, precision="f32", async_reset=True, seed=2, device=dev)
    env.reset()
    ppo = P.BatchedPPO(env, hidden=128, K_epochs=3, seed=1)
    assert ppo.update_impl == "kernel"
    # the history entries built on the fly from the recorded observations / actions == the materialised (K,15,N) buffer
    b = ppo.collect(K)
    assert b["entries"] is None and b["logprob_pending"]
    g1, l1 = ppo.gradients(b, record_logprob=True)
    b2 = dict(b); b2["entries"] = P.BatchedPPO.history_entries({"obs": b["obs"], "actions": b["actions"]}); b2["obs"] = None
    g2, l2 = ppo.gradients(b2)
    assert float((g1 - g2).norm() / g1.norm()) < 1e-5 and torch.allclose(l1, l2, rtol=1e-6, atol=1e-9)
    before = ppo._flat.clone()
    out = ppo.iterate(K)
    assert all(np.isfinite(out["losses"])) and np.isfinite(out["mean_reward"]) and not torch.equal(before, ppo._flat)
    assert out["losses"][-1] < out["losses"][0]
    assert torch.equal(env._actor[1]["w1"], ppo.policy.actor[0].weight.detach())
    out2 = ppo.iterate(K)
    assert all(np.isfinite(out2["losses"]))
    # the same iteration on a sensor handle: rollout, values and update all see the SENSED observation
    envs = BatchedQuad(1024, 0.01, 1000, training=True, direct_control=1, T=5, precision="f32", async_reset=True, sensor_noise=True, seed=2, device=dev)
    envs.reset()
    ppos = P.BatchedPPO(envs, hidden=128, K_epochs=2, seed=1)
    bs = ppos.collect(16)
    assert bs["obs"].shape == (16, 14, 1024) and bool(torch.isfinite(bs["obs"]).all()) and bool(torch.isfinite(bs["value"]).all())
    outs = ppos.update(bs)
    assert all(np.isfinite(outs))
