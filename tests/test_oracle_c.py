"""CPU test: the plain-C restatement (oracle/quad_oracle.c, used as the CPU baseline) against the pinned NumPy oracle
and the reference-generated fixtures."""
import numpy as np
import pytest

from conftest import load_golden, rel_err
from oracle import quad_oracle as qo
from oracle.c_oracle import COracle


@pytest.mark.parametrize("name,direct,training", [("step_direct.npz", 1, True), ("step_indirect.npz", 0, True),
                                                   ("step_eval.npz", 1, False)])
def test_c_oracle_reproduces_reference_trajectories(name, direct, training):
    g = load_golden(name)
    n_env = g["init"].shape[0]
    env = COracle(n_env, 0.01, int(g["n"]), training=training, direct_control=direct, T=int(g["T"]), integrator="rk45")
    oh = env.reset(g["init"])
    assert rel_err(oh, g["reset_obs"]) < 1e-10
    worst = 0.0
    for t in range(g["actions"].shape[0]):
        obs, rew, done = env.step(g["actions"][t])
        ok = ~np.isnan(g["obs"][t]).any(axis=1)
        worst = max(worst, rel_err(obs[ok], g["obs"][t][ok]), rel_err(rew[ok], g["reward"][t][ok]))
        assert np.array_equal(done, g["done"][t])
        assert np.array_equal(env.nfev[ok], g["nfev"][t][ok])
    assert worst < 1e-9, worst


def test_c_oracle_rk4_matches_numpy_rk4():
    n = 64
    init, _ = qo.sample_reset_state(5, np.arange(n), 0)
    a = COracle(n, 0.01, 1000, integrator="rk4", substeps=2, threads=2)
    b = qo.BatchQuadOracle(n, 0.01, 1000, integrator="rk4", substeps=2)
    a.reset(init); b.reset(init)
    rng = np.random.default_rng(0)
    for _ in range(20):
        act = rng.uniform(-1, 1, (n, 4))
        oa, ra, da = a.step(act); ob, rb, db = b.step(act)
        assert rel_err(oa, ob) < 1e-11 and np.array_equal(da, db)
