"""The oracle's `sensor` restatement (oracle/quad_oracle.py:SensorOracle) against the REFERENCE class driven by a replayed
draw stream (tests/golden/sensor_vectors.npz, oracle/gen_golden.py:gen_sensor_vectors): every method, TRIAD, GPS, the bias
drift and the GPS blend of visual_landing/math_trajectory.py:71-77 at 1e-12 — CPU only."""
import numpy as np

from conftest import load_golden, rel_err
from oracle import quad_oracle as qo

TOL = 1e-12


def _oracle(g, pre, n, **kw):
    sen = qo.SensorOracle(n, 0.01, **kw)
    sen.reset_u(g[pre + "u"][:, 3:6], g[pre + "reset_state"])          # draws 0..2 were consumed by sensor.__init__'s bias_reset
    return sen


def _state_rows(sen):
    return np.concatenate([sen.a_b[:, None], sen.g_b[:, None], sen.a_b_d[:, None], sen.g_b_d[:, None], sen.vel, sen.pos, sen.quat,
                           sen.R[:, :, 2], sen.acc0], axis=1)


def test_every_method_in_canonical_order_vs_reference():
    g = load_golden("sensor_vectors.npz")
    K, n = g["A_z"].shape[:2]
    sen = _oracle(g, "A_", n)
    for t in range(K):
        z, y, ar, rot, fm = g["A_z"][t], g["A_state"][t], g["A_acc_read"][t], g["A_mat_rot"][t], g["A_f_m"][t]
        acc, vel, pos = sen.accel_int(z[:, 0:9], ar, rot, fm)
        assert rel_err(np.concatenate([acc, vel, pos], axis=1), g["A_accel_int"][t]) < TOL
        assert rel_err(sen.gyro_int(z[:, 9:12], y), g["A_gyro_int"][t]) < TOL
        assert rel_err(sen.gyro(z[:, 12:15], y), g["A_gyro"][t]) < TOL
        pg, vg = sen.gps(z[:, 15:21], y)
        assert rel_err(np.concatenate([pg, vg], axis=1), g["A_gps"][t]) < TOL
        q, R = sen.triad(z[:, 21:27], ar, rot, fm)
        assert rel_err(R, g["A_triad_R"][t]) < TOL
        assert rel_err(q, g["A_triad_q"][t]) < 1e-11                     # SciPy's matrix -> quaternion conversion
        assert rel_err(_state_rows(sen), g["A_sens_state"][t], floor=1e-6) < 1e-10, t


def test_canonical_step_composition_and_gps_blend_vs_reference_sensor_sp():
    g = load_golden("sensor_vectors.npz")
    for pre, blend in (("A_", 0.0), ("B_", float(g["B_gps_blend"]))):
        K, n = g[pre + "z"].shape[:2]
        sen = _oracle(g, pre, n, gps_blend=blend)
        for t in range(K):
            obs = sen.step(g[pre + "z"][t], g[pre + "state"][t], g[pre + "acc_read"][t], g[pre + "mat_rot"][t], g[pre + "f_m"][t])
            assert rel_err(obs, g[pre + "obs"][t]) < TOL, (pre, t)
            assert rel_err(_state_rows(sen), g[pre + "sens_state"][t], floor=1e-6) < 1e-10, (pre, t)
    # the blend really acted: position estimate pulled towards the (noisy) GPS reading
    assert np.abs(g["B_obs"][-1][:, 0:6:2] - g["B_state"][-1][:, 0:5:2]).max() > 0.05


def test_methods_in_another_order_vs_reference():
    g = load_golden("sensor_vectors.npz")
    K, n = g["C_z"].shape[:2]
    sen = _oracle(g, "C_", n)
    for t in range(K):
        z, y, ar, rot, fm = g["C_z"][t], g["C_state"][t], g["C_acc_read"][t], g["C_mat_rot"][t], g["C_f_m"][t]
        assert rel_err(sen.gyro(z[:, 0:3], y), g["C_gyro"][t]) < TOL
        q, R = sen.triad(z[:, 3:9], ar, rot, fm)
        assert rel_err(R, g["C_triad_R"][t]) < TOL and rel_err(q, g["C_triad_q"][t]) < 1e-11
        pg, vg = sen.gps(z[:, 9:15], y)
        assert rel_err(np.concatenate([pg, vg], axis=1), g["C_gps"][t]) < TOL
        assert rel_err(sen.gyro_int(z[:, 15:18], y), g["C_gyro_int"][t]) < TOL
        assert rel_err(sen.accel(z[:, 18:21], ar), g["C_accel"][t]) < TOL
        acc, vel, pos = sen.accel_int(z[:, 21:30], ar, rot, fm)
        assert rel_err(np.concatenate([acc, vel, pos], axis=1), g["C_accel_int"][t]) < TOL
        assert rel_err(_state_rows(sen), g["C_sens_state"][t], floor=1e-6) < 1e-10, t


def test_oracle_dynamics_reproduce_the_fixture_truth():
    """The true quantities the sensor reads (state, accelerometer_read, mat_rot, f_in/M) are what the quad oracle produces."""
    g = load_golden("sensor_vectors.npz")
    K, n = g["A_z"].shape[:2]
    ora = qo.BatchQuadOracle(n, 0.01, 10 ** 6, training=False, direct_control=1, T=2, integrator="rk45")
    ora.reset(g["A_init"])
    assert rel_err(ora.state, g["A_reset_state"]) < 1e-9
    for t in range(K):
        ora.step(g["A_actions"][t])
        assert rel_err(ora.state, g["A_state"][t]) < 1e-9 and rel_err(ora.accelerometer_read, g["A_acc_read"][t]) < 1e-9
        assert rel_err(ora.mat_rot, g["A_mat_rot"][t]) < 1e-9 and rel_err(ora.f_in / qo.M, g["A_f_m"][t]) < 1e-12
