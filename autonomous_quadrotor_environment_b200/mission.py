"""Reference-trajectory (set-point) generator — the reference's ``mission`` class
(mission_control/mission_control.py:3-83) without its Python loops: the same arrays, produced by whole-array NumPy
operations whose evaluation order matches the loops (so the results are identical to the last bit), plus the glue that
hands a trajectory to the batched controllers (``BatchedQuad.control_rollout(..., target_traj=...)``).

API and quirks kept as they are upstream:
  * ``gen_trajectory`` ramps from ``additive[0:3]`` (the first three entries of a 14-vector, i.e. x, vx, y — :13,:20) to
    ``position + additive[0:3]`` and then holds ``position`` WITHOUT the offset (:21); with ``velocity`` given, ``steps``
    must equal ``total_timesteps`` (the reference's slice assignment raises otherwise, :28);
  * ``sin_trajectory`` integrates z from the LAST row of the zero-initialised array, so z[0] = ascent_rate*dt (:42);
  * ``get_error`` keeps returning (and extrapolating, :70) the last point once the trajectory is exhausted.
"""
import numpy as np


class mission:
    def __init__(self, time_step):
        self.time_step = time_step

    # mission_control.py:7-31
    def gen_trajectory(self, total_timesteps, steps, position, velocity=None, additive=None):
        self.trajectory_step = 0
        self.trajectory_total_steps = steps
        initial_state = np.zeros(14) if additive is None else np.asarray(additive, dtype=np.float64)
        position = np.asarray(position, dtype=np.float64)
        self.trajectory = np.zeros([total_timesteps, 3])
        self.velocity = np.zeros([total_timesteps, 3])
        if velocity is None:
            for i in range(3):
                self.trajectory[:steps, i] = np.linspace(initial_state[i], position[i] + initial_state[i], steps)
                self.trajectory[steps:, i] = position[i]
            if steps > 1:
                self.velocity[1:steps] = (self.trajectory[1:steps] - self.trajectory[0:steps - 1]) / self.time_step
        else:
            velocity = np.asarray(velocity, dtype=np.float64)
            if steps != total_timesteps:
                raise ValueError("could not broadcast input array from shape (%d,) into shape (%d,)" % (steps, total_timesteps))
            for i in range(3):
                self.velocity[:, i] = np.linspace(0, velocity[i], steps)
            # trajectory[i+1] = trajectory[i] + velocity[i]*dt, i = 0 .. steps-2 (sequential sums, like the loop)
            self.trajectory[1:steps] = np.cumsum(self.velocity[0:steps - 1] * self.time_step, axis=0)

    # mission_control.py:33-46
    def sin_trajectory(self, steps, circular_rate, ascent_rate, center, axis):
        self.trajectory_step = 0
        self.trajectory_total_steps = steps
        center, axis = np.asarray(center, dtype=np.float64), np.asarray(axis, dtype=np.float64)
        self.trajectory_timesteps = np.arange(0, steps, 1)
        a = self.trajectory_timesteps * circular_rate * self.time_step
        self.trajectory = center[None, :] + np.sin(a)[:, None] * axis[None, :]
        self.trajectory[:, 2] = np.cumsum(np.full(steps, ascent_rate * self.time_step))   # z[k] = z[k-1] + c from z[-1] = 0
        self.velocity = np.zeros([steps, 3])
        self.velocity[1:] = (self.trajectory[1:] - self.trajectory[:-1]) / self.time_step

    # mission_control.py:48-65
    def spiral_trajectory(self, zsteps, steps, rate, circular_rate, radius, center):
        self.trajectory_step = 0
        self.trajectory_total_steps = steps
        center = np.asarray(center, dtype=np.float64)
        self.trajectory_timesteps = np.arange(0, steps, 1)
        t = self.trajectory_timesteps
        a = t * circular_rate * self.time_step
        z = np.where(t > zsteps, zsteps * rate * self.time_step, t * rate * self.time_step)
        xyz = np.stack([np.cos(a) * radius, np.sin(a) * radius, z], axis=1)
        self.trajectory = center[None, :] + xyz - np.array([radius, 0, 0])[None, :]
        self.velocity = np.zeros([steps, 3])
        self.velocity[1:] = (self.trajectory[1:] - self.trajectory[:-1]) / self.time_step

    # mission_control.py:68-83
    def get_error(self, time):
        if self.trajectory_step == self.trajectory_total_steps:
            self.trajectory[-1, :] = self.trajectory[-1, :] + self.velocity[-1, :] * self.time_step
            k = -1
        else:
            k = self.trajectory_step
            self.trajectory_step += 1
        tr, ve = self.trajectory[k], self.velocity[k]
        return np.array([tr[0], ve[0], tr[1], ve[1], tr[2], ve[2], 0, 0, 0, 0, 0, 0, 0, 0])

    # ---- glue to the batched controllers ---------------------------------------------------------------------
    def velocity_setpoints(self, horizon, device=None, dtype=None):
        """(horizon, 3) torch tensor of velocity set-points for ``BatchedQuad.control_rollout(target_traj=...)``: the rows
        ``get_error`` would return next (the last row repeats once the trajectory is exhausted); advances the pointer."""
        import torch
        k0 = self.trajectory_step
        idx = np.minimum(np.arange(k0, k0 + horizon), len(self.velocity) - 1)
        idx = np.minimum(idx, max(self.trajectory_total_steps - 1, 0))
        self.trajectory_step = min(k0 + horizon, self.trajectory_total_steps)
        return torch.as_tensor(self.velocity[idx], dtype=dtype or torch.float32, device=device).contiguous()
