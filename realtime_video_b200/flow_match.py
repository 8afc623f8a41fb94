"""Flow-matching noise schedule, inference subset — used only when the reference's own ``utils/scheduler.py`` is
not importable (stand-alone runs of this package: bench.py and the tests on a box without the reference checkout).
When the package is dropped into the reference, ``realtime_video_b200.wan_wrapper`` imports the reference's
``utils.scheduler.FlowMatchScheduler`` instead and this module is not used.

Definition (reference utils/scheduler.py:118-141, :159-176): with shift s,
    sigma'_i = linspace(sigma_start, sigma_min, n [+1 and drop the last if extra_one_step])
    sigma_i  = s * sigma'_i / (1 + (s - 1) * sigma'_i),      timestep_i = 1000 * sigma_i
    add_noise(x0, eps, t) = (1 - sigma_j) * x0 + sigma_j * eps,  j = argmin_i |timestep_i - t|, cast to eps.dtype
Everything is fp32 torch arithmetic on a [F, 16, h, w] latent: host-side plumbing, not a kernel.
"""
from __future__ import annotations

import torch


class FlowMatchSchedule:
    def __init__(self, num_inference_steps: int = 100, num_train_timesteps: int = 1000, shift: float = 3.0,
                 sigma_max: float = 1.0, sigma_min: float = 0.003 / 1.002, extra_one_step: bool = False, **unused):
        if any(unused.get(k) for k in ("inverse_timesteps", "reverse_sigmas")):
            raise NotImplementedError("reversed schedules are training-side options outside the hot path")
        self.num_train_timesteps, self.shift = num_train_timesteps, shift
        self.sigma_max, self.sigma_min, self.extra_one_step = sigma_max, sigma_min, extra_one_step
        self.set_timesteps(num_inference_steps)

    def set_timesteps(self, num_inference_steps: int = 100, denoising_strength: float = 1.0, training: bool = False):
        top = self.sigma_min + (self.sigma_max - self.sigma_min) * denoising_strength
        n = num_inference_steps + (1 if self.extra_one_step else 0)
        lin = torch.linspace(top, self.sigma_min, n)[:num_inference_steps]
        self.sigmas = self.shift * lin / (1 + (self.shift - 1) * lin)
        self.timesteps = self.sigmas * self.num_train_timesteps

    def sigma_at(self, timestep: torch.Tensor) -> torch.Tensor:
        """sigma of the tabulated timestep nearest to each entry of ``timestep`` ([N] or [B, F])."""
        t = timestep.flatten()
        self.sigmas, self.timesteps = self.sigmas.to(t.device), self.timesteps.to(t.device)
        return self.sigmas[(self.timesteps[None, :] - t[:, None]).abs().argmin(dim=1)]

    def add_noise(self, original_samples: torch.Tensor, noise: torch.Tensor, timestep: torch.Tensor) -> torch.Tensor:
        sigma = self.sigma_at(timestep.to(noise.device)).reshape(-1, 1, 1, 1)
        return ((1 - sigma) * original_samples + sigma * noise).type_as(noise)
