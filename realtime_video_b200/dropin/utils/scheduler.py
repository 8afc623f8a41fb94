"""Drop-in for ``utils/scheduler.py`` (reference utils/scheduler.py:106-194): flow-matching
noise schedule.  Tiny elementwise torch arithmetic on [F,16,h,w] latents — host-side plumbing
kept in torch with the reference's exact rounding (fp32 sigma, result cast to the noise dtype)."""
import torch


class SchedulerInterface:
    """Conversions of reference utils/scheduler.py:5-100 that apply to flow matching."""

    def convert_x0_to_noise(self, x0, xt, timestep):
        raise NotImplementedError("alphas_cumprod-based conversions are not defined for FlowMatchScheduler")

    convert_noise_to_x0 = convert_x0_to_noise
    convert_velocity_to_x0 = convert_x0_to_noise


class FlowMatchScheduler:
    def __init__(self, num_inference_steps=100, num_train_timesteps=1000, shift=3.0, sigma_max=1.0,
                 sigma_min=0.003 / 1.002, inverse_timesteps=False, extra_one_step=False,
                 reverse_sigmas=False):
        self.num_train_timesteps = num_train_timesteps
        self.shift, self.sigma_max, self.sigma_min = shift, sigma_max, sigma_min
        self.inverse_timesteps, self.extra_one_step, self.reverse_sigmas = \
            inverse_timesteps, extra_one_step, reverse_sigmas
        self.set_timesteps(num_inference_steps)

    def set_timesteps(self, num_inference_steps=100, denoising_strength=1.0, training=False):
        """scheduler.py:118-141."""
        sigma_start = self.sigma_min + (self.sigma_max - self.sigma_min) * denoising_strength
        if self.extra_one_step:
            self.sigmas = torch.linspace(sigma_start, self.sigma_min, num_inference_steps + 1)[:-1]
        else:
            self.sigmas = torch.linspace(sigma_start, self.sigma_min, num_inference_steps)
        if self.inverse_timesteps:
            self.sigmas = torch.flip(self.sigmas, dims=[0])
        self.sigmas = self.shift * self.sigmas / (1 + (self.shift - 1) * self.sigmas)
        if self.reverse_sigmas:
            self.sigmas = 1 - self.sigmas
        self.timesteps = self.sigmas * self.num_train_timesteps
        if training:
            x = self.timesteps
            y = torch.exp(-2 * ((x - num_inference_steps / 2) / num_inference_steps) ** 2)
            y_shifted = y - y.min()
            self.linear_timesteps_weights = y_shifted * (num_inference_steps / y_shifted.sum())

    def _sigma(self, timestep, device):
        if timestep.ndim == 2:
            timestep = timestep.flatten(0, 1)
        self.sigmas = self.sigmas.to(device)
        self.timesteps = self.timesteps.to(device)
        idx = torch.argmin((self.timesteps.unsqueeze(0) - timestep.unsqueeze(1)).abs(), dim=1)
        return idx, self.sigmas[idx].reshape(-1, 1, 1, 1)

    def step(self, model_output, timestep, sample, to_final=False):
        """scheduler.py:143-157."""
        idx, sigma = self._sigma(timestep, model_output.device)
        if to_final or (idx + 1 >= len(self.timesteps)).any():
            sigma_ = 1 if (self.inverse_timesteps or self.reverse_sigmas) else 0
        else:
            sigma_ = self.sigmas[idx + 1].reshape(-1, 1, 1, 1)
        return sample + model_output * (sigma_ - sigma)

    def add_noise(self, original_samples, noise, timestep):
        """scheduler.py:159-176: (1 - sigma) x0 + sigma * noise, cast to the noise dtype."""
        _, sigma = self._sigma(timestep, noise.device)
        return ((1 - sigma) * original_samples + sigma * noise).type_as(noise)

    def training_target(self, sample, noise, timestep):
        return noise - sample

    def training_weight(self, timestep):
        if timestep.ndim == 2:
            timestep = timestep.flatten(0, 1)
        self.linear_timesteps_weights = self.linear_timesteps_weights.to(timestep.device)
        idx = torch.argmin((self.timesteps.unsqueeze(1) - timestep.unsqueeze(0)).abs(), dim=0)
        return self.linear_timesteps_weights[idx]
