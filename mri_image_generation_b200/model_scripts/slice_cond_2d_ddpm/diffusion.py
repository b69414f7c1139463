"""Drop-in for model_scripts/slice_cond_2d_ddpm/diffusion.py (GaussianDiffusion, linear betas)."""
import torch

from ... import _lib, schedules
from ...diffusion_base import DiffusionBase, _require_cuda


class GaussianDiffusion(DiffusionBase):
    """slice_cond_2d_ddpm/diffusion.py:5-49."""

    WITH_SNR = True

    def __init__(self, model, image_size, channels=1, timesteps=1000, beta_start=1e-4, beta_end=0.02):
        super().__init__()
        self.model = model
        self.image_size = image_size
        self.channels = channels
        self.timesteps = timesteps
        print(f"Setting up Gaussian Diffusion with {timesteps} timesteps.")
        betas = schedules.linear_beta_schedule(timesteps, beta_start, beta_end)
        for k, v in schedules.make_buffers(betas, with_snr=self.WITH_SNR).items():
            self.register_buffer(k, v)

    def _extract(self, a, t, x_shape):
        """diffusion.py:51-58."""
        B = t.shape[0]
        out = a.gather(-1, t)
        return out.view(B, 1, 1, 1).expand(x_shape)

    def q_sample(self, x_start, t, noise=None):
        """diffusion.py:60-75."""
        if noise is None:  # noise = torch.randn_like(x_start), drawn inside the kernel
            return self._q_sample_draw(x_start, t)[0]
        return self._q_sample(x_start, t, noise)

    def p_losses(self, x_start, t, cond=None, noise=None, min_snr_gamma=5.0):
        """The LIVE definition, diffusion.py:91-107 (the second `def p_losses` wins): min-SNR
        weighted per-sample MSE, `cond` passed as the model's third positional argument (the
        training script passes z_pos there, model.py:164).  The reference's hard-coded
        mean(dim=(1,2,3,4)) raises IndexError on 4-D slices; this computes the same
        quantity rank-generically instead of reproducing the crash."""
        if noise is None:  # noise = torch.randn_like(x_start), drawn inside the q_sample kernel
            x_noisy, noise = self._q_sample_draw(x_start, t)
        else:
            x_noisy = self.q_sample(x_start=x_start, t=t, noise=noise)
        predicted_noise = self.model(x_noisy, t) if cond is None else self.model(x_noisy, t, cond)
        return self._loss(predicted_noise, noise, t, float(min_snr_gamma))

    @torch.no_grad()
    def p_sample(self, x, t, z_pos):
        """diffusion.py:110-132."""
        _require_cuda(x, "p_sample")
        eng = self._engine_model()
        if eng is not None:
            prog = eng.program(x.shape[0], x.shape[2:], x.shape[1], 0)
            prog.z_in.copy_(self._z_tensor(z_pos, x.shape[0], x.device).reshape(-1, 1))
            return self._p_sample_on(prog, x, t)
        eps_theta = self.model(x, t, z_pos)
        return self._p_update(x, t, eps_theta)  # z = randn_like(x) drawn inside the kernel

    def _z_tensor(self, z_pos, B, device):
        if not torch.is_tensor(z_pos):
            return torch.full((B,), float(z_pos), device=device)
        return z_pos.to(device).float()

    @torch.no_grad()
    def p_sample_loop(self, shape, z_pos):
        """diffusion.py:134-155."""
        device = self.betas.device
        B = shape[0]
        img = self._randn(shape, device)
        z_pos = self._z_tensor(z_pos, B, device)
        eng = self._engine_model()
        if eng is not None:
            _require_cuda(img, "p_sample_loop")
            prog = eng.program(B, shape[2:], shape[1], 0)
            prog.z_in.copy_(z_pos.reshape(-1, 1))
            return self._reverse_loop(prog, img, self.timesteps - 1, self.timesteps, "ddpm")
        for i in reversed(range(self.timesteps)):
            t = torch.full((B,), i, device=device, dtype=torch.long)
            img = self.p_sample(img, t, z_pos)
        return img

    @torch.no_grad()
    def sample(self, batch_size=16, z_pos=0.5):
        """diffusion.py:157-166."""
        return self.p_sample_loop(
            (batch_size, self.channels, self.image_size, self.image_size), z_pos=z_pos)
