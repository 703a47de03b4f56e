"""ORACLE (test infrastructure): restatement of the reference ``Diffusion`` class (train.py:174-251) and of the
DDIM eta=0 sampler the benchmark configs name (not present in the reference -- specified here in fp64).

``train.py`` cannot be imported (module-level ``diffusers`` import, writes JSON into the cwd), so these ~40 lines
are restated.  Pinned by tests/test_oracle_golden.py against closed forms and against a literal transcription run
inside oracle/make_golden.py."""
import torch


class DiffusionOracle:
    def __init__(self, noise_steps=1000, beta_start=1e-4, beta_end=0.02):
        """train.py:175-188."""
        self.noise_steps = noise_steps
        self.beta = torch.linspace(beta_start, beta_end, noise_steps)
        self.alpha = 1.0 - self.beta
        self.alpha_hat = torch.cumprod(self.alpha, dim=0)

    def noise_images(self, x, t, eps):
        """train.py:190-194 with the noise passed in."""
        sa = torch.sqrt(self.alpha_hat[t])[:, None, None, None]
        s1 = torch.sqrt(1 - self.alpha_hat[t])[:, None, None, None]
        return sa * x + s1 * eps

    def ddpm_step(self, x, eps, i, noise):
        """train.py:229-236: x <- 1/sqrt(a) (x - (1-a)/sqrt(1-ah) eps) + sqrt(b) z ; z = 0 when i == 1."""
        n = x.shape[0]
        t = (torch.ones(n) * i).long()
        alpha = self.alpha[t][:, None, None, None]
        alpha_hat = self.alpha_hat[t][:, None, None, None]
        beta = self.beta[t][:, None, None, None]
        z = noise if i > 1 else torch.zeros_like(x)
        return 1 / torch.sqrt(alpha) * (x - ((1 - alpha) / (torch.sqrt(1 - alpha_hat))) * eps) + torch.sqrt(beta) * z

    def ddpm_sample(self, eps_fn, x_T, noises):
        """train.py:217-236 with one UNet call per step.  noises[i] is the z used at step i (i = T-1 .. 2)."""
        x = x_T.clone()
        for i in reversed(range(1, self.noise_steps)):
            t = (torch.ones(x.shape[0]) * i).long()
            eps = eps_fn(x, t)
            x = self.ddpm_step(x, eps, i, noises[i] if i > 1 else None)
        return x

    # ---- DDIM, eta = 0, evenly strided timesteps (specification of this build; fp64 arithmetic) ----
    def ddim_timesteps(self, num_steps):
        stride = self.noise_steps // num_steps
        ts = list(range(0, self.noise_steps, stride))[:num_steps]
        return ts[::-1]

    def ddim_step(self, x, eps, t, t_prev):
        ah = self.alpha_hat.double()
        a_t = ah[t]
        a_p = ah[t_prev] if t_prev >= 0 else torch.tensor(1.0, dtype=torch.float64)
        x0 = (x.double() - torch.sqrt(1 - a_t) * eps.double()) / torch.sqrt(a_t)
        return (torch.sqrt(a_p) * x0 + torch.sqrt(1 - a_p) * eps.double()).float()

    # ---- reduced-call ("stale eps") sampler of the reference's production generator ----
    @staticmethod
    def reduced_call_predicate(i, noise_steps):
        """regenerateFromtrain2.py:536 with fullSampling = 0: the UNet is evaluated when
        ``i % 100 == 0 or i % 5 == 0 or i == T or i == T - 1`` (the epoch-dependent terms select multiples of 25 / 15 / 10,
        all multiples of 5 already; ``epoch>50==0`` is a chained comparison that is always False); every other step
        re-uses the last predicted noise."""
        return i % 100 == 0 or i % 5 == 0 or i == noise_steps or i == noise_steps - 1

    def reduced_call_sample(self, eps_fn, x_T):
        """regenerateFromtrain2.py:520-618, fullSampling = 0: stale predicted noise between evaluations and the NOISE-FREE
        update ``x <- 1/sqrt(a) (x - (1-a)/sqrt(1-ah) eps)`` (:615-618).  Returns (x_0, list of evaluated timesteps)."""
        x = x_T.clone()
        eps = None
        called = []
        for i in reversed(range(1, self.noise_steps)):
            t = (torch.ones(x.shape[0]) * i).long()
            if self.reduced_call_predicate(i, self.noise_steps):
                eps = eps_fn(x, t)
                called.append(i)
            alpha = self.alpha[t][:, None, None, None]
            alpha_hat = self.alpha_hat[t][:, None, None, None]
            x = 1 / torch.sqrt(alpha) * (x - ((1 - alpha) / (torch.sqrt(1 - alpha_hat))) * eps)
        return x, called
