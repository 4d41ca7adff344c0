"""GaussianMLPPolicy: the subset of rllab's policy the hot path touches (SURVEY.md Appendix A.1;
reference call sites training.py:81-117, samplers/vectorized_sampler.py:62-65).

Parameters live as torch tensors on the rollout device.  `get_actions` serves the step-granular
socket (B1); the fused sampler (B2) hands the same tensors to the kernel, which evaluates the
mean network itself."""
import numpy as np
import torch


class GaussianMLPPolicy:
    vectorized = True
    recurrent = False

    def __init__(self, obs_dim, action_dim, hidden_sizes=(32, 32), init_std=1.0,
                 output_tanh=False, device="cuda", seed=0):
        self.obs_dim, self.action_dim = int(obs_dim), int(action_dim)
        self.hidden_sizes = tuple(hidden_sizes)
        self.output_tanh = bool(output_tanh)
        self.device = torch.device(device)
        dims = [self.obs_dim] + list(self.hidden_sizes) + [self.action_dim]
        g = torch.Generator().manual_seed(seed)
        self.W, self.b = [], []
        for i in range(len(dims) - 1):   # Xavier-uniform W, zero b (rllab MLP defaults)
            lim = float(np.sqrt(6.0 / (dims[i] + dims[i + 1])))
            self.W.append(((torch.rand(dims[i], dims[i + 1], generator=g) * 2 - 1) * lim).to(self.device))
            self.b.append(torch.zeros(dims[i + 1], device=self.device))
        self.log_std = torch.full((self.action_dim,), float(np.log(init_std)), device=self.device)

    # -- rllab Parameterized surface used by the optimizer socket (B3) -----------------------
    def get_params(self, trainable=True):
        out = []
        for w, b in zip(self.W, self.b):
            out += [w, b]
        return out + [self.log_std]

    def get_param_values(self, trainable=True):
        return torch.cat([p.reshape(-1) for p in self.get_params()]).cpu().numpy()

    def set_param_values(self, flat, trainable=True):
        flat = torch.as_tensor(np.asarray(flat), dtype=torch.float32, device=self.device)
        o = 0
        for p in self.get_params():
            n = p.numel()
            p.copy_(flat[o:o + n].reshape(p.shape))
            o += n
        assert o == flat.numel()

    def flat_params(self):
        """Flat fp32 device vector in get_params(trainable=True) order (what the optimizer updates)."""
        return torch.cat([p.reshape(-1) for p in self.get_params()]).contiguous()

    def set_flat_params(self, flat):
        o = 0
        for p in self.get_params():
            n = p.numel()
            p.copy_(flat[o:o + n].reshape(p.shape))
            o += n
        assert o == flat.numel()

    def reset(self, dones=None):
        pass   # feed-forward policy: no state

    # -- forward -----------------------------------------------------------------------------
    def mean_and_log_std(self, obs):
        h = torch.as_tensor(obs, dtype=torch.float32, device=self.device)
        n = len(self.W)
        for i in range(n):
            h = h @ self.W[i] + self.b[i]
            if i < n - 1 or self.output_tanh:
                h = torch.tanh(h)
        log_std = torch.clamp(self.log_std, min=float(np.log(1e-6))).expand_as(h)
        return h, log_std

    def get_actions(self, observations, rng=None):
        """actions = rnd * exp(log_std) + mean, rnd ~ N(0,1) from NumPy like the reference."""
        mean, log_std = self.mean_and_log_std(observations)
        rng = np.random if rng is None else rng
        rnd = rng.normal(size=tuple(mean.shape)).astype(np.float32)
        mean_np, ls_np = mean.cpu().numpy(), log_std.cpu().numpy()
        return rnd * np.exp(ls_np) + mean_np, dict(mean=mean_np, log_std=ls_np)
