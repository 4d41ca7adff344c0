"""LinearFeatureBaseline (rllab): ridge regression of returns on
[o, o^2, t/100, (t/100)^2, (t/100)^3, 1] with o = clip(obs, -10, 10).  Host implementation used by
BaseSampler.process_samples (samplers/base.py:55,167)."""
import numpy as np


class LinearFeatureBaseline:
    def __init__(self, env_spec=None, reg_coeff=1e-5):
        self._coeffs = None
        self._coeffs_dev = None      # torch float64 [2S+4] when fitted on the device
        self._reg_coeff = reg_coeff

    # -- device-resident coefficients (process_samples_flat / metrpo_trpo_fit_baseline) ----------
    def device_coeffs(self, device):
        if self._coeffs_dev is None and self._coeffs is not None:
            import torch
            self._coeffs_dev = torch.as_tensor(self._coeffs, dtype=torch.float64).to(device)
        return self._coeffs_dev

    def set_device_coeffs(self, coeffs):
        self._coeffs_dev = coeffs
        self._coeffs = None          # host copy refreshed lazily by predict()

    @staticmethod
    def features(path):
        o = np.clip(path["observations"], -10, 10)
        L = len(path["rewards"])
        al = np.arange(L).reshape(-1, 1) / 100.0
        return np.concatenate([o, o ** 2, al, al ** 2, al ** 3, np.ones((L, 1))], axis=1)

    def fit(self, paths):
        F = np.concatenate([self.features(p) for p in paths])
        ret = np.concatenate([p["returns"] for p in paths])
        reg = self._reg_coeff
        self._coeffs_dev = None
        for _ in range(5):
            self._coeffs = np.linalg.lstsq(F.T.dot(F) + reg * np.identity(F.shape[1]), F.T.dot(ret),
                                           rcond=None)[0]
            if not np.any(np.isnan(self._coeffs)):
                break
            reg *= 10

    def predict(self, path):
        if self._coeffs is None and self._coeffs_dev is not None:
            self._coeffs = self._coeffs_dev.cpu().numpy()
        if self._coeffs is None:
            return np.zeros(len(path["rewards"]))
        return self.features(path).dot(self._coeffs)
