"""LinearFeatureBaseline (rllab): ridge regression of returns on
[o, o^2, t/100, (t/100)^2, (t/100)^3, 1] with o = clip(obs, -10, 10).  Host implementation used by
BaseSampler.process_samples (samplers/base.py:55,167)."""
import numpy as np


class LinearFeatureBaseline:
    def __init__(self, reg_coeff=1e-5):
        self._coeffs = None
        self._reg_coeff = reg_coeff

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
        for _ in range(5):
            self._coeffs = np.linalg.lstsq(F.T.dot(F) + reg * np.identity(F.shape[1]), F.T.dot(ret),
                                           rcond=None)[0]
            if not np.any(np.isnan(self._coeffs)):
                break
            reg *= 10

    def predict(self, path):
        if self._coeffs is None:
            return np.zeros(len(path["rewards"]))
        return self.features(path).dot(self._coeffs)
