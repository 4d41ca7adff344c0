"""Host helpers of the hot path mirrored from the reference's utils.py."""
import numpy as np


def stop_critereon(threshold, offset, percent_models_threshold=0.5):
    """utils.py:285-296.  scalar mode: relative increase above `threshold`; vector mode (one entry
    per dynamics model): stop when the FRACTION of models whose cost got worse exceeds
    percent_models_threshold (params 'policy_opt_params.stop_critereon')."""
    def f(loss_old, loss_new, mode="scalar"):
        if mode == "scalar":
            assert not hasattr(loss_new, "__iter__")
            return (loss_new - loss_old) / (np.abs(loss_old) + offset) > threshold
        assert mode == "vector"
        assert isinstance(loss_new, np.ndarray)
        out = loss_new > loss_old
        return np.mean(out) > percent_models_threshold
    return f
