from .vectorized_sampler import VectorizedSampler  # noqa: F401
from .base import BaseSampler  # noqa: F401
