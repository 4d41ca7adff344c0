"""me_trpo_b200 -- B200-native ME-TRPO inner loop (imaginary ensemble rollout + policy update).

Host-side mirror of the reference's sockets for the hot path over the C-ABI
library libmetrpo.so (include/metrpo.h).  The CUDA library loads lazily on first use and there
is no CPU fallback.
"""
__version__ = "0.1.0"
