"""numpyro_b200 -- B200-native batched NUTS/HMC engine behind numpyro's MCMC API.

Only the MCMC hot path is implemented; everything computes inside the CUDA
library ``csrc/libb200nuts.so`` and there is no CPU fallback.
"""
__version__ = "0.1.0"
