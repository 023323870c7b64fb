"""scamlgp_b200 -- B200-native hot path of ScaML-GP (Scalable Meta-Learning with Gaussian Processes).

Drop-in for the data-parallel path of the reference (`scamlgp/model.py`,
`scamlgp/optimizer.py`, `scamlgp/utils.py`): fitting and querying the per-task base GPs
plus the target GP.  Host code is Python/torch; the arithmetic runs in hand-written
sm_100a kernels behind the C ABI of `include/scaml_b200.h` (csrc/libscaml_b200.so).
"""
from ._capi import (  # noqa: F401
    KERNEL_MATERN12,
    KERNEL_MATERN32,
    KERNEL_MATERN52,
    KERNEL_RBF,
    PRIOR_GAMMA,
    PRIOR_LOGNORMAL,
    PRIOR_NONE,
    HyperSpec,
    ScamlError,
    load_cuda_library,
)

__version__ = "0.1.0"
