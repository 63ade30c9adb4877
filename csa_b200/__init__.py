"""csa_b200 -- B200-native rotation finder for circular DNA sequence sets: the `./CSA R` hot path
of fjdf/CSA as CUDA kernels for sm_100a behind a C ABI (include/csa_gpu.h).  No CPU fallback."""
from .api import Batch, CsaGpuError, RotationFinder, SetResult  # noqa: F401
