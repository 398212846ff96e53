"""cdae_b200 — B200-native training / scoring engine for the Collaborative Denoising
Auto-Encoder, behind the class surface of jasonyaw/CDAE (libcf::CDAE).

Python here is the test / benchmark mirror of the C++ host class in cdae_b200/host/ — both are
thin callers of the C ABI in include/cdae_b200.h, implemented by hand-written sm_100a kernels
in cdae_b200/csrc/."""
from ._lib import CdaeError  # noqa: F401
from .model import CDAE, CDAEConfig, CDAEGroup  # noqa: F401
from .data import Dataset  # noqa: F401
