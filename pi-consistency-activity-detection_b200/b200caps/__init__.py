"""b200caps -- B200-native (sm_100a) kernels behind a C ABI for the semi-supervised
video-action-detection training step.  See DESIGN.md."""
from . import _abi  # noqa: F401

__all__ = ["_abi"]
