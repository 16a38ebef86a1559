"""super_diffusion_b200 — B200-native SuperDiff sampling path.

Host-side mirror of the reference's sampling API (mo-rsa24/super-diffusion:
cifar/dynamics.py, cifar/eval_utils.py, cifar/models/utils.py and the OR/AND
superposition loops of notebooks/superposition_edu.ipynb and
applications/images/clip_eval.py) over hand-written sm_100a CUDA kernels reached
through the C ABI of include/superdiff_b200.h.  No CPU fallback.
"""
from . import _lib, ops  # noqa: F401

__version__ = "0.1.0"
