"""ct_pvae_b200 -- B200-native (sm_100a) differentiable Radon path for CT_PVAE.

Drop-in modules (same names and signatures as the reference's ``ctvae`` package):
  ct_pvae_b200.forward_functions : pad_phantom, project_tf_low_mem, project_tf_fast
  ct_pvae_b200.fbp_tensorflow    : iradon
The kernels live in ``csrc/`` and are reached through the C ABI of
``include/ctradon.h`` (``libctradon.so``, ctypes + zero-copy DLPack).
"""
from .forward_functions import backproject, num_proj_pix, pad_phantom, project_tf_fast, project_tf_low_mem  # noqa: F401
from .fbp_tensorflow import get_fourier_filter, iradon  # noqa: F401
from .likelihood import calculate_log_prob_M_given_R, log_prob_M_given_R_sum  # noqa: F401

__version__ = "0.2.0"
