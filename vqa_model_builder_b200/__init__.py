"""vqa_model_builder_b200 — B200-native (sm_100a) fusion + MOE hot path of the AutoViVQA model builder.

Drop-in nn.Modules with the reference's constructor / forward / state_dict contract
(src/modeling/{moe,fusion,meta_arch}) over hand-written CUDA kernels reached through a C-ABI shared
library (include/b200vqa.h).  There is no CPU fallback.
"""
from .runtime import (get_compute_dtype_mode, invalidate_all, prefetch_compute_weights, resolve_compute_dtype,  # noqa: F401
                      set_compute_dtype)

__version__ = "0.1.0"
