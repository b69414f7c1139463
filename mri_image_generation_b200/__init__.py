"""B200-native (sm_100a) diffusion hot path for NickB42/mri-image-generation.

Host side: Python classes mirroring the reference's model / diffusion API
(mri_image_generation_b200.model_scripts.*); compute: hand-written CUDA kernels in
libmri_b200.so reached through the C ABI in include/mri_b200.h.  No CPU fallback.
"""
__version__ = "0.1.0"
