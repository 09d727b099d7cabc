"""dwc_gan_b200: B200-native (sm_100a) generator+discriminator training step of DWC-GAN behind the
reference's Solver / nn.Module API.  The CUDA kernels live in libdwc_b200.so (C ABI in
include/dwc_b200.h); there is no CPU fallback."""
from .ops import RT  # noqa: F401


def set_mode(mode: str):
    """'bf16' (tcgen05 product path), 'bf16_simt' (same numerics on CUDA cores), 'fp32' (validation mode)."""
    RT.set_mode(mode)
