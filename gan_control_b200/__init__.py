"""gan_control_b200 -- the StyleGAN2 generator/discriminator hot path of amazon-science/gan-control,
rebuilt for NVIDIA B200 (sm_100a): hand-written CUDA (tcgen05/TMEM implicit-GEMM convolutions, TMA,
fused epilogues) behind the reference's operator signatures.  See DESIGN.md / INTEGRATION.md."""
from .install import install, install_augment  # noqa: F401

__all__ = ['install', 'install_augment']
