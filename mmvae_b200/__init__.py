"""mmvae_b200 -- B200-native (sm_100a) implementation of the CMMVAE training step.

Host side: Python/PyTorch mirroring the reference's ``src/cmmvae`` module API.
Device side: hand-written CUDA kernels in ``csrc/`` behind the C ABI of ``include/cmmvae_b200.h``.
"""
__version__ = "0.1.0"
