"""otgan_b200 -- B200-native (sm_100a CUDA behind a C ABI) implementation of the OT-GAN training hot path.

Package layout mirrors the reference checkout so that `from utils import matching` becomes
`from otgan_b200.utils import matching`:
    utils/matching.py           <- utils/matching.py            (cosine-cost blocks, Sinkhorn, matched features, distance)
    toy_example/matching_cpu.py <- toy_example/matching_cpu.py  (tensor API, squared-Euclidean/n cost)
    csrc/ + libotgan.so         the CUDA kernels and the C ABI (include/otgan.h)
"""
__version__ = "0.1.0"
