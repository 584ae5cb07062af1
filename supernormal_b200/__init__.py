"""supernormal_b200 -- B200-native (sm_100a) implementation of SuperNormal's patch-based NeuS
training hot path, behind the nerfacc-0.3.5 / tiny-cuda-nn API surface that the reference's
models/renderer.py and models/fields.py call.  Host side: Python/PyTorch.  Device side:
hand-written CUDA kernels reached only through the C ABI of include/snb200.h (libsnb200.so).
No Triton, no multi-backend dispatch, no CPU fallback.
"""
__version__ = "0.1.0"
