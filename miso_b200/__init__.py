"""miso_b200 — B200 (sm_100a) implementation of the detection post-processing and
region-feature hot path behind miso's torchvision Faster/Mask R-CNN pipelines.

Host side: Python over PyTorch tensors (device memory, streams, torch.distributed).
Compute: hand-written CUDA kernels in libmisob200.so behind the C ABI of include/misob200.h.
There is no CPU fallback; operators raise if the CUDA library is not built.
"""
from ._lib import MisoB200Error, load as load_library  # noqa: F401

__version__ = "0.1.0"
