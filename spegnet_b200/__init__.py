"""spegnet_b200 -- B200-native (sm_100a) implementation of SPEGNet's inference forward pass.

``SPEGNet`` is a drop-in for the reference ``models/spegnet.py::SPEGNet``;
``ops`` exposes the individual C-ABI kernels; ``_lib`` is the ctypes binding of libspegnet_b200.so.
"""
from .model import SPEGNet  # noqa: F401
from .pipeline import HostPipeline  # noqa: F401

__all__ = ["SPEGNet", "HostPipeline"]
__version__ = "0.1.0"
