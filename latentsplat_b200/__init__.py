"""latentsplat_b200 -- B200-native (sm_100a) implementation of latentSplat's render hot path.

The product path is CUDA only: importing works anywhere (so that host logic can be tested),
but every compute entry point raises if libls_raster.so or a CUDA device is missing.
"""
__version__ = "0.1.0"
