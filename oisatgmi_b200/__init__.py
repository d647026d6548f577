"""B200-native implementation of the OI-SAT-GMI satellite->model assimilation hot path.

Importing the package never initialises CUDA (the reference calls `interpolator`
inside joblib worker processes, /root/reference/oisatgmi/reader.py:1405); the
shared library is loaded on first use by `oisatgmi_b200._lib`.
"""
__version__ = "0.1.0"
