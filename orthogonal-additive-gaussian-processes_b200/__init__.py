"""B200-native OAK hot path: the orthogonal additive kernel Gram / cross-covariance, the SGPR
statistics and the Sobol-index building blocks as hand-written sm_100a CUDA behind the reference's
gpflow-Kernel-shaped API.  Import as ``oak_b200`` (this directory's name is not a valid Python
identifier; ``oak_b200/__init__.py`` at the repository root is the import alias).
"""
__version__ = "0.1.0"
