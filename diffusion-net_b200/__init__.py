"""B200-native DiffusionNetBlock hot path behind the reference module API.

Public surface mirrors ``diffusion_net`` (reference ``src/diffusion_net/__init__.py:1-3``):
``layers`` (DiffusionNet, DiffusionNetBlock, LearnedTimeDiffusion,
SpatialGradientFeatures, MiniMLP) and ``geometry`` (to_basis / from_basis).
"""
__version__ = "0.1.0"
