# -*- coding: utf-8 -*-
"""B200-native drop-in for the ``solidspy_opt`` optimisers (SIMP / BESO / ESO on structured quad4 meshes).

Same entry points and argument order as the reference package
(``src/solidspy_opt/optimize.py:16,116,215,360``); the per-iteration numerics run in hand-written
sm_100a CUDA kernels behind the C ABI of ``libtopo_b200.so`` (``include/topo_b200.h``).
The Keras surrogates of the reference (``models.py``) are out of scope.
"""
__all__ = ["optimize", "Utils"]
__version__ = "0.1.0+b200"
