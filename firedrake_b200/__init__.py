"""firedrake_b200 -- B200-native finite-element assembly engine.

Drop-in for ONE hot path of firedrakeproject/firedrake: the PyOP2 global kernel
(gather -> TSFC element kernel -> scatter-add) behind ``assemble()`` /
``par_loop()`` / ``ImplicitMatrixContext.mult()``.  See DESIGN.md.

Importing this package needs neither a GPU nor the shared library; any compute
entry point does, and raises :class:`EngineError` otherwise (no CPU fallback).
"""
from ._lib import EngineError, init, load  # noqa: F401

__version__ = "0.1.0"
