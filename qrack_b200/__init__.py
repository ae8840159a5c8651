"""qrack_b200 — B200-native state-vector engine behind Qrack's QEngine hot path (Apply2x2/ApplyM/Compose/Decompose/Prob).

Layout: ``csrc/`` hand-written sm_100a CUDA kernels + the C ABI of ``include/b200sv.h`` (built in-tree into
``libb200sv.so``); ``qengine.py`` the host-side mirror of the reference's QEngine interface; ``qscript.py`` the
circuit-script format shared with the oracle and the reference harness.
"""
from .qengine import QEngineCUDA, QEngineHost  # noqa: F401
from .qcircuit import QCircuit  # noqa: F401
from . import qscript  # noqa: F401

__all__ = ["QEngineCUDA", "QEngineHost", "QCircuit", "qscript"]
