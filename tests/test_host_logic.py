"""CPU tests of the host logic: qscript generators/parsers and the QEngine dispatch mirror (driven on the oracle
backend, since the CUDA backend needs a device)."""
import random

import numpy as np
import pytest

from oracle.restate_engine import QEngineRestate
from qrack_b200 import qscript

import util


def test_generators_are_deterministic_and_sized():
    a = qscript.random_htcnot(20, 40)
    assert a == qscript.random_htcnot(20, 40)
    assert qscript.count_gate_ops(a) == 1200           # BASELINE configs[0]
    assert qscript.count_gate_ops(qscript.random_htcnot(30, 40)) == 1800   # configs[1]
    q = qscript.qft(30)
    assert q.count("QFT 0 30") == 1


def test_matching_is_perfect():
    rng = random.Random(5)
    for n in (2, 7, 20, 33):
        m = qscript.random_matching(rng, n)
        flat = [x for p in m for x in p]
        assert len(flat) == len(set(flat)) == n - (n % 2)


def mk(n, prec=32, **kw):
    return QEngineRestate(n, 0, random.Random(1), 1.0 + 0j, False, False, precision=prec, **kw)


def test_bounds_errors_like_reference():
    q = mk(3)
    with pytest.raises(ValueError):
        q.H(3)
    with pytest.raises(ValueError):
        q.Apply2x2(0, 8, [1, 0, 0, 1], 1, [8], False)
    with pytest.raises(ValueError):
        q.Apply2x2(0, 2, [1, 0, 0, 1], 2, [2, 2], False)
    with pytest.raises(ValueError):
        q.Prob(5)
    with pytest.raises(ValueError):
        q.ProbMask(8, 0)
    with pytest.raises(ValueError):
        q.ForceM(0, True)            # zero-probability result


def test_identity_elision_and_global_phase():
    q = mk(2)
    before = q.GetQuantumState()
    q.Phase(1.0, 1.0, 0)            # exact identity: elided
    q.Mtrx([1, 0, 0, 1], 1)
    assert np.array_equal(before, q.GetQuantumState())
    q.Phase(1j, 1j, 0)              # global phase with randGlobalPhase=False must be applied
    assert abs(q.GetAmplitude(0) - 1j) < 1e-7


def test_zero_state_semantics():
    q = mk(4)
    q.ZeroAmplitudes()
    assert q.IsZeroAmplitude()
    q.H(1)
    assert q.Prob(1) == 0.0
    assert not q.GetQuantumState().any()
    q.SetAmplitudePage(np.array([0.6, 0.8j], dtype=np.complex64), 2)
    assert not q.IsZeroAmplitude()
    assert abs(q.ProbAll(3) - 0.64) < 1e-6


def test_do_normalize_running_norm():
    q = QEngineRestate(3, 0, random.Random(1), 1.0 + 0j, True, False)
    q.SetQuantumState(np.array([2, 0, 0, 0, 0, 0, 0, 0], dtype=np.complex64))
    assert abs(q.Prob(0)) < 1e-7
    assert abs(q.ProbAll(0) - 1.0) < 1e-6          # normalised on read
    q.H(0)
    st = q.GetQuantumState()
    assert abs(np.linalg.norm(st) - 1) < 1e-6


def test_mirror_circuit_returns():
    text = qscript.random_htcnot(8, 6, seed=2, timed=False)
    ops = [t for _, t in qscript.parse(text)][1:]
    q = mk(8)
    q.SetPermutation(37, 1.0)
    inv = {"H": "H", "T": "IT", "CNOT": "CNOT"}
    for t in ops:
        getattr(q, t[0])(*[int(x) for x in t[1:]])
    for t in reversed(ops):
        getattr(q, inv[t[0]])(*[int(x) for x in t[1:]])
    assert abs(q.ProbAll(37) - 1.0) < 1e-5
