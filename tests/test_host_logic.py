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


def _basis(q):
    st = q.GetQuantumState()
    nz = np.nonzero(np.abs(st) > 1e-9)[0]
    assert len(nz) == 1
    return int(nz[0]), complex(st[nz[0]])


def test_qalu_mirror_semantics_on_basis_states():
    """QAlu members through the host mirror (src/qalu.cpp wrappers + src/qengine/arithmetic.cpp loops) against the
    arithmetic they stand for, on computational basis states (an independent statement of the semantics)."""
    rng = random.Random(11)
    for _ in range(40):
        n, L = 12, 4
        x, a = rng.randrange(1 << L), rng.randrange(1, 1 << L)
        hi = rng.randrange(4) << 10                      # spectator qubits 10, 11
        q = QEngineRestate(n, x | hi, random.Random(1), 1.0 + 0j, False, False, precision=64)
        q.INC(a, 0, L)
        assert _basis(q)[0] == ((x + a) % 16) | hi
        q.DEC(a, 0, L)
        assert _basis(q)[0] == x | hi
        q.INCC(a, 0, L, 4)                               # carry qubit 4 starts at 0
        assert _basis(q)[0] == ((x + a) % 16) | (((x + a) >> 4) << 4) | hi
        q.SetPermutation(x | hi)
        q.CINC(a, 0, L, [10])
        assert _basis(q)[0] == ((((x + a) % 16) if hi & (1 << 10) else x) | hi)
        q.SetPermutation(x | hi)
        q.MUL(a, 0, 4, L)                                # product spread over in/out 0..3 and carry 4..7
        assert _basis(q)[0] == ((x * a) & 15) | ((((x * a) >> 4) & 15) << 4) | hi
        q.DIV(a, 0, 4, L)
        assert _basis(q)[0] == x | hi
        q.MULModNOut(a, 13, 0, 4, L)
        assert _basis(q)[0] == x | (((x * a) % 13) << 4) | hi
        q.IMULModNOut(a, 13, 0, 4, L)
        assert _basis(q)[0] == x | hi
        q.POWModNOut(3, 11, 0, 4, L)
        assert _basis(q)[0] == x | ((pow(3, x) % 11) << 4) | hi
        q.SetPermutation(x | hi)
        q.ROL(1, 0, L)
        assert _basis(q)[0] == (((x << 1) | (x >> 3)) & 15) | hi
        q.ROR(1, 0, L)
        assert _basis(q)[0] == x | hi
        tab = bytes(rng.randrange(16) for _ in range(16))
        q.IndexedLDA(0, L, 4, 4, tab)
        assert _basis(q)[0] == x | (tab[x] << 4) | hi
        q.IndexedADC(0, L, 4, 4, 8, tab)                 # value += tab[x], carry out into qubit 8
        v = 2 * tab[x]
        assert _basis(q)[0] == x | ((v & 15) << 4) | ((v >> 4) << 8) | hi
        perm = list(range(16))
        rng.shuffle(perm)
        q.SetPermutation(x | hi)
        q.Hash(0, L, bytes(perm))
        assert _basis(q)[0] == perm[x] | hi
        q.SetPermutation(x | hi)
        q.PhaseFlipIfLess(a, 0, L)
        assert abs(_basis(q)[1] - (-1.0 if x < a else 1.0)) < 1e-12
        q.INCS(a, 0, L, 10)                              # sign flip only if signed overflow AND flag qubit 10 set
        sx, sa = (x - 16 if x & 8 else x), (a - 16 if a & 8 else a)
        ovf = not (-8 <= sx + sa <= 7)
        want = (-1.0 if x < a else 1.0) * (-1.0 if (ovf and (hi & (1 << 10))) else 1.0)
        assert abs(_basis(q)[1] - want) < 1e-12, (x, a, hi)


def test_qalu_argument_errors_and_unsupported_backends():
    q = QEngineRestate(8, 0, random.Random(1), 1.0 + 0j, False, False, precision=32)
    with pytest.raises(ValueError):
        q.INC(1, 6, 4)
    with pytest.raises(ValueError):
        q.INCC(1, 0, 4, 9)
    with pytest.raises(ValueError):
        q.CINC(1, 0, 4, [8])
    with pytest.raises(ValueError):
        q.DIV(0, 0, 4, 4)
    with pytest.raises(ValueError):
        q.Hash(0, 4, b"\x00\x01")                        # table too short
    q.ZeroAmplitudes()
    q.INC(3, 0, 4)                                       # CHECK_ZERO_SKIP
    assert q.IsZeroAmplitude()


def test_prob_bits_all_order_and_multishot_histogram_path():
    """QInterface::ProbBitsAll (qinterface.cpp:446-476) indexes its output by the ORDER of the listed qubits; the small-mask path
    of MultiShotMeasureMask (qengine.cpp:542-576) draws from that histogram."""
    import numpy as np
    from oracle.restate_engine import QEngineRestate
    n = 7
    q = QEngineRestate(n, 0, random.Random(3), 1.0 + 0j, False, False, precision=64)
    for b in range(n):
        q.U(b, 0.3 + 0.4 * b, 0.1 * b, 0.2)
    q.CNOT(0, 5)
    q.CNOT(6, 2)
    st = q.GetQuantumState()
    pr = (st.real ** 2 + st.imag ** 2)
    bits = [5, 0, 6]
    want = np.zeros(8)
    for i in range(1 << n):
        k = sum(((i >> b) & 1) << p for p, b in enumerate(bits))
        want[k] += pr[i]
    got = q.ProbBitsAll(bits)
    assert np.abs(np.asarray(got, dtype=np.float64) - want).max() < 1e-12
    res = q.MultiShotMeasureMask([1 << b for b in bits], 4000)
    assert sum(res.values()) == 4000 and set(res) <= set(range(8))
    emp = np.array([res.get(k, 0) for k in range(8)]) / 4000.0
    assert np.abs(emp - want).max() < 0.04          # ~5 sigma of a 4000-shot binomial
    with pytest.raises(ValueError):
        q.MultiShotMeasureMask([3], 10)


def test_qcircuit_records_the_dispatch_and_replays_it():
    """QCircuit records every gate method through the same dispatch mirror the engines use and replays it on any
    engine: gate by gate here (what QCircuit::Run does in the reference), in one ABI call on QEngineCUDA."""
    import numpy as np
    from oracle.restate_engine import QEngineRestate
    from qrack_b200 import QCircuit
    n = 9
    text = qscript.random_htcnot(n, 5, seed=3, timed=False) + "QFT 1 7\nCCNOT 0 8 4\nAntiCNOT 2 6\nINC 5 2 6\nU 3 0.3 0.2 0.1\nXMask 24\nCZ 1 7\n"
    c = QCircuit(n, 32)
    for _, t in qscript.parse(text):
        if t[0] == "qubits":
            continue
        getattr(c, t[0])(*[(float(x) if "." in x else int(x)) for x in t[1:]])
    assert c.GetGateCount() > 100
    q = QEngineRestate(n, 0, random.Random(1), 1.0 + 0j, False, False, precision=32)
    c.Run(q)
    ref, _ = qscript.run(text, lambda nq, p: QEngineRestate(nq, p, random.Random(1), 1.0 + 0j, False, False, precision=32))
    assert np.abs(q.GetQuantumState() - ref[0].GetQuantumState()).max() < 1e-6
    with pytest.raises(NotImplementedError):
        c.Prob(0)                      # state access is an engine call, not a circuit element
    with pytest.raises(ValueError):
        c.Run(QEngineRestate(n + 1, 0, random.Random(1), 1.0 + 0j, False, False, precision=32))


def test_cuda_backend_marshalling_with_a_fake_library():
    """_CudaBackend.apply2x2 (every gate of the GPU tests and of bench.py goes through it) against a stand-in library: argument
    order, matrix / power arrays, the cached arrays on repeated gates, the norm read-back.  No device, no libb200sv call."""
    from qrack_b200.qengine import _CudaBackend
    be = object.__new__(_CudaBackend)
    calls = []

    class Lib:
        def b200sv_apply2x2(self, h, o1, o2, m8, n, pw, nrm, th, out):
            calls.append((h, o1, o2, list(m8), n, list(pw)[:n], nrm, th))
            if out is not None:
                out._obj.value = 0.75
            return 0
    be.lib, be.h = Lib(), 123
    be._ck = lambda rc: None
    s = 2 ** -0.5
    had = [complex(s), complex(s), complex(s), complex(-s)]
    for _ in range(2):
        assert be.apply2x2(0, 4, had, [4], 1.0, 0.0, False) is None
        assert be.apply2x2(1, 5, [0j, 1 + 0j, 1 + 0j, 0j], [1, 4], 1.0, 0.0, False) is None
    assert abs(be.apply2x2(0, 2, [1 + 0j, 0j, 0j, 1j], [2], 0.5, 1e-9, True) - 0.75) < 1e-12
    assert calls[0] == (123, 0, 4, [s, 0, s, 0, s, 0, -s, 0], 1, [4], 1.0, 0.0)
    assert calls[1] == (123, 1, 5, [0, 0, 1, 0, 1, 0, 0, 0], 2, [1, 4], 1.0, 0.0)
    assert calls[2] == calls[0] and calls[3] == calls[1]
    assert calls[4][3] == [1, 0, 0, 0, 0, 0, 0, 1] and calls[4][6] == 0.5 and calls[4][7] == 1e-9


def test_matrix_memo_keeps_the_dispatch_semantics():
    """the memoised rounding / classification of gate matrices (QEngineHost._mtrx, IsIdentity, IsPhase, IsInvert) returns what the
    direct computation returns, for both precisions, and hands out fresh lists"""
    import cmath
    for prec in (32, 64):
        q = QEngineRestate(3, 0, random.Random(1), 1.0 + 0j, False, False, precision=prec)
        mats = [[1, 0, 0, 1], [1, 0, 0, -1], [0, 1, 1, 0], [2 ** -0.5] * 3 + [-(2 ** -0.5)], [1, 0, 0, cmath.exp(0.25j * cmath.pi)],
                [0.6, 0.8j, 0.8j, 0.6], [1 + 1e-9, 0, 0, 1], [1, 1e-12, 0, 1]]
        for m in mats * 2:
            got = q._mtrx(m)
            assert got == [complex(q.cplx(x)) for x in m]
            got[0] = 99  # a caller may scribble on its copy
            assert q._mtrx(m)[0] == complex(q.cplx(m[0]))
            r = q._mtrx(m)
            assert q.IsPhase(r) == (q._norm(r[1]) <= q.FP_NORM_EPSILON and q._norm(r[2]) <= q.FP_NORM_EPSILON)
            assert q.IsInvert(r) == (q._norm(r[0]) <= q.FP_NORM_EPSILON and q._norm(r[3]) <= q.FP_NORM_EPSILON)
            for ctl in (False, True):
                assert q.IsIdentity(r, ctl) == q._is_identity(r, ctl)
