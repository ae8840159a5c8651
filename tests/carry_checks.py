"""Shared body of the b200sv_set_rank_bits / b200sv_flush_carry checks: runs against any engine whose backend offers
set_state / get_state / apply_gates / flush / flush_carry / set_rank_bits — the CUDA backend (tests/test_zz_carry_gpu.py) and the
host-interpreter backend (tests/test_fused_emulation.py), so that the test logic itself is exercised without a device."""
import cmath
import ctypes
import random

import numpy as np

from qrack_b200 import _abi

import util
from test_fused_emulation import _random_gate_arrays


def _gate_list(n, k, count, rng):
    """random single-target gates on n real qubits; some get a control on a virtual qubit (n .. n+k-1), some are diagonal gates
    whose own qubit is virtual"""
    g, o1, o2, pm, m8 = _random_gate_arrays(n, count, rng)
    gates = []
    for i in range(g):
        a, b, p, m = o1[i], o2[i], pm[i], [complex(m8[8 * i + 2 * j], m8[8 * i + 2 * j + 1]) for j in range(4)]
        if k and rng.random() < 0.25:
            v = n + rng.randrange(k)
            p |= 1 << v
            if rng.random() < 0.5:
                a |= 1 << v
                b |= 1 << v
        gates.append((a, b, p, m))
        if k and rng.random() < 0.08:
            v = n + rng.randrange(k)
            c = rng.randrange(n)
            ph = cmath.exp(1j * rng.uniform(0, 6.2))
            form = rng.randrange(3)
            if form == 0:      # T-like gate on the virtual qubit
                gates.append((0, 1 << v, 1 << v, [1 + 0j, 0j, 0j, ph]))
            elif form == 1:    # controlled phase: real control, virtual "target"
                gates.append((1 << c, (1 << c) | (1 << v), (1 << c) | (1 << v), [1 + 0j, 0j, 0j, ph]))
            else:              # diag(ph, 1) on the virtual qubit, anti-controlled by a real one
                gates.append((0, 1 << v, (1 << c) | (1 << v), [ph, 0j, 0j, 1 + 0j]))
    return gates


def _pack(gates):
    n = len(gates)
    o1 = (ctypes.c_uint64 * n)(*[g[0] for g in gates])
    o2 = (ctypes.c_uint64 * n)(*[g[1] for g in gates])
    pm = (ctypes.c_uint64 * n)(*[g[2] for g in gates])
    m8 = (ctypes.c_double * (8 * n))()
    for i, g in enumerate(gates):
        for j in range(4):
            m8[8 * i + 2 * j] = complex(g[3][j]).real
            m8[8 * i + 2 * j + 1] = complex(g[3][j]).imag
    return n, o1, o2, pm, m8


def _specialise(gates, n, k, rank):
    """what the gates mean on the page of rank `rank`: predicates on virtual qubits evaluated, diagonal gates on a virtual qubit
    reduced to their entry (the reference form the engine has to agree with)"""
    vmask = ((1 << k) - 1) << n
    rv = rank << n
    out = []
    for (a, b, p, m) in gates:
        diff = a ^ b
        cm = p & ~diff
        cv = a & ~diff
        if (cv & vmask & cm) != (rv & vmask & cm):
            continue
        cm &= ~vmask
        cv &= ~vmask
        if diff & vmask:   # diagonal gate on a virtual qubit: one entry, as a (controlled) scalar
            d = m[3] if (rv & diff) else m[0]
            if cm:
                c = cm.bit_length() - 1
                want = (cv >> c) & 1
                cm &= ~(1 << c)
                cv &= ~(1 << c)
                out.append((cv, cv | (1 << c), cm | (1 << c), [1 + 0j, 0j, 0j, d] if want else [d, 0j, 0j, 1 + 0j]))
            else:
                out.append((0, 1, 1, [d, 0j, 0j, d]))
            continue
        out.append((cv, cv | diff, cm | diff, m))
    return out


def check_rank_bits_and_carry(make_engine, prec, n=16, k=2, count=260, seed=0):
    """make_engine(n) -> engine with .be; returns the number of ops that were handed back (summed over the ranks)"""
    lib = _abi.load()
    cplx = np.complex64 if prec == 32 else np.complex128
    rng = random.Random(1000 * seed + n + k)
    nrng = np.random.default_rng(seed)
    st = (nrng.standard_normal(1 << n) + 1j * nrng.standard_normal(1 << n))
    st = (st / np.linalg.norm(st)).astype(cplx)
    gates = _gate_list(n, k, count, rng)
    must = 0
    for b in rng.sample(range(1, n), 2):
        must |= 1 << b
    handed = 0
    for rank in range(1 << k):
        # the reference: the same gates specialised for this rank, through the plain host interpreter of the fused programs
        want = st.copy()
        g2, a1, a2, ap, am = _pack(_specialise(gates, n, k, rank))
        _abi.check(lib, lib.b200sv_emulate_fused(n, prec, g2, a1, a2, ap, am, want.ctypes.data_as(ctypes.c_void_p)))
        # A: one flush
        qa = make_engine(n)
        qa.be.set_state(st)
        qa.be.set_rank_bits(k, rank)
        qa.be.apply_gates(*_pack(gates))
        qa.be.flush()
        got_a = qa.be.get_state()
        # B: the tail is handed back (no non-diagonal op on the `must` qubits among it), then submitted again
        qb = make_engine(n)
        qb.be.set_state(st)
        qb.be.set_rank_bits(k, rank)
        qb.be.apply_gates(*_pack(gates))
        back = qb.be.flush_carry(1000000, must)
        for (a, b, p, m) in back:
            t = (a ^ b).bit_length() - 1
            assert not (((must >> t) & 1) and not (m[1] == 0 and m[2] == 0)), "a non-diagonal op on a must-qubit was handed back"
            assert t < n or (m[1] == 0 and m[2] == 0)
        handed += len(back)
        if back:
            qb.be.apply_gates(*_pack(back))
        got_b = qb.be.get_state()
        for name, got in (("one flush", got_a), ("carry + resubmit", got_b)):
            d = float(np.abs(got.astype(np.complex128) - want.astype(np.complex128)).max())
            assert d <= 2 * util.AMP_TOL[prec], "rank %d, %s: max |delta amp| = %.3e" % (rank, name, d)
        del qa, qb
    return handed
