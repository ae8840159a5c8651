"""CPU tests of the fused-sweep scheduler and program encoder: the real planner/encoder of libb200sv.so + the host
interpreter of the encoded programs (b200sv_emulate_fused, no device) against the oracle restatement and the golden
fixtures of the compiled reference.  Same tolerances as the GPU parity tests."""
import os
import random

import numpy as np
import pytest

from oracle.restate_engine import QEngineRestate
from qrack_b200 import qscript

import util
from emu_engine import QEngineEmu


def run_emu(text, prec):
    regs, results = qscript.run(text, util.make_factory(QEngineEmu, prec))
    return {k: v.GetQuantumState() for k, v in regs.items()}, results, regs


@pytest.mark.parametrize("prec", [32, 64])
@pytest.mark.parametrize("name", [n for n in util.golden_names() if not n.startswith("alu_")])
def test_emulated_sweeps_reproduce_golden_fixtures(name, prec):
    text, regs, results = util.load_golden(name, prec)
    got, gres, _ = run_emu(text, prec)
    util.assert_states_close(got, regs, prec, name)
    util.assert_results_close(gres, results, prec, name)


@pytest.mark.parametrize("prec", [32, 64])
@pytest.mark.parametrize("gen", ["htcnot", "u3", "qft", "qv", "grover"])
def test_emulated_sweeps_vs_oracle_multi_tile(gen, prec):
    """16-17 qubits: 8-16 tiles per sweep, high tile qubits, outer controls (ballots), DIAG slots, several passes."""
    n = 16 if prec == 32 else 15
    text = {"htcnot": lambda: qscript.random_htcnot(n, 12, seed=5, timed=False),
            "u3": lambda: qscript.random_u3_cnot(n, 5, seed=6),
            "qft": lambda: "qubits %d\nSetPermutation 12345\nH 3\nH 9\nQFT 0 %d\nT 2\nIQFT 1 %d\n" % (n, n, n - 2),
            "qv": lambda: qscript.quantum_volume(n, depth=5, seed=8, timed=False),
            "grover": lambda: qscript.grover(n, 2, target=77, timed=False)}[gen]()
    want, _ = util.run_engine(text, QEngineRestate, prec)
    got, _, regs = run_emu(text, prec)
    util.assert_states_close(got, want, prec, gen)
    assert regs[0].be.flushes >= 1          # the gates really went through the planner + emulator


@pytest.mark.parametrize("knobs,search", [("3,5,4,3,3", "0"), ("4,6,6,7,3", "2"), ("4,7,7,3,4", "1"), ("4,6,6,0,3", "0"),
                                          ("3,9,8,1,3", "4"), ("4,6,6,3,3,0,0,0", "2"), ("4,6,6,3,3,0,0,3", "0")])
def test_emulated_sweeps_under_every_tile_shape(knobs, search):
    """The tile-shape / bundling knobs (B200SV_FUSED) and the tile-qubit search (B200SV_PLAN_SEARCH) change the choice of
    high qubits, pass tables and DIAG/LAYER grouping; each setting must
    still reproduce the oracle.  Runs in a subprocess because the library reads the knobs once."""
    import subprocess
    import sys
    code = (
        "import sys, random; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "import numpy as np\n"
        "from oracle.restate_engine import QEngineRestate\n"
        "from qrack_b200 import qscript\n"
        "import util\n"
        "from emu_engine import QEngineEmu\n"
        "for prec in (32, 64):\n"
        "    text = qscript.random_htcnot(15, 8, seed=3, timed=False) + 'QFT 2 9\\nCCNOT 1 14 7\\nMCPhase 2 3 13 8 0.6 0.8 1 0\\n'\n"
        "    want, _ = util.run_engine(text, QEngineRestate, prec)\n"
        "    regs, _ = qscript.run(text, util.make_factory(QEngineEmu, prec))\n"
        "    util.assert_states_close({k: v.GetQuantumState() for k, v in regs.items()}, want, prec, 'knobs')\n"
        "print('ok')\n"
    ) % (util.ROOT, os.path.join(util.ROOT, "tests"))
    # half of the settings also switch the rotation stages (rewrite R5) and the lazy diagonals off
    env = dict(os.environ, B200SV_FUSED=knobs, B200SV_PLAN_SEARCH=search, B200SV_ROT=("0" if search in ("0", "4") else "1"),
               B200SV_LAZY_DIAG=("0" if search == "1" else "1"))
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr[-2000:]


def test_emulation_hook_rejects_bad_arguments():
    import ctypes
    from qrack_b200 import _abi
    lib = _abi.load()
    st = np.zeros(32, dtype=np.complex64)
    one = (ctypes.c_uint64 * 1)(3)
    two = (ctypes.c_uint64 * 1)(0)
    pm = (ctypes.c_uint64 * 1)(3)
    m = (ctypes.c_double * 8)(1, 0, 0, 0, 0, 0, 1, 0)
    # off1 ^ off2 has two bits: not a single-target gate
    assert lib.b200sv_emulate_fused(5, 32, 1, one, two, pm, m, st.ctypes.data_as(ctypes.c_void_p)) == _abi.B200SV_EINVAL
    assert lib.b200sv_emulate_fused(3, 32, 0, None, None, None, None, st.ctypes.data_as(ctypes.c_void_p)) == _abi.B200SV_EINVAL


def _random_mixed_circuit(rng, n, ngates):
    import math
    L = ["qubits %d" % n] + ["H %d" % q for q in range(n) if rng.random() < 0.7]

    def cx():
        a = rng.uniform(0, 2 * math.pi)
        return "%.17g %.17g" % (math.cos(a), math.sin(a))

    def unitary():
        th, ph, la = (rng.uniform(-math.pi, math.pi) for _ in range(3))
        c, s = math.cos(th / 2), math.sin(th / 2)
        m = [c, -s * complex(math.cos(la), math.sin(la)), s * complex(math.cos(ph), math.sin(ph)),
             c * complex(math.cos(ph + la), math.sin(ph + la))]
        return " ".join("%.17g %.17g" % (complex(z).real, complex(z).imag) for z in m)

    for _ in range(ngates):
        r, qs = rng.random(), rng.sample(range(n), 4)
        if r < 0.2:
            L.append("%s %d" % (rng.choice(["H", "T", "X", "S", "Z", "Y", "IT", "SqrtX"]), qs[0]))
        elif r < 0.4:
            L.append("%s %d %d" % (rng.choice(["CNOT", "CZ", "AntiCNOT", "CY"]), qs[0], qs[1]))
        elif r < 0.5:
            L.append("CCNOT %d %d %d" % tuple(qs[:3]))
        elif r < 0.6:
            L.append("U %d %.17g %.17g %.17g" % (qs[0], rng.uniform(-3, 3), rng.uniform(-3, 3), rng.uniform(-3, 3)))
        elif r < 0.7:
            L.append("MCMtrx 2 %d %d %d %s" % (qs[0], qs[1], qs[2], unitary()))
        elif r < 0.78:
            L.append("MACMtrx 2 %d %d %d %s" % (qs[0], qs[1], qs[2], unitary()))
        elif r < 0.86:
            L.append("MCPhase 2 %d %d %d %s %s" % (qs[0], qs[1], qs[2], cx(), cx()))
        elif r < 0.92:
            L.append("MACInvert 1 %d %d %s %s" % (qs[0], qs[1], cx(), cx()))
        elif r < 0.96:
            L.append("CPhaseRootN %d %d %d" % (rng.randrange(1, 6), qs[0], qs[1]))
        else:
            L.append("UCMtrx 3 %d %d %d %d %d %s" % (qs[0], qs[1], qs[2], qs[3], rng.randrange(8), unitary()))
    return "\n".join(L) + "\n"


def test_emulated_sweeps_fuzz_mixed_gates():
    """Random circuits mixing every single-target form the dispatch produces (controls, anti-controls, phases, inverts,
    uniformly-controlled selections) on 13-17 qubits, both precisions, through planner + encoder + interpreter."""
    rng = random.Random(2025)
    for trial in range(10):
        prec = rng.choice([32, 64])
        n = rng.randrange(13, 18) if prec == 32 else rng.randrange(13, 17)
        text = _random_mixed_circuit(rng, n, rng.randrange(60, 300))
        want, _ = util.run_engine(text, QEngineRestate, prec)
        got, _, _ = run_emu(text, prec)
        util.assert_states_close(got, want, prec, "fuzz %d" % trial)
