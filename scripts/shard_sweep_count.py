"""Host-only model of a sharded step: replays a circuit through QEngineSharded's dispatch with a shard that moves no data, captures the
gate batch of every window (what `_submit_batch` hands to b200sv_apply_gates between two exchanges) and asks the REAL fused planner
(b200sv_plan_gates, host only) how many sweeps / passes / device ops each window costs.  No GPU, no torch.distributed."""
import ctypes
import os
import random
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qrack_b200 import _abi, qscript, sharded  # noqa: E402


class _Be:
    def __init__(self, owner):
        self.owner = owner

    def apply_gates(self, n, o1, o2, pm, m8):
        self.owner.windows.append((n, o1, o2, pm, m8))

    def flush(self):
        pass


class _NullEngine:
    runningNorm = 1.0

    def __init__(self):
        self.windows = []
        self.be = _Be(self)

    def Finish(self):
        pass

    def SetAmplitude(self, *a):
        pass


class _NullShard:
    min_victim_bit = 8
    needs_top = False

    def __init__(self):
        self.engine = _NullEngine()
        self.cuts = []

    def zero_live(self):
        pass

    def exchange(self, dist, world, rank, k, vbits):
        self.cuts.append(len(self.engine.windows))
        return 0


def windows(n_local, world, rank, text):
    class Eng(sharded.QEngineSharded):
        def _make_backend(self, n_qubits):
            return sharded._ShardedBackend(n_qubits, 32, _NullShard(), None, world, rank)

    def make(n, perm):
        return Eng(n, perm, random.Random(1), 1.0 + 0j, precision=32, world=world, rank=rank)
    regs, _ = qscript.run(text, make)
    q = regs[0]
    q.be.flush()
    return q.be.exchanges, q.be.shard.engine.windows


def plan(lib, n_local, w):
    n, o1, o2, pm, m8 = w
    sw, ps, ops = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    _abi.check(lib, lib.b200sv_plan_gates(n_local, 32, n, o1, o2, pm, m8, ctypes.byref(sw), ctypes.byref(ps), ctypes.byref(ops)))
    return n, sw.value, ps.value, ops.value


def report(name, n_local, world, text, ranks=(0,), verbose=True):
    lib = _abi.load()
    out = []
    for rank in ranks:
        ex, wins = windows(n_local, world, rank, text)
        rows = [plan(lib, n_local, w) for w in wins]
        tot = [sum(r[i] for r in rows) for i in range(4)]
        out.append((ex, tot))
        if verbose:
            print("%s world %d rank %d: %d exchanges, %d windows, %d local gates -> %d sweeps, %d passes, %d device ops" % (
                name, world, rank, ex, len(rows), tot[0], tot[1], tot[2], tot[3]))
            print("   per window (gates/sweeps/passes): " + " ".join("%d/%d/%d" % r[:3] for r in rows))
    return out


if __name__ == "__main__":
    worlds = [int(a) for a in sys.argv[1:]] or [1, 2, 8]
    for world in worlds:
        k = world.bit_length() - 1
        n = 30 + k
        report("htcnot", 30, world, qscript.random_htcnot(n, 40, seed=20250921, timed=False), ranks=(0, world - 1) if world > 1 else (0,))
        report("qv", 30, world, qscript.quantum_volume(n, depth=n, seed=33, timed=False))
