"""Host-only dry run of the sharded engine's exchange planner: replays a circuit through QEngineSharded's dispatch with
a shard that moves no data and counts exchanges and local gate batches (no GPU, no torch.distributed)."""
import os
import random
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qrack_b200 import qscript, sharded  # noqa: E402


class _NullEngine:
    runningNorm = 1.0

    def __init__(self):
        self.gates = 0
        self.batches = 0
        self._open = False

    def UCMtrx(self, *a):
        self.gates += 1
        if not self._open:
            self.batches += 1
            self._open = True

    Mtrx = lambda self, *a: self.UCMtrx()
    Swap = lambda self, *a: self.UCMtrx()

    def Finish(self):
        self._open = False

    def SetAmplitude(self, *a):
        pass


class _NullShard:
    def __init__(self, p2p=True):
        self.engine = _NullEngine()
        self.min_victim_bit = 8 if p2p else 0
        self.needs_top = not p2p

    def zero_live(self):
        pass

    def exchange(self, dist, world, rank, k, vbits):
        self.engine.Finish()     # an exchange cuts the local fused window
        return 0


def count(n_local, world, text, defer, p2p=True):
    k = world.bit_length() - 1
    os.environ["B200SV_SHARD_DEFER"] = "1" if defer else "0"

    class Eng(sharded.QEngineSharded):
        def _make_backend(self, n_qubits):
            return sharded._ShardedBackend(n_qubits, 32, _NullShard(p2p), None, world, 0)

    def make(n, perm):
        return Eng(n, perm, random.Random(1), 1.0 + 0j, precision=32, world=world, rank=0)
    regs, _ = qscript.run(text, make)
    q = regs[0]
    q.be.flush()
    return q.be.exchanges, q.be.shard.engine.batches, q.be.shard.engine.gates


if __name__ == "__main__":
    for world in (2, 4, 8):
        k = world.bit_length() - 1
        n = 30 + k
        for name, text in (("htcnot", qscript.random_htcnot(n, 40, seed=20250921, timed=False)),
                           ("qv", qscript.quantum_volume(n, depth=40, seed=33, timed=False)),
                           ("qft", qscript.qft(n, seed=11, timed=False))):
            a = count(30, world, text, False)
            b = count(30, world, text, True)
            print("world %d %-7s in-order: %3d exchanges (%d local batches)   deferred: %3d exchanges (%d local batches)" % (
                world, name, a[0], a[1], b[0], b[1]))
