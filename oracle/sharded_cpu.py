"""oracle/sharded_cpu.py — TEST INFRASTRUCTURE: local-engine factory that runs qrack_b200.sharded.QEngineSharded on the
oracle restatement over torch CPU buffers (gloo), so the sharding logic can be checked without a GPU."""
import random

import numpy as np

from oracle.restate_engine import QEngineRestate


def restate_engine_factory(precision: int = 32):
    cplx = np.complex64 if precision == 32 else np.complex128

    def make(buf, n_local):
        q = QEngineRestate(n_local, 0, random.Random(1), 1.0 + 0j, False, False, precision=precision)
        q.be.amps = buf.numpy().view(cplx)  # shares memory with the torch page
        return q
    return make
