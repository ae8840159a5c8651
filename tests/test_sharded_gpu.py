"""Multi-GPU parity (needs >= 2 CUDA devices; skipped otherwise): QEngineSharded over QEngineCUDA local engines + NCCL
all-to-all must reproduce the single-engine oracle state."""
import os
import random

import numpy as np
import pytest

from oracle.restate_engine import QEngineRestate
from qrack_b200 import qscript

import util
from test_sharded_cpu import _free_port

pytestmark = pytest.mark.gpu


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _worker(rank, world, port, text, prec, out_path, p2p):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from qrack_b200.sharded import QEngineSharded, cuda_engine_factory

        def make(n, perm):
            return QEngineSharded(n, perm, random.Random(1), 1.0 + 0j, precision=prec, dist=dist, world=world, rank=rank,
                                  device=torch.device("cuda", rank), make_engine=cuda_engine_factory(rank, prec), p2p=p2p)
        regs, results = qscript.run(text, make)
        st = regs[0].GetQuantumState()
        if rank == 0:
            np.savez(out_path, state=st, exchanges=regs[0].be.exchanges)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("p2p", [False, True], ids=["nccl_all_to_all", "p2p_scatter_kernel"])
@pytest.mark.parametrize("prec", [32, 64])
def test_sharded_nccl_matches_oracle(prec, p2p, tmp_path):
    world = 2 if _ngpu() < 4 else 4
    if _ngpu() < 2:
        pytest.skip("needs >= 2 GPUs")
    import torch.multiprocessing as mp
    text = qscript.random_htcnot(18, 12, seed=8, timed=False)
    want, _ = util.run_engine(text, QEngineRestate, prec)
    out = str(tmp_path / "o.npz")
    for attempt in range(3):  # the rendezvous port can be taken between probing and binding
        try:
            mp.spawn(_worker, args=(world, _free_port(), text, prec, out, p2p), nprocs=world, join=True)
            break
        except Exception as e:
            if "EADDRINUSE" not in str(e) or attempt == 2:
                raise
    z = np.load(out)
    d = float(np.abs(z["state"].astype(np.complex128) - want[0].astype(np.complex128)).max())
    assert d <= util.AMP_TOL[prec], d
    assert int(z["exchanges"]) >= 1
