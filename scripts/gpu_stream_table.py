"""Roofline table of the streaming kernels behind SURVEY §8 rows a3-a8/a10 at BASELINE width (run under gpurun, 1 GPU):

    ApplyM/ForceM, Prob(q), ProbMask, ProbParity, ProbAll marginals, NormalizeState (norm + scale), UpdateRunningNorm,
    XMask (dedicated sweep), PhaseParity, PhaseRootNMask, Compose (n-1)+1, Decompose n -> (n-1)+1, Dispose(perm),
    ShuffleBuffers (two pages on one GPU), CopyStateVec, unfused Apply2x2 (k_apply2x2), MAll sampling.

ms = CUDA events on the engine's stream (b200sv_timer_*), warm-up + repetitions, state (8 GiB) >> L2.  bytes = the
ALGORITHMIC bytes of the op (what has to be read and written at minimum), so frac = algorithmic GB/s over the
measured copy peak (MEASURED_PEAKS.json); a kernel that re-reads the state shows up as a low fraction.
"""
import json
import os
import random
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qrack_b200 import QEngineCUDA  # noqa: E402

n = int(os.environ.get("N", "30"))
peak = json.load(open("MEASURED_PEAKS.json")).get("hbm_gbs", 6486.8) if os.path.exists("MEASURED_PEAKS.json") else 6486.8
rows = []


def timed(q, fn, reps=5, warm=2, setup=None):
    tot = 0.0
    for r in range(warm + reps):
        if setup:
            setup()
        q.Finish()
        q.be.timer_begin()
        fn()
        ms = q.be.timer_end()
        if r >= warm:
            tot += ms
    return tot / reps


def row(name, ms, nbytes, note=""):
    gbs = nbytes / (ms * 1e-3) / 1e9
    rows.append({"op": name, "ms": round(ms, 4), "algorithmic_bytes": nbytes, "gbs": round(gbs, 1), "frac_of_peak": round(gbs / peak, 3), "note": note})
    print("%-34s %9.3f ms  %8.1f GB/s  %.3f  %s" % (name, ms, gbs, gbs / peak, note), flush=True)


for prec in [int(p) for p in os.environ.get("PRECS", "32").split(",")]:
    S = 8 if prec == 32 else 16
    nq = n if prec == 32 else n - 1
    N = 1 << nq
    tag = "fp%d %dq " % (prec, nq)
    q = QEngineCUDA(nq, 0, random.Random(1), 1.0 + 0j, False, False, precision=prec)
    for b in range(nq):
        q.H(b)
    q.T(3)
    q.Finish()

    q.be.set_fusion(0)
    row(tag + "Apply2x2 H(20) unfused", timed(q, lambda: (q.H(20), q.be.flush())), 2 * N * S, "k_apply2x2")
    row(tag + "Apply2x2 CNOT(3,20) unfused", timed(q, lambda: (q.CNOT(3, 20), q.be.flush())), N * S, "half the amplitudes move")
    q.be.set_fusion(1)
    row(tag + "fused sweep, 2 H", timed(q, lambda: (q.H(0), q.H(1), q.be.flush())), 2 * N * S, "k_fused_sweep, one pass")
    row(tag + "XMask 5 qubits (dedicated sweep)", timed(q, lambda: (q.be.set_fusion(0), q.XMask(0b1000100010001000100000), q.be.set_fusion(1))), 2 * N * S)
    row(tag + "PhaseParity 4 qubits", timed(q, lambda: q.PhaseParity(0.7, 0b10001000100010000)), 2 * N * S)
    row(tag + "PhaseRootNMask 4 qubits", timed(q, lambda: q.PhaseRootNMask(3, 0b10001000100010000)), 2 * N * S)
    # (Prob of ONE qubit is served by the memoised all-marginals sweep below)
    row(tag + "ProbReg 2 high qubits", timed(q, lambda: q.be.prob_mask((1 << 20) | (1 << 21), 1 << 20)), N * S // 4, "reads the matching quarter")
    row(tag + "ProbMask 2 low qubits", timed(q, lambda: q.be.prob_mask(0b110, 0b010)), N * S, "low qubits: every sector is touched")
    row(tag + "ProbMask 3 qubits", timed(q, lambda: q.be.prob_mask((1 << 20) | (1 << 11) | (1 << 25), 1 << 20)), N * S // 8)
    row(tag + "ProbParity 4 qubits", timed(q, lambda: q.be.prob_parity(0b10001000100010000)), N * S)
    row(tag + "all single-qubit marginals", timed(q, lambda: q.Prob(5), setup=lambda: q.be.set_amplitude(0, q.be.get_amplitude(0))), N * S,
        "k_prob_all_bits: one sweep serves every Prob(q)")
    row(tag + "UpdateRunningNorm", timed(q, lambda: q.UpdateRunningNorm()), N * S)
    row(tag + "NormalizeState (scale only)", timed(q, lambda: q.be.normalize(0.999, 0.0, 0.0)), 2 * N * S)
    row(tag + "MAll sampling (no collapse)", timed(q, lambda: q.be.sample(0.4321)), N * S, "chunk sums + one 16 K-amplitude chunk")
    row(tag + "HighestProbAll", timed(q, lambda: q.be.highest_prob()), N * S)
    # ApplyM / ForceM keep one half: restore the state between repetitions (not timed)
    def reset():
        q.SetPermutation(0)
        for b in range(nq):
            q.H(b)
        q.Finish()
    row(tag + "ApplyM (ForceM q=20 -> 0)", timed(q, lambda: q.be.apply_m(1 << 20, 0, 1.4142135623730951 + 0j), reps=3, warm=1, setup=reset),
        3 * N * S // 2, "read+write the kept half, write zeros to the dropped half")
    del q

    # structure ops: (n-1) + 1 qubits
    a = QEngineCUDA(nq - 1, 0, random.Random(1), 1.0 + 0j, False, False, precision=prec)
    for b in range(nq - 1):
        a.H(b)
    a.T(2)
    a.Finish()

    def compose_once():
        b1 = QEngineCUDA(1, 0, random.Random(2), 1.0 + 0j, False, False, precision=prec)
        b1.H(0)
        a.Compose(b1)

    def undo_compose():
        if a.GetQubitCount() == nq:
            a.Dispose(nq - 1, 1, 0)
    # Dispose(perm) as the inverse keeps the (n-1)-qubit factor exactly (pure gather)
    ms_c = timed(a, compose_once, reps=3, warm=1, setup=undo_compose)
    row(tag + "Compose (%d)+1" % (nq - 1), ms_c, (N // 2) * S + N * S, "read 2^(n-1) + 2, write 2^n")
    undo_compose()

    def setup_n():
        if a.GetQubitCount() == nq - 1:
            compose_once()
    dest = QEngineCUDA(1, 0, random.Random(3), 1.0 + 0j, False, False, precision=prec)
    ms_d = timed(a, lambda: a.Decompose(nq - 1, dest), reps=3, warm=1, setup=setup_n)
    row(tag + "Decompose %d -> (%d)+1" % (nq, nq - 1), ms_d, 2 * N * S + (N // 2) * S, "two reads of 2^n (marginals, rebuild), write 2^(n-1)")
    ms_p = timed(a, lambda: a.Dispose(nq - 1, 1, 0), reps=3, warm=1, setup=setup_n)
    row(tag + "Dispose(perm) %d -> %d" % (nq, nq - 1), ms_p, (N // 2) * S * 2, "gather: read 2^(n-1), write 2^(n-1)")
    undo_compose()
    # page ops between two (n-1)-qubit engines on the same GPU (QPager's ShuffleBuffers / CopyStateVec)
    c = QEngineCUDA(nq - 1, 5, random.Random(4), 1.0 + 0j, False, False, precision=prec)
    row(tag + "ShuffleBuffers 2 x 2^%d (same GPU)" % (nq - 1), timed(a, lambda: a.ShuffleBuffers(c)), 2 * (N // 2) * S,
        "swap a's upper half with c's lower half: read+write 2 x 2^(n-2) amplitudes")
    row(tag + "CopyStateVec 2^%d" % (nq - 1), timed(a, lambda: c.CopyStateVec(a)), 2 * (N // 2) * S)
    del a, c, dest

os.makedirs("gpurun_out", exist_ok=True)
json.dump({"n": n, "peak_gbs": peak, "rows": rows}, open("gpurun_out/stream_table.json", "w"), indent=1)
