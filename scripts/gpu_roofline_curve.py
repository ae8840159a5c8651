"""Roofline curve of the fused sweep and of the QAlu basis-map kernel on one B200 (run under gpurun).

 (1) fused single-qubit sweep: W single-qubit gates (H / T / general U alternating) on W distinct tile qubits, one
     sweep each; physical GB/s = 2 * 2^n * S / (ms per sweep) against MEASURED_PEAKS.json.
 (2) QAlu: one out-of-place basis map per call (2 * 2^n * S algorithmic bytes; +1 x 2^n * S when the destination has
     to be cleared first).
Timing: CUDA events on the engine's stream (b200sv_timer_*), 2 warm-up + 5 timed repetitions, state >> L2.
"""
import json
import math
import os
import random
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qrack_b200 import QEngineCUDA  # noqa: E402

n = int(os.environ.get("N", "30"))
peak = json.load(open("MEASURED_PEAKS.json")).get("hbm_gbs", 6486.8) if os.path.exists("MEASURED_PEAKS.json") else 6486.8
out = {"n": n, "peak_gbs": peak, "fused": [], "alu": []}


def timed(q, fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    q.Finish()
    q.be.timer_begin()
    for _ in range(reps):
        fn()
    return q.be.timer_end() / reps


for prec in [int(p) for p in os.environ.get("PRECS", "32,64").split(",")]:
    S = 8 if prec == 32 else 16
    nq = n if prec == 32 else n - 1
    q = QEngineCUDA(nq, 0, random.Random(1), 1.0 + 0j, False, False, precision=prec)
    q.H(0)
    q.Finish()
    bytes_sweep = 2.0 * (1 << nq) * S
    # tile qubits reachable in one sweep: the low 6 plus any 7 (fp32) / 6 (fp64) others
    tile = [0, 1, 2, 3, 4, 5, 9, 14, 17, 20, 23, 26, nq - 1][: (13 if prec == 32 else 12)]
    for kind in os.environ.get("FUSED_KINDS", "H,T,U,mix").split(","):
        if not kind:
            continue
        for W in (1, 2, 4, 6, 8, 10, len(tile)):
            qs = tile[:W]

            def layer():
                for i, t in enumerate(qs):
                    k = kind if kind != "mix" else ("H", "T", "U")[i % 3]
                    if k == "H":
                        q.H(t)
                    elif k == "T":
                        q.T(t)
                    else:
                        q.U(t, 0.3 + 0.1 * i, 0.2, 0.1)
                q.be.flush()
            q.be.reset_stats()
            ms = timed(q, layer)
            st = q.be.stats()
            sweeps = st["fused_sweeps"] / 7.0
            gbs = bytes_sweep * sweeps / (ms * 1e-3) / 1e9
            rec = {"prec": prec, "kind": kind, "W": W, "ms": ms, "sweeps_per_layer": sweeps, "phys_gbs": gbs,
                   "phys_frac": gbs / peak, "alg_gbs": bytes_sweep * W / (ms * 1e-3) / 1e9}
            out["fused"].append(rec)
            print("fused fp%d %-3s W=%2d  %.3f ms  sweeps=%.1f  phys %.0f GB/s (%.2f of peak)  alg %.0f GB/s" % (
                prec, kind, W, ms, sweeps, gbs, gbs / peak, rec["alg_gbs"]), flush=True)
    # ---- QAlu
    L = 12
    tab = bytes(random.Random(3).randrange(256) for _ in range(2 << L))
    perm = list(range(1 << L))
    random.Random(4).shuffle(perm)
    hashtab = b"".join(int(v).to_bytes(2, "little") for v in perm)
    alu = [
        ("INC", 2, lambda: q.be.alu_inc(12345, 3, 20, 0)),
        ("CINC(2 ctrl)", 2, lambda: q.be.alu_inc(12345, 3, 20, (1 << 25) | (1 << 1))),
        ("ROL", 2, lambda: q.be.alu_rol(7, 0, nq)),
        ("INCDECC", 3, lambda: q.be.alu_incdecc(999, 2, 18, 24)),
        ("INCS", 2, lambda: q.be.alu_incs(999, 2, 18, 24)),
        ("MUL", 3, lambda: q.be.alu_muldiv(0, 5, 0, 10, 10, 0)),
        ("POWModNOut", 3, lambda: q.be.alu_modnout(2, 3, 1021, 0, 10, 10, 0)),
        ("IndexedADC", 3, lambda: q.be.alu_indexed(1, 0, L, L, 16, 28, 0, tab)),
        ("Hash", 3, lambda: q.be.alu_hash(4, L, hashtab)),
        ("PhaseFlipIfLess", 2, lambda: q.be.alu_phase_flip_if_less(777, 3, 12, -1)),
    ]
    for name, passes, fn in alu:
        ms = timed(q, fn, reps=3, warm=1)
        alg = 2.0 * (1 << nq) * S
        rec = {"prec": prec, "op": name, "ms": ms, "alg_gbs": alg / (ms * 1e-3) / 1e9, "alg_frac": alg / (ms * 1e-3) / 1e9 / peak,
               "phys_bytes_model": passes / 2.0 * alg}
        out["alu"].append(rec)
        print("alu   fp%d %-16s %.3f ms  alg %.0f GB/s (%.2f of peak)" % (prec, name, ms, rec["alg_gbs"], rec["alg_frac"]), flush=True)
    del q
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open(os.environ.get("OUT", "gpurun_out/roofline_curve.json"), "w"), indent=1)
