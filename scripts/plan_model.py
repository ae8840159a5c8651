"""Offline cost model of the fused sweep schedule (no GPU): plans a circuit with the real planner/encoder
(b200sv_plan_dry_run + B200SV_FUSED_DEBUG dump) and predicts the step time from

    T = 2.9 ms x sweeps + 0.9 ms x (passes - sweeps + staged copies) + 0.165 ms x active ops      (30 qubits, fp32)

fitted on the r1 measurements of the v9 kernel (51 sweeps / 156 passes / 1046 active ops = 416 ms measured, 415 modelled;
single-pass sweep 2.9 ms, +0.9 ms per extra pass from profiles/r1_roofline_curve_v9_H.json).  'active ops' counts an op
with an outer (per-tile) control as one half.  Scale: every term is proportional to 2^n amplitudes.

    python scripts/plan_model.py [htcnot|qv|qft|u3] [qubits] [precision]     (knobs via B200SV_FUSED / B200SV_PLAN_SEARCH)
"""
import collections
import ctypes
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def gate_arrays(text):
    T, C, K = [], [], []
    for ln in text.splitlines():
        w = ln.split()
        if not w:
            continue
        if w[0] == "H":
            T.append(int(w[1])); C.append(0); K.append(4)
        elif w[0] in ("T", "S", "Z", "IT", "IS"):
            T.append(int(w[1])); C.append(0); K.append(1)
        elif w[0] == "X":
            T.append(int(w[1])); C.append(0); K.append(2)
        elif w[0] == "CNOT":
            T.append(int(w[2])); C.append(1 << int(w[1])); K.append(2)
        elif w[0] == "CZ":
            T.append(int(w[2])); C.append(1 << int(w[1])); K.append(1)
        elif w[0] in ("AI", "IAI", "U"):
            T.append(int(w[1])); C.append(0); K.append(3)
    return T, C, K


def child(kind, n, prec):
    from qrack_b200 import _abi, qscript
    text = {"htcnot": lambda: qscript.random_htcnot(n, 40, seed=20250921, timed=False),
            "qv": lambda: qscript.quantum_volume(n, depth=40, seed=33, timed=False),
            "u3": lambda: qscript.random_u3_cnot(n, 20, seed=7)}.get(kind, lambda: None)()
    if kind == "qft":
        T, C, K = [], [], []
        for i in range(n):
            hb = n - 1 - i
            for j in range(i):
                T.append(hb + 1 + j); C.append(1 << hb); K.append(1)
            T.append(hb); C.append(0); K.append(4)
    else:
        T, C, K = gate_arrays(text)
    N = len(T)
    ns, npass = ctypes.c_int(), ctypes.c_int()
    lib = _abi.load()
    rc = lib.b200sv_plan_dry_run(n, prec, N, (ctypes.c_int * N)(*T), (ctypes.c_uint64 * N)(*C), (ctypes.c_int * N)(*K),
                                 ctypes.byref(ns), ctypes.byref(npass))
    print("PLAN %d %d %d %d" % (rc, N, ns.value, npass.value))


def main():
    kind = sys.argv[1] if len(sys.argv) > 1 else "htcnot"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    prec = int(sys.argv[3]) if len(sys.argv) > 3 else 32
    if os.environ.get("_PLAN_MODEL_CHILD"):
        return child(kind, n, prec)
    env = dict(os.environ, _PLAN_MODEL_CHILD="1", B200SV_FUSED_DEBUG="1")
    r = subprocess.run([sys.executable, os.path.abspath(__file__), kind, str(n), str(prec)], env=env, capture_output=True, text=True)
    plan = [l for l in r.stdout.splitlines() if l.startswith("PLAN")][0].split()
    gates, sweeps, passes = int(plan[2]), int(plan[3]), int(plan[4])
    # the dump also contains the attempts that did not fit; keep the last `sweeps` committed programs' statistics simple:
    ops = collections.Counter()
    staged = 0
    dumped_sweeps = 0
    for l in r.stderr.splitlines():
        ls = l.strip()
        if ls.startswith("pass"):
            for t in ls.split(":", 1)[1].split():
                ops[re.sub(r"\(.*\)|\.\d", "", t)] += 1
        elif ls.startswith("sweep:"):
            dumped_sweeps += 1
            staged += ("directIn 0" in ls) + ("directOut 0" in ls)
    scale_dump = sweeps / max(1, dumped_sweeps)          # retried encodings inflate the dump
    total = sum(ops.values()) * scale_dump
    active = (sum(ops.values()) - 0.5 * sum(v for k, v in ops.items() if k.endswith("o") or k.endswith("so"))) * scale_dump
    amp = 2.0 ** (n - 30) * (2.0 if prec == 64 else 1.0)
    t = amp * (2.9 * sweeps + 0.9 * (passes - sweeps + staged * scale_dump) + 0.165 * active)
    print("%s n=%d fp%d: %d gates -> %d sweeps, %d passes, %.0f device ops (%.0f active), %d staged copies" % (
        kind, n, prec, gates, sweeps, passes, total, active, staged))
    print("  op mix: %s" % dict(ops))
    print("  modelled step: %.1f ms  = %.0f gates/s   (fp64 and n != 30 scaled by bytes; model fitted on fp32 n=30)" % (t, gates / t * 1e3))


if __name__ == "__main__":
    main()
