#!/usr/bin/env python
"""bench.py — gates/sec on BASELINE.json's headline workload + Apply2x2 HBM roofline (contract: see DESIGN.md §Measurement).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--qubits n] [--depth d]

A "step" = one replay of the whole seeded H/T/CNOT circuit (BASELINE configs[1]: 30 qubits, depth 40, 1800 gates, fp32)
on a state vector that is already resident in HBM.  `value` = gates/s over exactly K timed steps (CUDA events on the
engine's own stream, max over ranks); `e2e` = the same metric through the public QEngineCUDA API with host-side gate
submission, state (re)initialisation and a device->host read of per-qubit probabilities inside the timed region.
`--impl reference` times the compiled reference QEngineCPU (oracle/_ref) — or the oracle port when that binary did not
travel — on a bounded sample of the same workload on the host cores.
"""
import argparse
import json
import os
import random
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from qrack_b200 import qscript  # noqa: E402

METRIC = "gates/sec at 30q random circuit; Apply2x2 HBM GB/s vs roofline"


def measured_traffic(kernel_bytes_per_launch):
    """dram__bytes_read+write per launch of the fused sweep from the committed `ncu --set full` capture
    (profiles/r1_fused_v9_ncu_full.json, 28 qubits), scaled to this run's state size."""
    p = os.path.join(ROOT, "profiles", "r1_fused_v9_ncu_full.json")
    try:
        j = json.load(open(p))
        l = j["launches"][0]
        per28 = (float(l["dram__bytes_read.sum"]) + float(l["dram__bytes_write.sum"])) * 1e9
        return per28 / (2.0 * (1 << 28) * 8) * kernel_bytes_per_launch
    except Exception:
        return None


def fused_single_qubit_sweep_probe(q, n, amp_bytes, peak, reps=5, warm=2):
    """The north-star's own roofline target, measured live: ONE fused sweep of single-qubit gates (H on qubits 0 and 1)
    over the resident 2^n state; physical bytes 2 * 2^n * S / CUDA-event time on the engine's stream.  Never fatal."""
    try:
        def layer():
            q.H(0)
            q.H(1)
            q.be.flush()
        for _ in range(warm):
            layer()
        q.Finish()
        s0 = q.be.stats()
        q.be.timer_begin()
        for _ in range(reps):
            layer()
        ms = q.be.timer_end() / reps
        s1 = q.be.stats()
        sweeps = (s1["fused_sweeps"] - s0["fused_sweeps"]) / float(reps)
        if ms <= 0 or sweeps != 1.0:
            return {"gates": 2, "ms": ms, "sweeps": sweeps, "physical_gbs": None, "frac": None}
        gbs = 2.0 * (1 << n) * amp_bytes / (ms * 1e-3) / 1e9
        return {"gates": 2, "ms": ms, "sweeps": sweeps, "physical_gbs": gbs, "frac": gbs / peak}
    except Exception as e:  # diagnostic only: the headline numbers above are already taken
        return {"error": repr(e)[:200]}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.rows = []
        self.stop_flag = False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        sm = [float(r[1]) for r in self.rows if len(r) > 2 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for k, nm in enumerate(names):
                if len(r) > 5 + k and r[5 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


def gate_calls(text):
    """Pre-parse the script into (method name, int args) so the timed loop is only API calls."""
    calls = []
    for _, t in qscript.parse(text):
        if t[0] in ("qubits", "TIC", "TOC"):
            continue
        calls.append((t[0], tuple((float(x) if ("." in x or "e" in x) else int(x)) for x in t[1:])))
    return calls


def algorithmic_bytes(calls, n, amp_bytes):
    """B(gate) = 2 * 2^(n-c) * S, c = number of control qubits."""
    tot = 0
    for name, args in calls:
        c = 1 if name in ("CNOT", "CZ", "CY", "AntiCNOT", "Swap", "CPhaseRootN") else 0
        tot += 2 * (1 << (n - c)) * amp_bytes
    return tot


def cpu_reference_sample(n, depth, seed, prec, budget_gates, threads=None):
    """Time the compiled reference (oracle/_ref/ref_harness) on the first `budget_gates` gates of the workload."""
    harness = os.path.join(ROOT, "oracle", "_ref", "ref_harness_f%d" % prec)
    full = qscript.random_htcnot(n, depth, seed=seed, timed=False).splitlines()
    sample = [full[0], "TIC"] + full[1:1 + budget_gates] + ["TOC"]
    if os.path.exists(harness):
        with tempfile.TemporaryDirectory() as td:
            sp = os.path.join(td, "s.qs")
            open(sp, "w").write("\n".join(sample) + "\n")
            cmd = [harness, sp, "--time"] + (["--threads", str(threads)] if threads else [])
            out = subprocess.run(cmd, capture_output=True, text=True, check=True).stdout
        j = json.loads(out.strip().splitlines()[-1])
        return {"value": j["ops"] / j["seconds"], "unit": "gates/s", "cores": j["threads"], "kind": "reference",
                "sample": "first %d gates of the %d-qubit depth-%d H/T/CNOT circuit on QEngineCPU (fp%d), %.1f s" %
                          (j["ops"], n, depth, prec, j["seconds"]), "seconds": j["seconds"], "gates": j["ops"]}
    # oracle port (single-threaded C restatement)
    from oracle.restate_engine import QEngineRestate
    q = QEngineRestate(n, 0, random.Random(1), 1.0 + 0j, False, False, precision=prec)
    calls = gate_calls("\n".join(sample) + "\n")
    t0 = time.perf_counter()
    for name, args in calls:
        getattr(q, name)(*args)
    dt = time.perf_counter() - t0
    return {"value": len(calls) / dt, "unit": "gates/s", "cores": 1, "kind": "port",
            "sample": "first %d gates of the %d-qubit circuit on the oracle C restatement (fp%d), %.1f s" % (len(calls), n, prec, dt),
            "seconds": dt, "gates": len(calls)}


def dist_setup(n_gpus):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        return rank, world, local, dist
    return rank, world, local, None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--qubits", type=int, default=30)
    ap.add_argument("--depth", type=int, default=40)
    ap.add_argument("--precision", type=int, default=32)
    ap.add_argument("--seed", type=int, default=20250921)
    ap.add_argument("--fusion", type=int, default=1)
    ap.add_argument("--cpu-sample-gates", type=int, default=45)
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"],
                    help="N>1: qubit exchange by the fused NVLink peer-store kernel (default) or NCCL all_to_all + local swaps")
    ap.add_argument("--workload", default="htcnot", choices=["htcnot", "qft", "qv"],
                    help="htcnot = BASELINE configs[1] (default, the headline); qft = configs[2]; qv = configs[3]-style layers")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    shard_bits = max(0, args.gpus).bit_length() - 1 if args.gpus > 1 else 0
    if args.impl == "reference" and shard_bits:
        # same configuration as the b200 arm at this N (weak scaling: 2^qubits amplitudes per GPU => qubits + log2 N in all);
        # the sample shrinks with the state so that the run stays within minutes
        args.qubits += shard_bits
        args.cpu_sample_gates = max(12, args.cpu_sample_gates >> shard_bits)
    n, depth, prec = args.qubits, args.depth, args.precision
    amp_bytes = 8 if prec == 32 else 16
    if args.workload == "qft":
        text = qscript.qft(n, seed=11, timed=False)
    elif args.workload == "qv":
        text = qscript.quantum_volume(n, depth=depth, seed=33, timed=False)
    else:
        text = qscript.random_htcnot(n, depth, seed=args.seed, timed=False)
    calls = gate_calls(text)
    gates = len(calls)
    if args.workload == "qft":
        gates = sum(1 for c in calls if c[0] == "H") + n + n * (n - 1) // 2  # init H's + QFT's own H and CPhaseRootN gates
    workload = {"htcnot": "%d-qubit random circuit (H/T/CNOT, depth %d, %d gates), fp%d amplitudes" % (n, depth, gates, prec),
                "qft": "%d-qubit QFT (%d H + %d controlled-phase), fp%d amplitudes" % (n, n, n * (n - 1) // 2, prec),
                "qv": "%d-qubit quantum-volume layers (AI + CNOT matching, depth %d, %d gates), fp%d" % (n, depth, gates, prec)}[args.workload]
    dtype = "f32" if prec == 32 else "f64"

    if args.impl == "reference":
        rank = int(os.environ.get("RANK", "0"))
        if rank != 0:
            return
        steps = max(args.steps, 1)
        per_step = max(4, args.cpu_sample_gates // max(1, steps))
        vals = []
        cb = None
        for _ in range(max(args.warmup, 0) and 1):
            cpu_reference_sample(min(n, 24), depth, args.seed, prec, per_step)
        t_total, g_total = 0.0, 0
        for _ in range(steps):
            cb = cpu_reference_sample(n, depth, args.seed, prec, per_step)
            t_total += cb["seconds"]
            g_total += cb["gates"]
        v = g_total / t_total
        cb["value"] = v
        line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "gates/s", "n_gpus": args.gpus, "steps": steps,
                "warmup": args.warmup, "ms_per_step": 1e3 * t_total / steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": dtype, "data": "synthetic",
                "config": {"workload": workload, "sample_gates_per_step": per_step, "engine": "QEngineCPU (reference, host cores)"},
                "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": v, "unit": "gates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    rank, world, local, dist = dist_setup(args.gpus)
    from qrack_b200 import QEngineCUDA

    sharded = world > 1
    if sharded:
        # ONE state vector of n = qubits + log2(N) qubits sharded over the N GPUs (weak scaling: 2^qubits amplitudes per GPU)
        import torch
        from qrack_b200.sharded import QEngineSharded, cuda_engine_factory
        k = world.bit_length() - 1
        n = args.qubits + k
        text = {"qft": lambda: qscript.qft(n, seed=11, timed=False),
                "qv": lambda: qscript.quantum_volume(n, depth=depth, seed=33, timed=False),
                "htcnot": lambda: qscript.random_htcnot(n, depth, seed=args.seed, timed=False)}[args.workload]()
        calls = gate_calls(text)
        gates = len(calls)
        workload = "%d-qubit %s circuit (%d gates), fp%d amplitudes, 2^%d amplitudes per GPU" % (n, args.workload, gates, prec, args.qubits)
        os.environ["B200SV_FUSED"] = os.environ.get("B200SV_FUSED", "")
        q = QEngineSharded(n, 0, random.Random(1), 1.0 + 0j, precision=prec, dist=dist, world=world, rank=rank,
                           device=torch.device("cuda", local), make_engine=cuda_engine_factory(local, prec),
                           p2p=(args.exchange == "p2p"))
    else:
        q = QEngineCUDA(n, 0, random.Random(1), 1.0 + 0j, False, False, deviceId=local, precision=prec)
        q.be.set_fusion(args.fusion)

    def replay():
        for name, a in calls:
            getattr(q, name)(*a)

    def barrier():
        if dist is not None:
            import torch
            dist.barrier()
            torch.cuda.synchronize()

    def get_stats():
        return q.be.shard.stats() if sharded else q.be.stats()

    # ---- device-timed arm: state resident in HBM, CUDA events on the stream the kernels run on --------------------
    for _ in range(args.warmup):
        q.SetPermutation(0, 1.0 + 0j)
        replay()
        q.Finish()
    sampler = ClockSampler(local)
    sampler.start()
    ms_steps = []
    ex0 = q.be.exchanges if sharded else 0
    stats0 = get_stats()
    if not sharded:
        q.be.reset_stats()
        stats0 = {k: 0 for k in stats0}
    barrier()
    for _ in range(args.steps):
        q.SetPermutation(0, 1.0 + 0j)
        q.Finish()
        if sharded:
            import torch
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()          # the local engines run on torch's current stream (b200sv_set_stream), like NCCL
            replay()
            q.be.flush()
            q.be.loc.be.flush()
            e1.record()
            e1.synchronize()
            ms_steps.append(e0.elapsed_time(e1))
        else:
            q.be.timer_begin()
            replay()
            ms_steps.append(q.be.timer_end())
    barrier()
    stats1 = get_stats()
    stats = {k: stats1[k] - stats0.get(k, 0) for k in stats1}
    exchanges = (q.be.exchanges - ex0) if sharded else 0
    # ---- end-to-end arm: public API, host submission + init + result read inside the timed region -----------
    h2d = gates * (8 * 8 + 8 * 4)  # per gate: 8 doubles of matrix + offsets/powers words crossing the C ABI
    d2h = n * 8
    e2e_steps = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        q.SetPermutation(0, 1.0 + 0j)
        replay()
        probs = [q.Prob(b) for b in range(n)]
        e2e_steps.append(time.perf_counter() - t0)
    barrier()
    sampler.stop_flag = True
    sampler.join(timeout=2)

    ms_total = sum(ms_steps)
    e2e_total = sum(e2e_steps)
    if dist is not None:
        import torch
        t = torch.tensor([ms_total, e2e_total], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total, e2e_total = t.tolist()
    total_gates = gates * args.steps  # one circuit on one (possibly sharded) state vector: whole-job gate count
    value = total_gates / (ms_total / 1e3)
    e2e_value = total_gates / e2e_total

    if rank == 0:
        peak, peak_src = peaks()
        launches = stats["kernel_launches"]
        swept = stats["bytes_swept"]
        kernel_ms = ms_total / max(1, launches) * 1.0
        achieved = (swept / 1e9) / (ms_total / 1e3) if ms_total > 0 else 0.0
        alg = algorithmic_bytes(calls, n, amp_bytes) * args.steps
        if sharded:
            swept = swept  # bytes swept by rank 0's local kernels; the roofline entry is per GPU
        line = {
            "metric": METRIC, "value": value, "unit": "gates/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": dtype, "data": "synthetic",
            "config": {"workload": workload, "l2_policy": "state vector (%.1f GiB) is far larger than the 126 MB L2" %
                       ((1 << n) * amp_bytes / 2 ** 30), "fusion": args.fusion,
                       "parallelism": "1 GPU" if world == 1 else
                       "1 state vector sharded over %d GPUs (top %d qubits = rank), qubit exchange: %s" % (world, world.bit_length() - 1, "fused NVLink peer-store kernel" if args.exchange == "p2p" else "NCCL all_to_all_single + local swap sweeps"),
                       "exchanges_per_step": (exchanges / max(1, args.steps)) if sharded else 0},
            "gpu_launches": int(launches),
            "e2e": {"value": e2e_value, "unit": "gates/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "includes": "SetPermutation + host gate submission through the C ABI + Prob(q) for every qubit"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": (measured_traffic(swept / max(1, launches)) if stats["fused_sweeps"] else None), "peak_source": peak_src, "kernel": "fused sweep" if stats["fused_sweeps"] else "k_apply2x2",
                         "bytes_per_launch": swept / max(1, launches), "ms_per_launch": kernel_ms,
                         "algorithmic_gbs": (alg / 1e9) / (ms_total / 1e3),
                         "algorithmic_frac": (alg / 1e9) / (ms_total / 1e3) / peak,
                         "note": "achieved/frac = PHYSICAL bytes of the launch (2 * 2^n * S) / duration; algorithmic_* credits "
                                 "sum_W 2*2^(n-c)*S per fused gate (SURVEY 8d); light sweeps reach 0.91 of peak "
                                 "(profiles/r1_roofline_curve_v9_H.json)",
                         "fused_sweeps": int(stats["fused_sweeps"]),
                         "fused_gates": int(stats["fused_gates"])},
            "clocks": sampler.summary(),
        }
        if world == 1 and stats["fused_sweeps"]:
            line["roofline"]["fused_single_qubit_sweep"] = fused_single_qubit_sweep_probe(q, n, amp_bytes, peak)
        if world == 1 and not args.skip_cpu_baseline:
            cb = cpu_reference_sample(n, depth, args.seed, prec, args.cpu_sample_gates)
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
