"""bench.py's JSON contract, exercised without a GPU: the N=1 main path runs end to end against a stand-in engine (the
oracle restatement with the CUDA backend's stats/timer hooks stubbed), and the reference arm runs for real on host cores.
Guards the keys the driver reads (metric/value/unit/..., roofline, cpu_baseline, e2e, gpu_launches, clocks)."""
import json
import os
import subprocess
import sys

import pytest

import util

SHIM = r'''
import sys, random, time
sys.path.insert(0, %r)
import bench
from oracle.restate_engine import QEngineRestate
class Fake(QEngineRestate):
    def __init__(self, n, perm, rng, phase, a, b, deviceId=0, precision=32):
        super().__init__(n, perm, rng, phase, a, b, precision=precision)
        be = self.be
        st = {'kernel_launches': 0, 'fused_sweeps': 0, 'fused_gates': 0, 'bytes_swept': 0, 'gates_submitted': 0, 'single_launches': 0}
        t = [0.0]
        be.set_fusion = lambda f: None
        def flush():
            st['kernel_launches'] += 1; st['fused_sweeps'] += 1; st['bytes_swept'] += 2 * (1 << n) * 8
        be.flush = flush
        be.stats = lambda: dict(st)
        be.reset_stats = lambda: [st.__setitem__(k, 0) for k in st]
        be.timer_begin = lambda: t.__setitem__(0, time.perf_counter())
        def tend():
            flush(); return (time.perf_counter() - t[0]) * 1e3
        be.timer_end = tend
import qrack_b200
qrack_b200.QEngineCUDA = Fake
sys.argv = ['bench.py', '--qubits', '9', '--steps', '2', '--warmup', '3', '--skip-cpu-baseline']
bench.main()
'''


def _last_json(out):
    return json.loads([l for l in out.splitlines() if l.startswith("{")][-1])


def test_b200_arm_json_contract():
    r = subprocess.run([sys.executable, "-c", SHIM % util.ROOT], capture_output=True, text=True, timeout=600, cwd=util.ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    j = _last_json(r.stdout)
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "gpu_launches", "e2e", "roofline", "clocks"):
        assert k in j, k
    assert j["n_gpus"] == 1 and j["steps"] == 2 and j["warmup"] >= 3 and j["unit"] == "gates/s" and j["dtype"] == "f32"
    assert j["vs_baseline"] is None and "workload" in j["config"] and "model" not in j["config"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic", "algorithmic_gbs", "algorithmic_frac", "fused_single_qubit_sweep"):
        assert k in j["roofline"], k
    for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
        assert k in j["e2e"], k
    assert j["gpu_launches"] > 0
    # self-checks computed in the run: norm, marginals, mirror-circuit return
    assert j["check"]["ok"] is True and abs(j["check"]["norm_minus_1"]) < 1e-4 and abs(j["check"]["mirror_return_prob"] - 1) < 1e-3


@pytest.mark.parametrize("workload,extra", [("grover", ["--depth", "2"]), ("qft", ["--precision", "64"]), ("qv", ["--depth", "3"])])
def test_b200_arm_other_baseline_workloads(workload, extra):
    shim = (SHIM % util.ROOT).replace("'--skip-cpu-baseline']", "'--skip-cpu-baseline', '--workload', %r] + %r" % (workload, extra))
    r = subprocess.run([sys.executable, "-c", shim], capture_output=True, text=True, timeout=600, cwd=util.ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    j = _last_json(r.stdout)
    assert j["check"]["ok"] is True, j["check"]
    if workload == "grover":
        assert abs(j["check"]["grover_success_prob"] - j["check"]["grover_law_sin2((2k+1)asin(2^-n/2))"]) < 1e-4


def test_reference_arm_runs_on_host_cores():
    if util.ref_harness(32) is None:
        pytest.skip("oracle/_ref not built")
    r = subprocess.run([sys.executable, os.path.join(util.ROOT, "bench.py"), "--impl", "reference", "--qubits", "16", "--steps", "2",
                        "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=util.ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    j = _last_json(r.stdout)
    assert j["impl"] == "reference" and j["unit"] == "gates/s" and j["value"] > 0
    assert j["cpu_baseline"]["kind"] == "reference" and j["cpu_baseline"]["cores"] >= 1
    assert j["e2e"]["h2d_bytes_per_step"] == 0 and j["e2e"]["d2h_bytes_per_step"] == 0
    # the sample is a fixed prefix of the circuit (>= one full layer incl. CNOTs), whatever --steps is, and `config` is the b200 arm's
    assert j["reference_sample"]["ops_total"] >= 60 and len(j["reference_sample"]["segment_seconds"]) == 2
    r5 = subprocess.run([sys.executable, os.path.join(util.ROOT, "bench.py"), "--impl", "reference", "--qubits", "16", "--steps", "20",
                         "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=util.ROOT)
    j5 = _last_json(r5.stdout)
    assert j5["reference_sample"]["ops_total"] == 60 and j5["config"] == j["config"]
    assert set(j["config"]) == {"workload", "l2_policy", "fusion", "parallelism"}
    # other ranks of a torchrun launch do no work
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(util.ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--qubits", "16"],
                       capture_output=True, text=True, timeout=120, cwd=util.ROOT, env=env)
    assert r.returncode == 0 and not r.stdout.strip()
