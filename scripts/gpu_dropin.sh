#!/bin/bash
# The reference's own unit tests / harness running on the B200 drop-in (dropin/_build, prebuilt in the build container)
set -u
export LD_LIBRARY_PATH=$PWD/qrack_b200:${LD_LIBRARY_PATH:-}
mkdir -p gpurun_out
U=dropin/_build/f32/unittest_b200
H=dropin/_build/harness_b200_f32
R=oracle/_ref/ref_harness_f32
LIST="test_cnot,test_apply_single_bit,test_global_phase,test_qft_h,test_compose,test_decompose,test_dispose,test_dispose_perm,test_allocate,test_trydecompose,test_prob*,test_cprob,test_forcem,test_getamplitude,test_getquantumstate,test_getprobs,test_normalize,test_grover,test_h_cnot_rand,test_m,test_mreg,test_swap,test_t,test_ccnot,test_ucmtrx,test_multishotmeasuremask,test_bell_m,test_mirror_circuit*"
echo "== reference unittest on QEngineCUDA drop-in (--layer-qengine --proc-cuda)"
timeout 900 $U --layer-qengine --proc-cuda --disable-hardware-rng "$LIST" > gpurun_out/dropin_unittest_qengine_full.log 2>&1; grep -B3 -A12 "FAILED" gpurun_out/dropin_unittest_qengine_full.log | head -60; tail -4 gpurun_out/dropin_unittest_qengine_full.log | tee gpurun_out/dropin_unittest_qengine.log
ALU="test_rol,test_ror,test_inc,test_incs,test_incc,test_incsc,test_cinc,test_dec,test_decs,test_decc,test_decsc,test_cdec,test_mul,test_div,test_mulmodnout,test_imulmodnout,test_powmodnout,test_cmul,test_cdiv,test_cmulmodnout,test_cimulmodnout,test_cpowmodnout,test_c_phase_flip_if_less,test_superposition_reg,test_adc_superposition_reg,test_sbc_superposition_reg,test_superposition_reg_long,test_adc_superposition_reg_long_index,test_sbc_superposition_reg_long_index,test_hash,test_fulladd,test_ifulladd,test_adc,test_iadc,test_cfulladd,test_cifulladd,test_cadc,test_ciadc,test_set_reg,test_amplitude_amplification,test_basis_change"
echo "== reference ALU unittests on the drop-in's native QAlu kernels (--layer-qengine --proc-cuda)"
timeout 900 $U --layer-qengine --proc-cuda --disable-hardware-rng "$ALU" > gpurun_out/dropin_unittest_alu_full.log 2>&1; grep -B3 -A12 "FAILED" gpurun_out/dropin_unittest_alu_full.log | head -60; tail -4 gpurun_out/dropin_unittest_alu_full.log | tee gpurun_out/dropin_unittest_alu.log
echo "== reference unittest with QUnit over the drop-in (--layer-qunit --proc-cuda): callers above the engine"
timeout 900 $U --layer-qunit --proc-cuda --disable-hardware-rng "$LIST,$ALU" > gpurun_out/dropin_unittest_qunit_full.log 2>&1; grep -B12 "FAILED" gpurun_out/dropin_unittest_qunit_full.log | grep "^test_" | sort | uniq -c | head -20; tail -4 gpurun_out/dropin_unittest_qunit_full.log | tee gpurun_out/dropin_unittest_qunit.log
echo "== reference unittest with QHybrid over the drop-in (--layer-qunit --proc-hybrid: test_main.cpp has no qengine-layer slot for QHybrid)"
timeout 900 $U --layer-qunit --proc-hybrid --disable-hardware-rng "$LIST,$ALU" > gpurun_out/dropin_unittest_qhybrid_full.log 2>&1; grep -B12 "FAILED" gpurun_out/dropin_unittest_qhybrid_full.log | grep "^test_" | sort | uniq -c | head -20; tail -4 gpurun_out/dropin_unittest_qhybrid_full.log | tee gpurun_out/dropin_unittest_qhybrid.log
echo "== reference unittest on QPager over the drop-in (--layer-qpager --proc-cuda)"
timeout 900 $U --layer-qpager --proc-cuda --disable-hardware-rng "$LIST" > gpurun_out/dropin_unittest_qpager_full.log 2>&1; grep -A2 "^tests.cpp.*FAILED\|^\S.*tests.cpp:[0-9]*: FAILED" gpurun_out/dropin_unittest_qpager_full.log | grep -v "^--" | head -40; grep -B12 "FAILED" gpurun_out/dropin_unittest_qpager_full.log | grep "^test_" | sort | uniq -c | head -20; tail -4 gpurun_out/dropin_unittest_qpager_full.log | tee gpurun_out/dropin_unittest_qpager.log
echo "== harness parity: drop-in vs compiled reference on the C1 circuit (20 q)"
python - <<'PY' 2>&1 | tee gpurun_out/dropin_parity.log
import subprocess, numpy as np, sys
sys.path.insert(0,'.')
from qrack_b200 import qscript
open('/tmp/c1.qs','w').write(qscript.random_htcnot(20,40,seed=20250921,timed=False))
subprocess.run(['oracle/_ref/ref_harness_f32','/tmp/c1.qs','--dump','/tmp/ref'],check=True)
for eng in ('cuda','pager-cuda:17','hybrid'):
    subprocess.run(['dropin/_build/harness_b200_f32','/tmp/c1.qs','--dump','/tmp/dev','--engine',eng],check=True)
    a=np.fromfile('/tmp/ref.0.bin',dtype=np.complex64); b=np.fromfile('/tmp/dev.0.bin',dtype=np.complex64)
    print(eng, 'max |delta amp| vs QEngineCPU = %.3e'%np.abs(a-b).max())
PY
echo "== done"
