"""GPU test (sorted last on purpose: the newest entry points, after the established suites) of b200sv_set_rank_bits (the rank index as
constant virtual qubits) and b200sv_flush_carry (the under-filled tail of a window is handed back) on one device; the check body is
shared with the host-interpreter test (tests/carry_checks.py)."""
import random

import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("prec", [32, 64])
def test_rank_bits_and_carry_on_the_device(prec):
    from carry_checks import check_rank_bits_and_carry
    from qrack_b200.qengine import QEngineCUDA
    handed = 0
    for seed in range(3):
        handed += check_rank_bits_and_carry(lambda n: QEngineCUDA(n, 0, random.Random(1), 1.0 + 0j, False, False, deviceId=0, precision=prec),
                                            prec, seed=seed)
    assert handed > 0
