/* qalu_restate_impl.h — TEST INFRASTRUCTURE (oracle), NOT product code.
 *
 * Plain-C, single-threaded restatement of the reference's QEngineCPU arithmetic
 * (unitaryfoundation/qrack, /root/reference/src/qengine/arithmetic.cpp, dense branches), included once per precision by
 * qengine_restate.c.  One loop per reference method, written as the reference iterates (source index -> destination),
 * so that it is an independent check of the descriptor-driven CUDA kernel.  PARITY PIN: tests/test_oracle_pin.py replays
 * tests/golden/alu_*.qs on the compiled reference (oracle/_ref) and on this restatement.
 */

#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, SUF)

static REAL* FN(alu_new)(int nq) { return (REAL*)calloc((size_t)2 << nq, sizeof(REAL)); }
static void FN(alu_adopt)(REAL* psi, REAL* out, int nq)
{
    memcpy(psi, out, ((size_t)2 << nq) * sizeof(REAL));
    free(out);
}
#define ALU_MOVE(dst, src, sgn)                                                                                        \
    do {                                                                                                               \
        out[2 * (dst)] = (sgn)*psi[2 * (src)];                                                                         \
        out[2 * (dst) + 1] = (sgn)*psi[2 * (src) + 1];                                                                 \
    } while (0)

/* src/common/functions.cpp:214-233 */
static int FN(alu_is_overflow_add)(uint64_t a, uint64_t b, uint64_t signMask, uint64_t lengthPower)
{
    if ((a & b & signMask) != 0) {
        a = ((~a) & (lengthPower - 1U)) + 1U;
        b = ((~b) & (lengthPower - 1U)) + 1U;
        if ((a + b) > signMask) {
            return 1;
        }
    } else if (((~a) & (~b) & signMask) != 0) {
        if ((a + b) >= signMask) {
            return 1;
        }
    }
    return 0;
}

/* src/common/functions.cpp:77-95 (recursive there; same 64-bit wrap-around) */
static uint64_t FN(alu_int_pow)(uint64_t base, uint64_t power)
{
    if (power == 0U) {
        return 1U;
    }
    if (power == 1U) {
        return base;
    }
    uint64_t tmp = FN(alu_int_pow)(base, power >> 1U);
    tmp *= tmp;
    if (power & 1U) {
        tmp *= base;
    }
    return tmp;
}

static uint64_t FN(alu_value)(const unsigned char* values, uint64_t idx, int bytes)
{
    uint64_t v = 0;
    for (int j = 0; j < bytes; ++j) {
        v |= (uint64_t)values[idx * (uint64_t)bytes + (uint64_t)j] << (8U * j);
    }
    return v;
}

/* QEngineCPU::ROL — arithmetic.cpp:23-70 */
void FN(orc_rol)(REAL* psi, int nq, int shift, int start, int length)
{
    if (!length) {
        return;
    }
    shift %= length;
    if (!shift) {
        return;
    }
    const uint64_t n = 1ULL << nq, lengthMask = (1ULL << length) - 1U, regMask = lengthMask << start;
    REAL* out = FN(alu_new)(nq);
    for (uint64_t lcv = 0; lcv < n; ++lcv) {
        const uint64_t otherRes = lcv & ~regMask;
        const uint64_t regInt = (lcv & regMask) >> start;
        const uint64_t outInt = (regInt >> (length - shift)) | ((regInt << shift) & lengthMask);
        ALU_MOVE((outInt << start) | otherRes, lcv, 1);
    }
    FN(alu_adopt)(psi, out, nq);
}

/* QEngineCPU::INC :73-118 and CINC :121-172 (controlMask == 0: INC) */
void FN(orc_inc)(REAL* psi, int nq, uint64_t toAdd, int start, int length, uint64_t controlMask)
{
    if (!length) {
        return;
    }
    const uint64_t n = 1ULL << nq, lengthMask = (1ULL << length) - 1U, inOutMask = lengthMask << start;
    toAdd &= lengthMask;
    if (!toAdd) {
        return;
    }
    REAL* out = FN(alu_new)(nq);
    memcpy(out, psi, ((size_t)2 << nq) * sizeof(REAL)); /* CINC: nStateVec->copy(stateVec) (:155) */
    for (uint64_t lcv = 0; lcv < n; ++lcv) {
        if ((lcv & controlMask) != controlMask) {
            continue;
        }
        const uint64_t otherRes = lcv & ~inOutMask;
        const uint64_t outInt = (((lcv & inOutMask) >> start) + toAdd) & lengthMask;
        ALU_MOVE((outInt << start) | otherRes, lcv, 1);
    }
    FN(alu_adopt)(psi, out, nq);
}

/* QEngineCPU::INCDECC :175-224 — only sources with the carry qubit clear are moved */
void FN(orc_incdecc)(REAL* psi, int nq, uint64_t toMod, int start, int length, int carryIndex)
{
    if (!length) {
        return;
    }
    const uint64_t n = 1ULL << nq, lengthPower = 1ULL << length, lengthMask = lengthPower - 1U;
    toMod &= lengthMask;
    if (!toMod) {
        return;
    }
    const uint64_t carryMask = 1ULL << carryIndex, inOutMask = lengthMask << start;
    REAL* out = FN(alu_new)(nq);
    for (uint64_t lcv = 0; lcv < n; ++lcv) {
        if (lcv & carryMask) {
            continue;
        }
        const uint64_t otherRes = lcv & ~(inOutMask | carryMask);
        const uint64_t outInt = ((lcv & inOutMask) >> start) + toMod;
        const uint64_t outRes = (outInt < lengthPower) ? ((outInt << start) | otherRes)
                                                       : (((outInt - lengthPower) << start) | otherRes | carryMask);
        ALU_MOVE(outRes, lcv, 1);
    }
    FN(alu_adopt)(psi, out, nq);
}

/* QEngineCPU::INCS :227-309 */
void FN(orc_incs)(REAL* psi, int nq, uint64_t toAdd, int start, int length, int overflowIndex)
{
    if (!length) {
        return;
    }
    const uint64_t n = 1ULL << nq, lengthPower = 1ULL << length, lengthMask = lengthPower - 1U;
    toAdd &= lengthMask;
    if (!toAdd) {
        return;
    }
    const uint64_t overflowMask = 1ULL << overflowIndex, signMask = 1ULL << (length - 1), inOutMask = lengthMask << start;
    REAL* out = FN(alu_new)(nq);
    for (uint64_t lcv = 0; lcv < n; ++lcv) {
        const uint64_t otherRes = lcv & ~inOutMask;
        const uint64_t inOutInt = (lcv & inOutMask) >> start;
        const uint64_t outInt = inOutInt + toAdd;
        const uint64_t outRes =
            (outInt < lengthPower) ? ((outInt << start) | otherRes) : (((outInt - lengthPower) << start) | otherRes);
        const int isOverflow = FN(alu_is_overflow_add)(inOutInt, toAdd, signMask, lengthPower);
        if (isOverflow && ((outRes & overflowMask) == overflowMask)) {
            ALU_MOVE(outRes, lcv, -1);
        } else {
            ALU_MOVE(outRes, lcv, 1);
        }
    }
    FN(alu_adopt)(psi, out, nq);
}

/* QEngineCPU::INCDECSC — carry only :312-361 (overflowIndex < 0) and overflow flag + carry :364-419 */
void FN(orc_incdecsc)(REAL* psi, int nq, uint64_t toMod, int start, int length, int overflowIndex, int carryIndex)
{
    if (!length) {
        return;
    }
    const uint64_t n = 1ULL << nq, lengthPower = 1ULL << length, lengthMask = lengthPower - 1U;
    toMod &= lengthMask;
    if (!toMod) {
        return;
    }
    const uint64_t overflowMask = (overflowIndex < 0) ? 0 : (1ULL << overflowIndex);
    const uint64_t signMask = 1ULL << (length - 1), carryMask = 1ULL << carryIndex, inOutMask = lengthMask << start;
    REAL* out = FN(alu_new)(nq);
    for (uint64_t lcv = 0; lcv < n; ++lcv) {
        if (lcv & carryMask) {
            continue;
        }
        const uint64_t otherRes = lcv & ~(inOutMask | carryMask);
        const uint64_t inOutInt = (lcv & inOutMask) >> start;
        const uint64_t outInt = inOutInt + toMod;
        const uint64_t outRes = (outInt < lengthPower) ? ((outInt << start) | otherRes)
                                                       : (((outInt - lengthPower) << start) | otherRes | carryMask);
        int flip = FN(alu_is_overflow_add)(inOutInt, toMod, signMask, lengthPower);
        if (overflowIndex >= 0) {
            flip = flip && ((outRes & overflowMask) == overflowMask);
        }
        if (flip) {
            ALU_MOVE(outRes, lcv, -1);
        } else {
            ALU_MOVE(outRes, lcv, 1);
        }
    }
    FN(alu_adopt)(psi, out, nq);
}

/* QEngineCPU::MULDIV :422-456 / CMULDIV :488-551.  inverse == 0: out[mulRes] = in[orig] (MUL); 1: out[orig] = in[mulRes]
 * (DIV).  Sources are the indices whose carry register (and, iterated by par_for_mask, control qubits) read zero; with
 * controls, every partial control pattern of such an index is copied unchanged (:536-547). */
void FN(orc_muldiv)(REAL* psi, int nq, int inverse, uint64_t toMul, int start, int carryStart, int length, uint64_t controlMask)
{
    const uint64_t n = 1ULL << nq, lowMask = (1ULL << length) - 1U, highMask = lowMask << length;
    const uint64_t inOutMask = lowMask << start, carryMask = lowMask << carryStart;
    const uint64_t otherMask = (n - 1U) ^ (inOutMask | carryMask | controlMask);
    REAL* out = FN(alu_new)(nq);
    for (uint64_t lcv = 0; lcv < n; ++lcv) {
        if (lcv & (carryMask | controlMask)) {
            continue; /* par_for_skip / par_for_mask domain */
        }
        const uint64_t otherRes = lcv & otherMask;
        const uint64_t mulInt = ((lcv & inOutMask) >> start) * toMul;
        const uint64_t mulRes =
            ((mulInt & lowMask) << start) | (((mulInt & highMask) >> length) << carryStart) | otherRes | controlMask;
        const uint64_t origRes = lcv | controlMask;
        if (inverse) {
            ALU_MOVE(origRes, mulRes, 1);
        } else {
            ALU_MOVE(mulRes, origRes, 1);
        }
        if (controlMask) {
            /* all proper subsets of the control mask, including the empty one */
            for (uint64_t sub = (controlMask - 1U) & controlMask;; sub = (sub - 1U) & controlMask) {
                ALU_MOVE(lcv | sub, lcv | sub, 1);
                if (!sub) {
                    break;
                }
            }
        }
    }
    FN(alu_adopt)(psi, out, nq);
}

/* QEngineCPU::ModNOut :595-632 / CModNOut :670-735.  kind 0: in*toMod, 1: the same, inverse direction, 2: toMod^in */
void FN(orc_modnout)(REAL* psi, int nq, int kind, uint64_t toMod, uint64_t modN, int inStart, int outStart, int length,
    uint64_t controlMask)
{
    const uint64_t n = 1ULL << nq, lowMask = (1ULL << length) - 1U;
    const uint64_t inMask = lowMask << inStart, skipMask = (lowMask << outStart) & (n - 1U);
    REAL* out = FN(alu_new)(nq);
    for (uint64_t lcv = 0; lcv < n; ++lcv) {
        if (lcv & (skipMask | controlMask)) {
            continue;
        }
        const uint64_t inRes = lcv & inMask;
        const uint64_t inInt = inRes >> inStart;
        const uint64_t k = (kind == 2) ? FN(alu_int_pow)(toMod, inInt) : (inInt * toMod);
        const uint64_t outRes = (k % modN) << outStart;
        /* lcv has a zero output register, so inRes | outRes | otherRes == lcv | outRes */
        if (kind == 1) {
            ALU_MOVE(lcv | controlMask, lcv | outRes | controlMask, 1);
        } else {
            ALU_MOVE(lcv | outRes | controlMask, lcv | controlMask, 1);
        }
        if (controlMask) {
            for (uint64_t sub = (controlMask - 1U) & controlMask;; sub = (sub - 1U) & controlMask) {
                ALU_MOVE(lcv | sub, lcv | sub, 1);
                if (!sub) {
                    break;
                }
            }
        }
    }
    FN(alu_adopt)(psi, out, nq);
}

/* QEngineCPU::IndexedLDA :983-1083 (kind 0), IndexedADC :1086-1260 (1), IndexedSBC :1263-1444 (2); dense branches.
 * The M/X of the carry qubit and LDA's SetReg stay with the caller; carryIn is what the caller measured. */
void FN(orc_indexed)(REAL* psi, int nq, int kind, int indexStart, int indexLength, int valueStart, int valueLength, int carryIndex,
    int carryIn, const unsigned char* values)
{
    const uint64_t n = 1ULL << nq;
    const int valueBytes = (valueLength + 7) >> 3;
    const uint64_t lengthPower = 1ULL << valueLength;
    const uint64_t inputMask = ((1ULL << indexLength) - 1U) << indexStart;
    const uint64_t outputMask = (lengthPower - 1U) << valueStart;
    REAL* out = FN(alu_new)(nq);
    if (kind == 0) {
        for (uint64_t lcv = 0; lcv < n; ++lcv) {
            if (lcv & outputMask) {
                continue; /* par_for_skip(valueStart, valueLength) */
            }
            const uint64_t v = FN(alu_value)(values, (lcv & inputMask) >> indexStart, valueBytes);
            ALU_MOVE(lcv | (v << valueStart), lcv, 1);
        }
    } else {
        const uint64_t carryMask = 1ULL << carryIndex;
        /* ADC skips 1 bit at the carry qubit (:1251); SBC skips valueLength bits from there (:1436) */
        const uint64_t skipMask = (kind == 1) ? carryMask : ((((1ULL << valueLength) - 1U) << carryIndex) & (n - 1U));
        const uint64_t otherMask = (n - 1U) & ~(inputMask | outputMask | carryMask);
        for (uint64_t lcv = 0; lcv < n; ++lcv) {
            if (lcv & skipMask) {
                continue;
            }
            const uint64_t otherRes = lcv & otherMask, inputRes = lcv & inputMask;
            uint64_t outputInt = FN(alu_value)(values, inputRes >> indexStart, valueBytes);
            const uint64_t cur = (lcv & outputMask) >> valueStart;
            if (kind == 1) {
                outputInt += cur + (uint64_t)carryIn;
            } else {
                outputInt = cur + (lengthPower - (outputInt + (uint64_t)carryIn));
            }
            uint64_t carryRes = 0;
            if (outputInt >= lengthPower) {
                outputInt -= lengthPower;
                carryRes = carryMask;
            }
            ALU_MOVE((outputInt << valueStart) | inputRes | otherRes | carryRes, lcv, 1);
        }
    }
    FN(alu_adopt)(psi, out, nq);
}

/* QEngineCPU::Hash :1447-1506 */
void FN(orc_hash)(REAL* psi, int nq, int start, int length, const unsigned char* values)
{
    const uint64_t n = 1ULL << nq;
    const int bytes = (length + 7) >> 3;
    const uint64_t inputMask = ((1ULL << length) - 1U) << start;
    REAL* out = FN(alu_new)(nq);
    for (uint64_t lcv = 0; lcv < n; ++lcv) {
        const uint64_t inputRes = lcv & inputMask;
        const uint64_t v = FN(alu_value)(values, inputRes >> start, bytes);
        ALU_MOVE((v << start) | (lcv & ~inputRes), lcv, 1);
    }
    FN(alu_adopt)(psi, out, nq);
}

/* QEngineCPU::PhaseFlipIfLess :1703-1720 (flagIndex < 0) / CPhaseFlipIfLess :1678-1701 */
void FN(orc_phase_flip_if_less)(REAL* psi, int nq, uint64_t greaterPerm, int start, int length, int flagIndex)
{
    const uint64_t n = 1ULL << nq, regMask = ((1ULL << length) - 1U) << start;
    const uint64_t flagMask = (flagIndex < 0) ? 0 : (1ULL << flagIndex);
    for (uint64_t lcv = 0; lcv < n; ++lcv) {
        if ((((lcv & regMask) >> start) < greaterPerm) && ((lcv & flagMask) == flagMask)) {
            psi[2 * lcv] = -psi[2 * lcv];
            psi[2 * lcv + 1] = -psi[2 * lcv + 1];
        }
    }
}
#undef ALU_MOVE
#undef FN
#undef CAT
#undef CAT_
