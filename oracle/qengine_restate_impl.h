/* qengine_restate_impl.h — TEST INFRASTRUCTURE (oracle), NOT product code.
 *
 * Plain-C, single-threaded restatement of the reference's QEngineCPU hot path
 * (unitaryfoundation/qrack, /root/reference/src/qengine/state.cpp), included once per precision by
 * qengine_restate.c with REAL / SUF defined.  Every function cites the reference lines it follows.
 * Amplitudes are interleaved (re,im) REALs (reference include/statevector.hpp:94).
 * PARITY PIN: checked against the compiled reference itself (oracle/_ref/ref_harness_f{32,64}) by
 * tests/test_oracle_pin.py and against the committed fixtures in tests/golden/.
 */

#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, SUF)

/* "insert zero bits" index map: reference src/common/parallel_for.cpp:118-149 (par_for_mask) */
static uint64_t FN(push_apart)(uint64_t i, int n, const uint64_t* pows)
{
    for (int m = 0; m < n; ++m) {
        const uint64_t low = pows[m] - 1U;
        const uint64_t high = ~(low + pows[m]);
        i = ((i << 1U) & high) | (i & low);
    }
    return i;
}

static int FN(is_norm_0)(REAL re, REAL im) { return (re * re + im * im) <= FP_NORM_EPS; } /* qrack_types.hpp:28 */

/* QEngineCPU::Apply2x2 — reference src/qengine/state.cpp:392-533 (dense path).
 * nrm is applied only when the caller decided doApplyNorm (state.cpp:422-425); normOut may be NULL (doCalcNorm false).
 * The three matrix classes (diagonal / anti-diagonal / general, :450-468) are selected with IS_NORM_0 as there. */
void FN(orc_apply2x2)(REAL* psi, int nq, uint64_t off1, uint64_t off2, const REAL* m, int bitCount, const uint64_t* pows,
    REAL nrm, REAL norm_thresh, double* normOut)
{
    const uint64_t items = (1ULL << nq) >> bitCount;
    const int diag = FN(is_norm_0)(m[2], m[3]) && FN(is_norm_0)(m[4], m[5]);
    const int anti = FN(is_norm_0)(m[0], m[1]) && FN(is_norm_0)(m[6], m[7]);
    REAL acc = 0;
    for (uint64_t lcv = 0; lcv < items; ++lcv) {
        const uint64_t i = FN(push_apart)(lcv, bitCount, pows);
        REAL* pa = psi + 2U * (i + off1);
        REAL* pb = psi + 2U * (i + off2);
        const REAL ar = pa[0], ai = pa[1], br = pb[0], bi = pb[1];
        REAL xr, xi, yr, yi;
        if (diag) { /* mtrxPhase * qubit, :451-455 */
            xr = m[0] * ar - m[1] * ai;
            xi = m[0] * ai + m[1] * ar;
            yr = m[6] * br - m[7] * bi;
            yi = m[6] * bi + m[7] * br;
        } else if (anti) { /* read2(offset2, offset1), :456-461 */
            xr = m[2] * br - m[3] * bi;
            xi = m[2] * bi + m[3] * br;
            yr = m[4] * ar - m[5] * ai;
            yi = m[4] * ai + m[5] * ar;
        } else { /* matrixMul, :462-467 */
            xr = (m[0] * ar - m[1] * ai) + (m[2] * br - m[3] * bi);
            xi = (m[0] * ai + m[1] * ar) + (m[2] * bi + m[3] * br);
            yr = (m[4] * ar - m[5] * ai) + (m[6] * br - m[7] * bi);
            yi = (m[4] * ai + m[5] * ar) + (m[6] * bi + m[7] * br);
        }
        xr *= nrm;
        xi *= nrm;
        yr *= nrm;
        yi *= nrm;
        if (normOut) { /* NORM_THRESH_KERNEL :361-382 / NORM_CALC_KERNEL :384-390 */
            REAL d = xr * xr + xi * xi;
            if (d < norm_thresh) {
                xr = 0;
                xi = 0;
            } else {
                acc += d;
            }
            d = yr * yr + yi * yi;
            if (d < norm_thresh) {
                yr = 0;
                yi = 0;
            } else {
                acc += d;
            }
        }
        pa[0] = xr;
        pa[1] = xi;
        pb[0] = yr;
        pb[1] = yi;
    }
    if (normOut) {
        *normOut = (double)acc;
    }
}

/* QEngineCPU::ApplyM — state.cpp:2167-2196 */
void FN(orc_apply_m)(REAL* psi, int nq, uint64_t mask, uint64_t result, REAL nre, REAL nim)
{
    const uint64_t n = 1ULL << nq;
    for (uint64_t i = 0; i < n; ++i) {
        if ((i & mask) == result) {
            const REAL r = psi[2 * i], im = psi[2 * i + 1];
            psi[2 * i] = nre * r - nim * im;
            psi[2 * i + 1] = nre * im + nim * r;
        } else {
            psi[2 * i] = 0;
            psi[2 * i + 1] = 0;
        }
    }
}

/* Prob / ProbReg / ProbMask — state.cpp:1751-1947: sum |psi[lcv|perm]|^2 over the complement of mask */
double FN(orc_prob_mask)(const REAL* psi, int nq, uint64_t mask, uint64_t perm)
{
    uint64_t pows[64];
    int np = 0;
    for (uint64_t v = mask; v;) { /* :1925-1931 */
        const uint64_t old = v;
        v &= v - 1U;
        pows[np++] = (v ^ old) & old;
    }
    const uint64_t items = (1ULL << nq) >> np;
    REAL acc = 0;
    for (uint64_t lcv = 0; lcv < items; ++lcv) {
        const uint64_t i = FN(push_apart)(lcv, np, pows) | perm;
        acc += psi[2 * i] * psi[2 * i] + psi[2 * i + 1] * psi[2 * i + 1];
    }
    return (double)acc;
}

/* ProbParity — state.cpp:1949-1993 */
double FN(orc_prob_parity)(const REAL* psi, int nq, uint64_t mask)
{
    const uint64_t n = 1ULL << nq;
    REAL acc = 0;
    for (uint64_t i = 0; i < n; ++i) {
        if (__builtin_popcountll(i & mask) & 1) {
            acc += psi[2 * i] * psi[2 * i] + psi[2 * i + 1] * psi[2 * i + 1];
        }
    }
    return (double)acc;
}

/* ForceMParity collapse — state.cpp:2083-2091 */
double FN(orc_collapse_parity)(REAL* psi, int nq, uint64_t mask, int result)
{
    const uint64_t n = 1ULL << nq;
    REAL acc = 0;
    for (uint64_t i = 0; i < n; ++i) {
        if ((__builtin_popcountll(i & mask) & 1) == result) {
            acc += psi[2 * i] * psi[2 * i] + psi[2 * i + 1] * psi[2 * i + 1];
        } else {
            psi[2 * i] = 0;
            psi[2 * i + 1] = 0;
        }
    }
    return (double)acc;
}

/* UpdateRunningNorm -> par_norm — state.cpp:2250-2268, parallel_for.cpp:244-300 */
double FN(orc_norm)(const REAL* psi, int nq, REAL thresh)
{
    const uint64_t n = 1ULL << nq;
    REAL acc = 0;
    for (uint64_t i = 0; i < n; ++i) {
        const REAL v = psi[2 * i] * psi[2 * i] + psi[2 * i + 1] * psi[2 * i + 1];
        if (v >= thresh) {
            acc += v;
        }
    }
    return (double)acc;
}

/* NormalizeState — state.cpp:2198-2248 (nrm already resolved by the caller; thresh <= 0 means no floor) */
void FN(orc_normalize)(REAL* psi, int nq, REAL nrm, REAL thresh, REAL phaseArg)
{
    const uint64_t n = 1ULL << nq;
    const REAL f = (REAL)1 / (REAL)sqrt((double)nrm);
    const REAL cr = f * (REAL)cos((double)phaseArg), ci = f * (REAL)sin((double)phaseArg);
    for (uint64_t i = 0; i < n; ++i) {
        REAL r = psi[2 * i], im = psi[2 * i + 1];
        if (thresh > 0 && (r * r + im * im) < thresh) {
            r = 0;
            im = 0;
        }
        psi[2 * i] = cr * r - ci * im;
        psi[2 * i + 1] = cr * im + ci * r;
    }
}

/* XMask — state.cpp:965-1007 */
void FN(orc_xmask)(REAL* psi, int nq, uint64_t mask)
{
    const uint64_t n = 1ULL << nq;
    const uint64_t otherMask = (n - 1U) ^ mask;
    for (uint64_t lcv = 0; lcv < n; ++lcv) {
        const uint64_t otherRes = lcv & otherMask;
        uint64_t setInt = lcv & mask;
        uint64_t resetInt = setInt ^ mask;
        if (setInt < resetInt) {
            continue;
        }
        setInt |= otherRes;
        resetInt |= otherRes;
        const REAL r = psi[2 * resetInt], im = psi[2 * resetInt + 1];
        psi[2 * resetInt] = psi[2 * setInt];
        psi[2 * resetInt + 1] = psi[2 * setInt + 1];
        psi[2 * setInt] = r;
        psi[2 * setInt + 1] = im;
    }
}

/* PhaseParity — state.cpp:1009-1054; UniformParityRZ / CUniformParityRZ — :1200-1264 (cmask == 0: uncontrolled).
 * odd-parity amplitudes get (cs + i sn), even-parity ones (cs - i sn). */
void FN(orc_phase_parity)(REAL* psi, int nq, uint64_t cmask, uint64_t mask, REAL cs, REAL sn)
{
    const uint64_t n = 1ULL << nq;
    for (uint64_t i = 0; i < n; ++i) {
        if ((i & cmask) != cmask) {
            continue;
        }
        const REAL s = (__builtin_popcountll(i & mask) & 1) ? sn : -sn;
        const REAL r = psi[2 * i], im = psi[2 * i + 1];
        psi[2 * i] = cs * r - s * im;
        psi[2 * i + 1] = cs * im + s * r;
    }
}

/* PhaseRootNMask — state.cpp:1056-1092 */
void FN(orc_phase_root_n_mask)(REAL* psi, int nq, int nroot, uint64_t mask)
{
    const uint64_t n = 1ULL << nq;
    const uint64_t nPhases = 1ULL << nroot;
    const REAL radians = -(REAL)M_PI / (REAL)(1ULL << (nroot - 1));
    for (uint64_t i = 0; i < n; ++i) {
        const uint64_t steps = (uint64_t)__builtin_popcountll(i & mask) % nPhases;
        if (steps) {
            const REAL a = radians * (REAL)steps;
            const REAL cs = (REAL)cos((double)a), sn = (REAL)sin((double)a);
            const REAL r = psi[2 * i], im = psi[2 * i + 1];
            psi[2 * i] = cs * r - sn * im;
            psi[2 * i + 1] = cs * im + sn * r;
        }
    }
}

/* Compose(toCopy, start) — state.cpp:1368-1459 (start == nA gives the append form :1271-1362) */
void FN(orc_compose)(REAL* out, const REAL* a, int nA, const REAL* b, int nB, int start)
{
    const uint64_t n = 1ULL << (nA + nB);
    const uint64_t startMask = (1ULL << start) - 1U;
    const uint64_t midMask = ((1ULL << nB) - 1U) << start;
    const uint64_t endMask = (n - 1U) & ~(startMask | midMask);
    for (uint64_t l = 0; l < n; ++l) {
        const uint64_t ia = (l & startMask) | ((l & endMask) >> nB);
        const uint64_t ib = (l & midMask) >> start;
        const REAL ar = a[2 * ia], ai = a[2 * ia + 1], br = b[2 * ib], bi = b[2 * ib + 1];
        out[2 * l] = ar * br - ai * bi;
        out[2 * l + 1] = ar * bi + ai * br;
    }
}

/* DecomposeDispose — state.cpp:1551-1696.  rem/part receive the rebuilt factors (part may be NULL = Dispose). */
void FN(orc_decompose)(const REAL* psi, int nq, int start, int length, REAL* rem, REAL* part, REAL floorv)
{
    const int nl = nq - length;
    const uint64_t partPower = 1ULL << length, remPower = 1ULL << nl;
    const uint64_t startMask = (1ULL << start) - 1U;
    REAL* remProb = (REAL*)calloc(remPower, sizeof(REAL));
    REAL* remAngle = (REAL*)calloc(remPower, sizeof(REAL));
    REAL* partProb = (REAL*)calloc(partPower, sizeof(REAL));
    REAL* partAngle = (REAL*)calloc(partPower, sizeof(REAL));
    for (uint64_t lcv = 0; lcv < remPower; ++lcv) { /* :1605-1617 */
        uint64_t j = lcv & startMask;
        j |= (lcv ^ j) << length;
        for (uint64_t k = 0; k < partPower; ++k) {
            const uint64_t i = j | (k << start);
            const REAL nrm = psi[2 * i] * psi[2 * i] + psi[2 * i + 1] * psi[2 * i + 1];
            remProb[lcv] += nrm;
            if (nrm > floorv) {
                partAngle[k] += (REAL)atan2((double)psi[2 * i + 1], (double)psi[2 * i]) * nrm;
            }
        }
    }
    for (uint64_t lcv = 0; lcv < partPower; ++lcv) { /* :1618-1638 */
        const uint64_t j = lcv << start;
        for (uint64_t k = 0; k < remPower; ++k) {
            uint64_t l = k & startMask;
            l |= j | ((k ^ l) << length);
            const REAL nrm = psi[2 * l] * psi[2 * l] + psi[2 * l + 1] * psi[2 * l + 1];
            partProb[lcv] += nrm;
            if (nrm > floorv) {
                remAngle[k] += (REAL)atan2((double)psi[2 * l + 1], (double)psi[2 * l]) * nrm;
            }
        }
        if (partProb[lcv] > floorv) {
            partAngle[lcv] /= partProb[lcv];
        }
    }
    for (uint64_t lcv = 0; lcv < remPower; ++lcv) { /* :1639-1644 */
        if (remProb[lcv] > floorv) {
            remAngle[lcv] /= remProb[lcv];
        }
    }
    if (part) { /* :1680-1683 */
        for (uint64_t k = 0; k < partPower; ++k) {
            const REAL mag = (REAL)sqrt((double)partProb[k]);
            part[2 * k] = mag * (REAL)cos((double)partAngle[k]);
            part[2 * k + 1] = mag * (REAL)sin((double)partAngle[k]);
        }
    }
    for (uint64_t r = 0; r < remPower; ++r) { /* :1693-1695 */
        const REAL mag = (REAL)sqrt((double)remProb[r]);
        rem[2 * r] = mag * (REAL)cos((double)remAngle[r]);
        rem[2 * r + 1] = mag * (REAL)sin((double)remAngle[r]);
    }
    free(remProb);
    free(remAngle);
    free(partProb);
    free(partAngle);
}

/* Dispose(start, length, perm) — state.cpp:1708-1748 */
void FN(orc_dispose_perm)(REAL* out, const REAL* psi, int nq, int start, int length, uint64_t perm)
{
    const uint64_t remPower = 1ULL << (nq - length);
    const uint64_t skipMask = (1ULL << start) - 1U;
    const uint64_t disposedRes = perm << start;
    for (uint64_t h = 0; h < remPower; ++h) {
        const uint64_t lo = h & skipMask;
        const uint64_t i = lo | ((h ^ lo) << length) | disposedRes;
        out[2 * h] = psi[2 * i];
        out[2 * h + 1] = psi[2 * i + 1];
    }
}

/* SumSqrDiff inner product — state.cpp:2152-2164: sum conj(a)*b */
void FN(orc_inner)(const REAL* a, const REAL* b, int nq, double* re, double* im)
{
    const uint64_t n = 1ULL << nq;
    REAL r = 0, i_ = 0;
    for (uint64_t i = 0; i < n; ++i) {
        r += a[2 * i] * b[2 * i] + a[2 * i + 1] * b[2 * i + 1];
        i_ += a[2 * i] * b[2 * i + 1] - a[2 * i + 1] * b[2 * i];
    }
    *re = (double)r;
    *im = (double)i_;
}

/* ShuffleBuffers — state.cpp:134-163, statevector.hpp:229-237: swap a[half..] with b[..half] */
void FN(orc_shuffle)(REAL* a, REAL* b, int nq)
{
    const uint64_t half = (1ULL << nq) >> 1;
    for (uint64_t i = 0; i < 2 * half; ++i) {
        const REAL t = a[2 * half + i];
        a[2 * half + i] = b[i];
        b[i] = t;
    }
}

/* UniformlyControlledSingleBit — state.cpp:1094-1198 */
void FN(orc_uniformly_controlled)(REAL* psi, int nq, int nc, const int* controls, int target, const REAL* mtrxs, int nskip,
    const uint64_t* skipPowers, uint64_t skipValueMask, REAL nrm)
{
    const uint64_t tpow = 1ULL << target;
    const uint64_t half = (1ULL << nq) >> 1;
    for (uint64_t j = 0; j < half; ++j) {
        const uint64_t lo = j & (tpow - 1U);
        const uint64_t lcv = ((j ^ lo) << 1) | lo; /* par_for_skip(0, max, targetPower, 1) :1192 */
        uint64_t offset = 0;
        for (int c = 0; c < nc; ++c) {
            if (lcv & (1ULL << controls[c])) {
                offset |= 1ULL << c;
            }
        }
        uint64_t i = 0, iHigh = offset;
        for (int p = 0; p < nskip; ++p) {
            const uint64_t iLow = iHigh & (skipPowers[p] - 1U);
            i |= iLow;
            iHigh = (iHigh ^ iLow) << 1U;
        }
        i |= iHigh;
        const REAL* m = mtrxs + 8U * (i | skipValueMask);
        REAL* pa = psi + 2 * lcv;
        REAL* pb = psi + 2 * (lcv | tpow);
        const REAL ar = pa[0], ai = pa[1], br = pb[0], bi = pb[1];
        pa[0] = nrm * ((m[0] * ar - m[1] * ai) + (m[2] * br - m[3] * bi));
        pa[1] = nrm * ((m[0] * ai + m[1] * ar) + (m[2] * bi + m[3] * br));
        pb[0] = nrm * ((m[4] * ar - m[5] * ai) + (m[6] * br - m[7] * bi));
        pb[1] = nrm * ((m[4] * ai + m[5] * ar) + (m[6] * bi + m[7] * br));
    }
}

/* GetExpectation — utility.cpp */
double FN(orc_expectation)(const REAL* psi, int nq, int start, int length)
{
    const uint64_t n = 1ULL << nq;
    const uint64_t lm = (1ULL << length) - 1U;
    double acc = 0;
    for (uint64_t i = 0; i < n; ++i) {
        acc += (double)(psi[2 * i] * psi[2 * i] + psi[2 * i + 1] * psi[2 * i + 1]) * (double)((i >> start) & lm);
    }
    return acc;
}

#undef FN
#undef CAT
#undef CAT_
