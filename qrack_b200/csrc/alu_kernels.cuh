// alu_kernels.cuh — basis-state maps of the QAlu family as ONE out-of-place sweep.
//
// Every member of the reference's QEngineCPU arithmetic (src/qengine/arithmetic.cpp) is a map over computational basis
// indices: out[f(i)] = ±in[i] (or, for the inverse forms, out[i] = in[f(i)]), restricted to a domain where some "carry"
// or "output" register reads zero, and the identity where the control qubits are not all set.  The map is evaluated per
// amplitude by k_alu_map from a small descriptor; the sweep reads each source amplitude once and writes each destination
// once, i.e. it is bound by HBM bandwidth (2 x 2^n x sizeof(amplitude) algorithmic bytes), like a single gate.
// Included by b200sv.cu (same translation unit as the other kernels).
#pragma once

namespace b200sv {

enum AluKind {
    ALU_ROL = 0,    // arithmetic.cpp:23-70
    ALU_INC,        // :73-118   (and CINC :121-172 through ctrlMask)
    ALU_INCC,       // :175-224  INCDECC
    ALU_INCS,       // :227-309
    ALU_INCSC,      // :312-361  INCDECSC, carry only
    ALU_INCSC_OVF,  // :364-419  INCDECSC, overflow flag + carry
    ALU_MUL,        // :422-471  MULDIV forward (and CMULDIV :488-551)
    ALU_DIV,        //           MULDIV inverse
    ALU_MULMODN,    // :595-655  ModNOut, in*toMod % modN
    ALU_IMULMODN,   //           inverse of the above
    ALU_POWMODN,    // :658-667  ModNOut, toMod^in % modN
    ALU_LDA,        // :983-1083 IndexedLDA
    ALU_ADC,        // :1086-1260 IndexedADC
    ALU_SBC,        // :1263-1444 IndexedSBC
    ALU_HASH,       // :1447-1506
};

struct AluDesc {
    int kind;
    int start, length;   // in/out (or index / input) register
    int start2, length2; // carry / value / output register
    uint64_t arg;        // toAdd, toMul, base, shift
    uint64_t modN;
    uint64_t carryMask;    // single carry qubit (0: none)
    uint64_t overflowMask; // single overflow qubit (0: none)
    uint64_t ctrlMask;     // all of these must be set for the map to act
    uint64_t zeroMask;     // sources with any of these bits set are dropped (destination stays zero)
    uint64_t carryIn;      // ADC / SBC
    int valueBytes;        // bytes per table entry (LDA / ADC / SBC / HASH)
    int gather;            // 0: out[f(i)] = in[i]; 1: out[i] = in[f(i)]
    const unsigned char* table; // device copy of the classical table
};

__device__ __forceinline__ uint64_t alu_pow_wrap(uint64_t base, uint64_t power)
{
    // square-and-multiply with 64-bit wrap-around, as intPowOcl (src/common/functions.cpp:77-95)
    uint64_t r = 1U;
    while (power) {
        if (power & 1U) {
            r *= base;
        }
        base *= base;
        power >>= 1U;
    }
    return r;
}

// signed-addition overflow test of the reference (src/common/functions.cpp:214-233)
__device__ __forceinline__ bool alu_overflow_add(uint64_t a, uint64_t b, uint64_t signMask, uint64_t lengthPower)
{
    if (a & b & signMask) {
        a = ((~a) & (lengthPower - 1U)) + 1U;
        b = ((~b) & (lengthPower - 1U)) + 1U;
        return (a + b) > signMask;
    }
    if ((~a) & (~b) & signMask) {
        return (a + b) >= signMask;
    }
    return false;
}

__device__ __forceinline__ uint64_t alu_table(const AluDesc& d, uint64_t idx)
{
    uint64_t v = 0;
    for (int j = 0; j < d.valueBytes; ++j) {
        v |= (uint64_t)d.table[idx * (uint64_t)d.valueBytes + (uint64_t)j] << (8U * j);
    }
    return v;
}

// f(i) and the sign of the mapped amplitude
__device__ __forceinline__ uint64_t alu_map(const AluDesc& d, uint64_t i, bool& neg)
{
    neg = false;
    const uint64_t lenPow = 1ULL << d.length;
    const uint64_t lenMask = lenPow - 1U;
    const uint64_t regMask = lenMask << d.start;
    const uint64_t reg = (i & regMask) >> d.start;
    switch (d.kind) {
    case ALU_ROL: {
        const uint64_t o = (reg >> (d.length - (int)d.arg)) | ((reg << d.arg) & lenMask);
        return (i & ~regMask) | (o << d.start);
    }
    case ALU_INC: {
        const uint64_t o = (reg + d.arg) & lenMask;
        return (i & ~regMask) | (o << d.start);
    }
    case ALU_INCC:
    case ALU_INCS:
    case ALU_INCSC:
    case ALU_INCSC_OVF: {
        uint64_t o = reg + d.arg;
        uint64_t res = i & ~(regMask | d.carryMask);
        if (o >= lenPow) {
            o -= lenPow;
            res |= d.carryMask; // 0 for INCS
        }
        res |= o << d.start;
        if (d.kind != ALU_INCC) {
            const bool ovf = alu_overflow_add(reg, d.arg, lenPow >> 1U, lenPow);
            neg = ovf && ((d.kind == ALU_INCSC) || ((res & d.overflowMask) == d.overflowMask));
        }
        return res;
    }
    case ALU_MUL:
    case ALU_DIV: {
        const uint64_t carryRegMask = lenMask << d.start2;
        const uint64_t prod = reg * d.arg;
        return (i & ~(regMask | carryRegMask)) | ((prod & lenMask) << d.start) | (((prod >> d.length) & lenMask) << d.start2);
    }
    case ALU_MULMODN:
    case ALU_IMULMODN:
    case ALU_POWMODN: {
        const uint64_t k = (d.kind == ALU_POWMODN) ? alu_pow_wrap(d.arg, reg) : reg * d.arg;
        return i | ((k % d.modN) << d.start2);
    }
    case ALU_LDA:
        return i | (alu_table(d, reg) << d.start2);
    case ALU_ADC:
    case ALU_SBC: {
        const uint64_t vPow = 1ULL << d.length2;
        const uint64_t vMask = (vPow - 1U) << d.start2;
        const uint64_t cur = (i & vMask) >> d.start2;
        const uint64_t t = alu_table(d, reg);
        uint64_t o = (d.kind == ALU_ADC) ? (t + cur + d.carryIn) : (cur + (vPow - (t + d.carryIn)));
        uint64_t res = i & ~(vMask | d.carryMask);
        if (o >= vPow) {
            o -= vPow;
            res |= d.carryMask;
        }
        return res | (o << d.start2);
    }
    case ALU_HASH:
        return (i & ~regMask) | (alu_table(d, reg) << d.start);
    default:
        return i;
    }
}

template <typename C> __device__ __forceinline__ C alu_neg(C v)
{
    v.x = -v.x;
    v.y = -v.y;
    return v;
}

// one amplitude per thread-iteration; reads coalesced, writes follow the map (contiguous runs for the adders)
template <typename C>
__global__ void __launch_bounds__(256) k_alu_map(const C* __restrict__ in, C* __restrict__ out, uint64_t n, const AluDesc d)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        if (i & d.zeroMask) {
            continue;
        }
        if ((i & d.ctrlMask) != d.ctrlMask) {
            out[i] = in[i];
            continue;
        }
        bool neg;
        const uint64_t j = alu_map(d, i, neg);
        if (d.gather) {
            out[i] = in[j];
        } else {
            const C v = in[i];
            out[j] = neg ? alu_neg(v) : v;
        }
    }
}

// (C)PhaseFlipIfLess: in place (arithmetic.cpp:1678-1720)
template <typename C>
__global__ void __launch_bounds__(256)
    k_phase_flip_if_less(C* __restrict__ psi, uint64_t n, uint64_t regMask, int start, uint64_t greaterPerm, uint64_t flagMask)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        if ((((i & regMask) >> start) < greaterPerm) && ((i & flagMask) == flagMask)) {
            psi[i] = alu_neg(psi[i]);
        }
    }
}

} // namespace b200sv
