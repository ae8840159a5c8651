// fused.cu — the fused multi-gate sweep: a window of queued single-/controlled-single-qubit gates is applied in ONE
// pass over HBM (read 2^n amplitudes once, write once), instead of one read+write sweep per gate as in the reference
// (src/qengine/state.cpp:392-533 does one par_for_mask sweep per Apply2x2; src/qengine/cuda.cu:857-1059 one launch).
//
// Shape of one sweep:
//   * tile  = 2^KC 16-byte chunks (64 KB) living in shared memory.  Its index bits are the low L qubits (so every
//     global access is a >= 2^L-amplitude contiguous run) plus up to H arbitrary "high" qubits chosen by the scheduler.
//   * load  : coalesced 128-bit loads, stored to smem in the TMA SWIZZLE_128B layout (16-B chunk index ^= row index),
//             so the in-tile butterflies are bank-conflict free for any target bit.
//   * passes: each pass picks RB "register" chunk bits; every thread pulls a 2^RB-chunk sub-block into registers
//             (fp32: +qubit 0 inside the chunk), applies every queued gate whose target is one of those bits (controls
//             anywhere: tile-local bits become a per-amplitude predicate, outside bits a per-tile predicate) and
//             writes the sub-block back.  Diagonal gates are index-only and ride along in any pass, on any qubit.
//   * store : swizzled smem -> coalesced 128-bit streaming stores.
// The host-side scheduler below reorders only gates that commute (disjoint qubits, or shared qubits used diagonally by
// both) and never changes the product of the gate sequence.
#include "sv_common.cuh"

#include <algorithm>
#include <cstring>
#include <mutex>

namespace b200sv {

constexpr int MAX_HIGH = 8;
constexpr int MAX_PASS = 12;
constexpr int MAX_OPS = 64;
constexpr int MAX_NA = 32; // register amplitudes per sub-block

enum { OP_GENERAL = 0, OP_REALM = 1, OP_XSWAP = 2, OP_PHASE = 3 };

template <typename R> struct DevOp {
    uint64_t omask, oval; // predicate on the tile's global base index (qubits outside the tile)
    uint32_t lmaskSb, lvalSb; // predicate on the sub-block base (tile-local amplitude index bits outside the register set)
    uint32_t emask;           // bit e set: register amplitude e satisfies the register-resident part of the controls
    int kind;
    int jr; // register-bit index of the target (OP_PHASE: unused)
    int pad;
    R m[8];
};

struct DevPass {
    int opBegin, opEnd;
    int nsb;                 // number of sub-block index bits
    unsigned char sbit[16];  // sub-block index bit i -> tile chunk bit
    unsigned short roffc[MAX_NA]; // register chunk e -> tile chunk offset
};

struct DevSweep {
    int nHigh;        // high qubits in the tile
    int lowAmpBits;   // L: low qubits in the tile
    int kc;           // tile chunk bits actually used (<= KC)
    int nPass;
    int nOps;
    int pad[3];
    uint64_t highLow[MAX_HIGH]; // (2^q - 1) for push_apart of the tile base, ascending
    uint64_t highPow[MAX_HIGH]; // 2^q
    DevPass pass[MAX_PASS];
};

template <typename R> struct Apc {
    static constexpr int v = 16 / (2 * sizeof(R)); // amplitudes per 16-byte chunk: fp32 2, fp64 1
};

__device__ __forceinline__ uint4 ld_stream(const uint4* p)
{
    uint4 v;
    asm volatile("ld.global.cs.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_stream(uint4* p, const uint4 v)
{
    asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// TMA SWIZZLE_128B layout on 16-byte chunks: chunk index bits [0..2] ^= bits [3..5]
__device__ __forceinline__ uint32_t swz(uint32_t c) { return c ^ ((c >> 3) & 7U); }

template <int JR, int NA, typename C, typename R>
__device__ __forceinline__ void app_general(C (&a)[NA], const R* __restrict__ m, uint32_t em)
{
    const C m0 = mk<R>(m[0], m[1]), m1 = mk<R>(m[2], m[3]), m2 = mk<R>(m[4], m[5]), m3 = mk<R>(m[6], m[7]);
#pragma unroll
    for (int e = 0; e < NA; ++e) {
        if (!(e & (1 << JR))) {
            if (em & (1U << e)) {
                const C x = a[e], y = a[e | (1 << JR)];
                a[e] = cmad2(m0, x, m1, y);
                a[e | (1 << JR)] = cmad2(m2, x, m3, y);
            }
        }
    }
}
template <int JR, int NA, typename C, typename R>
__device__ __forceinline__ void app_realm(C (&a)[NA], const R* __restrict__ m, uint32_t em)
{
    const R m0 = m[0], m1 = m[2], m2 = m[4], m3 = m[6];
#pragma unroll
    for (int e = 0; e < NA; ++e) {
        if (!(e & (1 << JR))) {
            if (em & (1U << e)) {
                const C x = a[e], y = a[e | (1 << JR)];
                a[e] = mk<R>(m0 * x.x + m1 * y.x, m0 * x.y + m1 * y.y);
                a[e | (1 << JR)] = mk<R>(m2 * x.x + m3 * y.x, m2 * x.y + m3 * y.y);
            }
        }
    }
}
template <int JR, int NA, typename C> __device__ __forceinline__ void app_xswap(C (&a)[NA], uint32_t em)
{
#pragma unroll
    for (int e = 0; e < NA; ++e) {
        if (!(e & (1 << JR))) {
            if (em & (1U << e)) {
                const C x = a[e];
                a[e] = a[e | (1 << JR)];
                a[e | (1 << JR)] = x;
            }
        }
    }
}

template <int NA, typename C, typename R> struct TargetDispatch {
    static __device__ __forceinline__ void run(C (&a)[NA], int kind, int jr, const R* m, uint32_t em)
    {
#define SV_CASE(J)                                                                                                     \
    case J:                                                                                                            \
        if ((1 << J) < NA) {                                                                                           \
            if (kind == OP_GENERAL)                                                                                    \
                app_general<((1 << J) < NA ? J : 0), NA, C, R>(a, m, em);                                              \
            else if (kind == OP_REALM)                                                                                 \
                app_realm<((1 << J) < NA ? J : 0), NA, C, R>(a, m, em);                                                \
            else                                                                                                       \
                app_xswap<((1 << J) < NA ? J : 0), NA, C>(a, em);                                                      \
        }                                                                                                              \
        break;
        switch (jr) {
            SV_CASE(0)
            SV_CASE(1)
            SV_CASE(2)
            SV_CASE(3)
            SV_CASE(4)
        default:
            break;
        }
#undef SV_CASE
    }
};

template <typename R, int KC, int RB, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB)
    k_fused_sweep(typename Cx<R>::type* __restrict__ psi, const unsigned char* __restrict__ prog, uint32_t progBytes, uint64_t nTiles)
{
    typedef typename Cx<R>::type C;
    constexpr int APC = Apc<R>::v;
    constexpr int NCH = 1 << RB;
    constexpr int NA = NCH * APC;
    static_assert(NA <= MAX_NA, "register sub-block too large");
    extern __shared__ __align__(1024) unsigned char smem[];
    uint4* tile = reinterpret_cast<uint4*>(smem);
    unsigned char* sprog = smem + ((size_t)16 << KC);
    __shared__ uint64_t rowOff[1 << 8];

    const int tid = threadIdx.x;
    for (uint32_t i = tid; i < progBytes / 16; i += NT) {
        reinterpret_cast<uint4*>(sprog)[i] = reinterpret_cast<const uint4*>(prog)[i];
    }
    __syncthreads();
    const DevSweep& sw = *reinterpret_cast<const DevSweep*>(sprog);
    const DevOp<R>* ops = reinterpret_cast<const DevOp<R>*>(sprog + sizeof(DevSweep));
    const int kc = sw.kc;
    const uint32_t nChunk = 1U << kc;
    const int lcb = sw.lowAmpBits - (APC == 2 ? 1 : 0); // low (contiguous) chunk bits
    const uint32_t colMask = (1U << lcb) - 1U;
    // global amplitude offset of each chunk row (depends only on the sweep's high qubits)
    for (uint32_t r = tid; r < (nChunk >> lcb); r += NT) {
        uint64_t off = 0;
        for (int h = 0; h < sw.nHigh; ++h) {
            if ((r >> h) & 1U) {
                off |= sw.highPow[h];
            }
        }
        rowOff[r] = off;
    }
    __syncthreads();

    for (uint64_t t = blockIdx.x; t < nTiles; t += gridDim.x) {
        uint64_t base = t << sw.lowAmpBits;
        for (int h = 0; h < sw.nHigh; ++h) {
            const uint64_t lo = base & sw.highLow[h];
            base = ((base ^ lo) << 1) | lo;
        }
        // ---- load: global -> swizzled smem ----------------------------------------------------------------------
        for (uint32_t c0 = tid; c0 < nChunk; c0 += NT * 8) {
            uint4 v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const uint32_t c = c0 + u * NT;
                if (c < nChunk) {
                    const C* g = psi + base + rowOff[c >> lcb] + (uint64_t)(c & colMask) * APC;
                    v[u] = ld_stream(reinterpret_cast<const uint4*>(g));
                }
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const uint32_t c = c0 + u * NT;
                if (c < nChunk) {
                    tile[swz(c)] = v[u];
                }
            }
        }
        __syncthreads();
        // ---- passes ----------------------------------------------------------------------------------------------
        for (int p = 0; p < sw.nPass; ++p) {
            const DevPass& ps = sw.pass[p];
            const uint32_t nSub = nChunk >> RB;
            for (uint32_t s = tid; s < nSub; s += NT) {
                uint32_t sbc = 0;
                for (int i = 0; i < ps.nsb; ++i) {
                    sbc |= ((s >> i) & 1U) << ps.sbit[i];
                }
                C a[NA];
#pragma unroll
                for (int e = 0; e < NCH; ++e) {
                    const uint4 v = tile[swz(sbc | ps.roffc[e])];
                    *reinterpret_cast<uint4*>(&a[e * APC]) = v;
                }
                const uint32_t xsb = sbc * APC;
                for (int o = ps.opBegin; o < ps.opEnd; ++o) {
                    const DevOp<R>& op = ops[o];
                    if ((base & op.omask) != op.oval) {
                        continue;
                    }
                    const uint32_t em = ((xsb & op.lmaskSb) == op.lvalSb) ? op.emask : 0U;
                    if (op.kind == OP_PHASE) {
                        const C ph = mk<R>(op.m[0], op.m[1]);
#pragma unroll
                        for (int e = 0; e < NA; ++e) {
                            if (em & (1U << e)) {
                                a[e] = cmul<C>(ph, a[e]);
                            }
                        }
                    } else {
                        TargetDispatch<NA, C, R>::run(a, op.kind, op.jr, op.m, em);
                    }
                }
#pragma unroll
                for (int e = 0; e < NCH; ++e) {
                    tile[swz(sbc | ps.roffc[e])] = *reinterpret_cast<const uint4*>(&a[e * APC]);
                }
            }
            __syncthreads();
        }
        // ---- store: swizzled smem -> global ------------------------------------------------------------------------
        for (uint32_t c0 = tid; c0 < nChunk; c0 += NT * 8) {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const uint32_t c = c0 + u * NT;
                if (c < nChunk) {
                    C* g = psi + base + rowOff[c >> lcb] + (uint64_t)(c & colMask) * APC;
                    st_stream(reinterpret_cast<uint4*>(g), tile[swz(c)]);
                }
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------------------
// host-side scheduler
// ---------------------------------------------------------------------------------------------------------
struct HostOp {
    int kind;       // OP_*
    int tq;         // target qubit (-1 for OP_PHASE)
    uint64_t cmask; // predicate mask over qubits (controls; for OP_PHASE includes the phased qubit)
    uint64_t cval;
    double m[8];
};

static inline uint64_t bitq(int q) { return 1ULL << q; }

static void lower_queue(const std::vector<GateOp>& q, std::vector<HostOp>& out)
{
    out.clear();
    out.reserve(q.size() * 2);
    for (const GateOp& g : q) {
        HostOp h;
        memset(&h, 0, sizeof(h));
        if (g.kind == 1) { // diagonal: one predicated phase per non-unit diagonal entry
            const bool one0 = (g.m[0] == 1.0 && g.m[1] == 0.0), one3 = (g.m[6] == 1.0 && g.m[7] == 0.0);
            if (!one0) {
                h.kind = OP_PHASE;
                h.tq = -1;
                h.cmask = g.cmask | bitq(g.target);
                h.cval = g.cval;
                h.m[0] = g.m[0];
                h.m[1] = g.m[1];
                out.push_back(h);
            }
            if (!one3) {
                h.kind = OP_PHASE;
                h.tq = -1;
                h.cmask = g.cmask | bitq(g.target);
                h.cval = g.cval | bitq(g.target);
                h.m[0] = g.m[6];
                h.m[1] = g.m[7];
                out.push_back(h);
            }
            continue;
        }
        h.tq = g.target;
        h.cmask = g.cmask;
        h.cval = g.cval;
        memcpy(h.m, g.m, sizeof(h.m));
        if (g.kind == 2 && g.m[2] == 1.0 && g.m[3] == 0.0 && g.m[4] == 1.0 && g.m[5] == 0.0) {
            h.kind = OP_XSWAP;
        } else if (g.m[1] == 0.0 && g.m[3] == 0.0 && g.m[5] == 0.0 && g.m[7] == 0.0) {
            h.kind = OP_REALM;
        } else {
            h.kind = OP_GENERAL;
        }
        out.push_back(h);
    }
}

struct TileCfg {
    int n;    // qubits
    int apcLog; // 1 for fp32, 0 for fp64
    int KC;   // max tile chunk bits
    int RB;   // register chunk bits per pass
    int L;    // low (contiguous) amplitude bits
    int kA;   // tile amplitude bits actually used
    int H;    // capacity of high qubits
};

static TileCfg make_cfg(int n, int prec, int KC, int RB, int Lpref)
{
    TileCfg c;
    c.n = n;
    c.apcLog = (prec == 32) ? 1 : 0;
    c.KC = KC;
    c.RB = RB;
    const int kAmax = KC + c.apcLog;
    c.kA = std::min(n, kAmax);
    c.L = std::min(Lpref, c.kA);
    if (n <= kAmax) {
        c.L = c.kA; // whole state is one tile
    }
    c.H = c.kA - c.L;
    return c;
}

// Greedy, order-preserving selection with commutation-aware skipping.
//   fits(op)   : can the op be executed under the current resource set (may grow the set)
// Ops that are skipped block later ops that do not commute with them.
template <typename FitFn>
static void greedy_select(std::vector<HostOp>& pending, std::vector<HostOp>& taken, size_t maxTake, size_t lookahead, FitFn fits)
{
    uint64_t blockedT = 0, blockedD = 0;
    std::vector<HostOp> rest;
    rest.reserve(pending.size());
    size_t i = 0;
    for (; i < pending.size(); ++i) {
        const HostOp& op = pending[i];
        if (i >= lookahead || taken.size() >= maxTake) {
            break;
        }
        const uint64_t usesT = op.tq >= 0 ? bitq(op.tq) : 0;
        const uint64_t usesD = op.cmask;
        bool conflict = (usesT & (blockedT | blockedD)) || (usesD & blockedT);
        if (!conflict && !fits(op)) {
            conflict = true;
        }
        if (conflict) {
            blockedT |= usesT;
            blockedD |= usesD;
            rest.push_back(op);
        } else {
            taken.push_back(op);
        }
    }
    for (; i < pending.size(); ++i) {
        rest.push_back(pending[i]);
    }
    pending.swap(rest);
}

struct PassPlan {
    std::vector<HostOp> ops;
    std::vector<int> regQ; // register qubits of this pass (excluding the implicit qubit 0 of fp32)
};
struct SweepPlan {
    std::vector<int> highQ; // ascending
    std::vector<PassPlan> passes;
    size_t nOps = 0;
};

static void plan_sweep(std::vector<HostOp>& pending, const TileCfg& cfg, SweepPlan& sp)
{
    // ---- choose the tile's high qubits and the ops of this sweep ----
    uint64_t inTile = (cfg.L >= 64) ? ~0ULL : (bitq(cfg.L) - 1U);
    int freeHigh = cfg.H;
    std::vector<HostOp> sel;
    greedy_select(pending, sel, MAX_OPS, 2048, [&](const HostOp& op) {
        if (op.tq < 0 || (inTile & bitq(op.tq))) {
            return true;
        }
        if (freeHigh > 0) {
            inTile |= bitq(op.tq);
            --freeHigh;
            return true;
        }
        return false;
    });
    sp.highQ.clear();
    for (int q = cfg.L; q < cfg.n; ++q) {
        if (inTile & bitq(q)) {
            sp.highQ.push_back(q);
        }
    }
    // pad the tile with arbitrary high qubits so that its size is fixed (top-down, any unused qubit)
    for (int q = cfg.n - 1; q >= cfg.L && (int)sp.highQ.size() < cfg.H; --q) {
        if (!(inTile & bitq(q))) {
            inTile |= bitq(q);
            sp.highQ.push_back(q);
        }
    }
    std::sort(sp.highQ.begin(), sp.highQ.end());
    // ---- split into passes by register capacity ----
    sp.passes.clear();
    sp.nOps = sel.size();
    while (!sel.empty() && (int)sp.passes.size() < MAX_PASS) {
        PassPlan pp;
        uint64_t regSet = cfg.apcLog ? 1ULL : 0ULL; // fp32: qubit 0 is always register-resident
        int freeReg = cfg.RB;
        greedy_select(sel, pp.ops, MAX_OPS, 4096, [&](const HostOp& op) {
            if (op.tq < 0 || (regSet & bitq(op.tq))) {
                return true;
            }
            if (freeReg > 0) {
                regSet |= bitq(op.tq);
                --freeReg;
                pp.regQ.push_back(op.tq);
                return true;
            }
            return false;
        });
        sp.passes.push_back(pp);
    }
    if (!sel.empty()) {
        // more passes than the descriptor holds: give the remainder back (order among them is preserved)
        sp.nOps -= sel.size();
        sel.insert(sel.end(), pending.begin(), pending.end());
        pending.swap(sel);
    }
}

// tile-local amplitude bit of a tile qubit
static int tile_bit(const TileCfg& cfg, const std::vector<int>& highQ, int q)
{
    if (q < cfg.L) {
        return q;
    }
    for (size_t h = 0; h < highQ.size(); ++h) {
        if (highQ[h] == q) {
            return cfg.L + (int)h;
        }
    }
    return -1;
}

template <typename R> static size_t encode_sweep(const SweepPlan& sp, const TileCfg& cfg, std::vector<unsigned char>& buf)
{
    const size_t start = buf.size();
    const size_t bytes = ((sizeof(DevSweep) + sp.nOps * sizeof(DevOp<R>)) + 15U) & ~(size_t)15U;
    buf.resize(start + bytes, 0);
    DevSweep* ds = reinterpret_cast<DevSweep*>(buf.data() + start);
    DevOp<R>* dops = reinterpret_cast<DevOp<R>*>(buf.data() + start + sizeof(DevSweep));
    ds->nHigh = (int)sp.highQ.size();
    ds->lowAmpBits = cfg.L;
    ds->kc = cfg.kA - cfg.apcLog;
    ds->nPass = (int)sp.passes.size();
    ds->nOps = (int)sp.nOps;
    uint64_t tileMask = bitq(cfg.L) - 1U;
    for (int h = 0; h < ds->nHigh; ++h) {
        ds->highLow[h] = bitq(sp.highQ[h]) - 1U;
        ds->highPow[h] = bitq(sp.highQ[h]);
        tileMask |= bitq(sp.highQ[h]);
    }
    const int kc = ds->kc;
    const int APC = 1 << cfg.apcLog;
    const int NA = (1 << cfg.RB) * APC;
    int o = 0;
    for (int p = 0; p < ds->nPass; ++p) {
        const PassPlan& pp = sp.passes[p];
        DevPass& dp = ds->pass[p];
        // register chunk bits: targets first, then fill from the top with unused chunk bits
        std::vector<int> rb;
        uint32_t used = 0;
        for (int q : pp.regQ) {
            const int cb = tile_bit(cfg, sp.highQ, q) - cfg.apcLog;
            rb.push_back(cb);
            used |= 1U << cb;
        }
        for (int cb = kc - 1; cb >= 0 && (int)rb.size() < cfg.RB; --cb) {
            if (!(used & (1U << cb))) {
                rb.push_back(cb);
                used |= 1U << cb;
            }
        }
        std::sort(rb.begin(), rb.end());
        // sub-block index bits: lanes first take one free bit from each bank class {0,3},{1,4},{2,5}
        std::vector<int> sb;
        uint32_t taken = used;
        for (int pcl = 0; pcl < 3; ++pcl) {
            for (int cand : { pcl, pcl + 3 }) {
                if (cand < kc && !(taken & (1U << cand))) {
                    sb.push_back(cand);
                    taken |= 1U << cand;
                    break;
                }
            }
        }
        for (int cb = 0; cb < kc; ++cb) {
            if (!(taken & (1U << cb))) {
                sb.push_back(cb);
                taken |= 1U << cb;
            }
        }
        dp.nsb = (int)sb.size();
        for (size_t i = 0; i < sb.size(); ++i) {
            dp.sbit[i] = (unsigned char)sb[i];
        }
        // register chunk offsets and per-register-amplitude tile-local amplitude offsets
        uint32_t roffA[MAX_NA];
        uint32_t regAmpMask = cfg.apcLog ? 1U : 0U;
        for (int cb : rb) {
            regAmpMask |= 1U << (cb + cfg.apcLog);
        }
        for (int e = 0; e < (1 << cfg.RB); ++e) {
            uint32_t off = 0;
            for (int b = 0; b < cfg.RB; ++b) {
                if ((e >> b) & 1) {
                    off |= 1U << rb[b];
                }
            }
            dp.roffc[e] = (unsigned short)off;
            for (int w = 0; w < APC; ++w) {
                roffA[e * APC + w] = (off << cfg.apcLog) | (uint32_t)w;
            }
        }
        dp.opBegin = o;
        for (const HostOp& hop : pp.ops) {
            DevOp<R>& d = dops[o++];
            d.kind = hop.kind;
            d.omask = hop.cmask & ~tileMask;
            d.oval = hop.cval & ~tileMask;
            // tile-local predicate
            uint32_t lmask = 0, lval = 0;
            for (uint64_t m = hop.cmask & tileMask; m; m &= m - 1U) {
                const int q = __builtin_ctzll(m);
                const int tb = tile_bit(cfg, sp.highQ, q);
                lmask |= 1U << tb;
                if (hop.cval & bitq(q)) {
                    lval |= 1U << tb;
                }
            }
            d.lmaskSb = lmask & ~regAmpMask;
            d.lvalSb = lval & ~regAmpMask;
            const uint32_t lmr = lmask & regAmpMask, lvr = lval & regAmpMask;
            uint32_t em = 0;
            for (int e = 0; e < NA; ++e) {
                if ((roffA[e] & lmr) == lvr) {
                    em |= 1U << e;
                }
            }
            d.emask = em;
            d.jr = 0;
            if (hop.tq >= 0) {
                const int tb = tile_bit(cfg, sp.highQ, hop.tq);
                // register-bit index of the target: position among the register amplitude bits
                int jr = 0;
                for (int b = 0; b < tb; ++b) {
                    if (regAmpMask & (1U << b)) {
                        ++jr;
                    }
                }
                d.jr = jr;
            }
            for (int k = 0; k < 8; ++k) {
                d.m[k] = (R)hop.m[k];
            }
        }
        dp.opEnd = o;
    }
    return bytes;
}

// per-state program arena (device + pinned host), guarded by an event
struct Arena {
    unsigned char* dev = nullptr;
    unsigned char* host = nullptr;
    size_t cap = 0;
    cudaEvent_t done = nullptr;
    bool pending = false;
};
static std::vector<std::pair<State*, Arena>>& arenas()
{
    static std::vector<std::pair<State*, Arena>> a;
    return a;
}
static std::mutex& arena_mutex()
{
    static std::mutex m;
    return m;
}
static Arena* get_arena(State* s)
{
    std::lock_guard<std::mutex> lk(arena_mutex());
    for (auto& kv : arenas()) {
        if (kv.first == s) {
            return &kv.second;
        }
    }
    arenas().push_back({ s, Arena() });
    return &arenas().back().second;
}
void fused_release(State* s)
{
    std::lock_guard<std::mutex> lk(arena_mutex());
    auto& v = arenas();
    for (size_t i = 0; i < v.size(); ++i) {
        if (v[i].first == s) {
            Arena& a = v[i].second;
            if (a.dev) {
                cudaFree(a.dev);
            }
            if (a.host) {
                cudaFreeHost(a.host);
            }
            if (a.done) {
                cudaEventDestroy(a.done);
            }
            v.erase(v.begin() + i);
            return;
        }
    }
}

bool fused_accepts(const State* s, const GateOp&) { return s->nq >= 5 && s->nq <= 62; }

struct KernelCfg {
    int KC, RB, NT, MINB;
};

template <typename R, int KC, int RB, int NT, int MINB>
static int launch_sweep(State* s, const unsigned char* dprog, uint32_t progBytes, uint64_t nTiles)
{
    auto kern = k_fused_sweep<R, KC, RB, NT, MINB>;
    const size_t shm = ((size_t)16 << KC) + progBytes;
    static unsigned long long attr_set_mask = 0; // per device: the attribute is per-context
    if (!(attr_set_mask & (1ULL << s->dev))) {
        SV_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
            (int)(((size_t)16 << KC) + sizeof(DevSweep) + MAX_OPS * sizeof(DevOp<R>) + 16)));
        attr_set_mask |= 1ULL << s->dev;
    }
    const uint64_t maxGrid = (uint64_t)sm_count(s->dev) * MINB;
    const unsigned grid = (unsigned)std::min<uint64_t>(nTiles, maxGrid);
    kern<<<grid, NT, shm, s->stream>>>(reinterpret_cast<typename Cx<R>::type*>(s->amps), dprog, progBytes, nTiles);
    SV_CUDA(cudaGetLastError());
    return B200SV_OK;
}

int fused_flush(State* s)
{
    if (s->queue.empty()) {
        return B200SV_OK;
    }
    if (!s->amps) {
        s->queue.clear();
        return B200SV_OK;
    }
    std::vector<HostOp> pending;
    lower_queue(s->queue, pending);
    s->queue.clear();
    if (pending.empty()) {
        return B200SV_OK;
    }
    constexpr int KC = 12, RB = 3, NT = 256, MINB = 3;
    const int Lpref = (s->prec == 32) ? 7 : 6;
    const TileCfg cfg = make_cfg(s->nq, s->prec, KC, RB, Lpref);
    Arena* ar = get_arena(s);
    if (!ar->done) {
        SV_CUDA(cudaEventCreateWithFlags(&ar->done, cudaEventDisableTiming));
    }
    // build every sweep of this flush
    std::vector<unsigned char> buf;
    std::vector<std::pair<size_t, size_t>> segs; // (offset, bytes)
    std::vector<size_t> nops;
    while (!pending.empty()) {
        SweepPlan sp;
        plan_sweep(pending, cfg, sp);
        if (!sp.nOps) {
            set_error("fused scheduler made no progress");
            return B200SV_ESTATE;
        }
        const size_t off = buf.size();
        const size_t bytes = (s->prec == 32) ? encode_sweep<float>(sp, cfg, buf) : encode_sweep<double>(sp, cfg, buf);
        segs.push_back({ off, bytes });
        nops.push_back(sp.nOps);
    }
    if (ar->pending) {
        SV_CUDA(cudaEventSynchronize(ar->done));
        ar->pending = false;
    }
    if (ar->cap < buf.size()) {
        if (ar->dev) {
            cudaFree(ar->dev);
            cudaFreeHost(ar->host);
        }
        ar->cap = std::max<size_t>(buf.size() * 2, 1 << 20);
        SV_CUDA(cudaMalloc(&ar->dev, ar->cap));
        SV_CUDA(cudaMallocHost(&ar->host, ar->cap));
    }
    memcpy(ar->host, buf.data(), buf.size());
    SV_CUDA(cudaMemcpyAsync(ar->dev, ar->host, buf.size(), cudaMemcpyHostToDevice, s->stream));
    const uint64_t nTiles = s->dim() >> cfg.kA;
    for (size_t i = 0; i < segs.size(); ++i) {
        if (s->prec == 32) {
            SV_TRY((launch_sweep<float, KC, RB, NT, MINB>(s, ar->dev + segs[i].first, (uint32_t)segs[i].second, nTiles)));
        } else {
            SV_TRY((launch_sweep<double, KC, RB, NT, MINB>(s, ar->dev + segs[i].first, (uint32_t)segs[i].second, nTiles)));
        }
        s->stats.kernel_launches++;
        s->stats.fused_sweeps++;
        s->stats.fused_gates += nops[i];
        s->stats.bytes_swept += 2ULL * s->dim() * s->amp_bytes();
    }
    SV_CUDA(cudaEventRecord(ar->done, s->stream));
    ar->pending = true;
    return B200SV_OK;
}

int fused_plan_dry_run(int n_qubits, int precision, int n_gates, const int* targets, const uint64_t* cmasks, const int* kinds,
    int* n_sweeps, int* n_passes)
{
    std::vector<GateOp> q((size_t)n_gates);
    for (int i = 0; i < n_gates; ++i) {
        GateOp& g = q[i];
        memset(&g, 0, sizeof(g));
        g.target = targets[i];
        g.cmask = cmasks[i];
        g.cval = cmasks[i];
        g.kind = kinds[i];
        if (kinds[i] == 1) { // diagonal: T-like
            g.m[0] = 1.0;
            g.m[6] = 0.6;
            g.m[7] = 0.8;
        } else if (kinds[i] == 2) { // X-like
            g.m[2] = 1.0;
            g.m[4] = 1.0;
        } else {
            g.m[0] = 0.6;
            g.m[2] = 0.8;
            g.m[4] = 0.8;
            g.m[6] = -0.6;
            g.m[1] = (kinds[i] == 3) ? 0.1 : 0.0; // kind 3: complex general
        }
    }
    std::vector<HostOp> pending;
    lower_queue(q, pending);
    constexpr int KC = 12, RB = 3;
    const TileCfg cfg = make_cfg(n_qubits, precision, KC, RB, precision == 32 ? 7 : 6);
    int sweeps = 0, passes = 0;
    while (!pending.empty()) {
        SweepPlan sp;
        plan_sweep(pending, cfg, sp);
        if (!sp.nOps) {
            return B200SV_ESTATE;
        }
        ++sweeps;
        passes += (int)sp.passes.size();
    }
    *n_sweeps = sweeps;
    *n_passes = passes;
    return B200SV_OK;
}

} // namespace b200sv
