// Probe for VERDICT r1 #10 ("TMA, with a measurement either way"): does moving the fused sweep's tile staging onto the TMA engine pay?
//
// The sweep's tile is 2^7 rows of 2^6 amplitudes (512 B each for fp32): the rows are selected by 7 ARBITRARY high qubits and the tile
// base by the remaining 17 index bits, i.e. the tile is a 2x2x2x2x2x2x2x64 box out of a tensor whose other 17 bits are cut into up to 8
// separate bit ranges.  A cp.async.bulk.tensor map has at most 5 dimensions, so the tensor form cannot address such a tile; what TMA
// offers is the 1-D bulk copy (cp.async.bulk, SASS UBLKCP): one instruction per 512-byte row, 128 per tile, instead of 4096 16-byte
// loads.  This probe times exactly that against the per-thread path the engine uses, on the engine's geometry (30 qubits, 64 KB tiles,
// 2 CTAs of 256 threads per SM, persistent grid), as a copy through shared memory (global -> smem -> [optional per-thread touch of every
// chunk] -> global, in place).  Output: one JSON line per variant with ms and GB/s (read + write).
//
//   variant 0: ld.global.cs.v4 / st.shared ... ld.shared / st.global.cs.v4 by all threads (stage_in / stage_out of fused.cu)
//   variant 1: cp.async.bulk global->smem per row + mbarrier, cp.async.bulk smem->global per row + bulk_group
//   variant 2: as 1, double-buffered (2 x 32 KB half tiles in flight per CTA)
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e__ = (x); if (e__ != cudaSuccess) { printf("{\"error\": \"%s at %s:%d\"}\n", cudaGetErrorString(e__), __FILE__, __LINE__); return 1; } } while (0)

constexpr int NT = 256;
constexpr int LOW = 6;          // contiguous amplitudes per row: 2^6 * 8 B = 512 B
constexpr int NHIGH = 7;
constexpr int ROWS = 1 << NHIGH;
constexpr int ROW_BYTES = (1 << LOW) * 8;
constexpr int TILE_BYTES = ROWS * ROW_BYTES; // 64 KB

struct Geo {
    uint64_t highLow[NHIGH]; // (2^q - 1), ascending
    uint64_t highPow[NHIGH]; // 2^q
    uint64_t nTiles;
};

__device__ __forceinline__ uint64_t tile_base(const Geo& g, uint64_t t)
{
    uint64_t base = t << LOW;
    for (int h = 0; h < NHIGH; ++h) {
        const uint64_t lo = base & g.highLow[h];
        base = ((base ^ lo) << 1) | lo;
    }
    return base;
}
__device__ __forceinline__ uint64_t row_off(const Geo& g, uint32_t r)
{
    uint64_t off = 0;
    for (int h = 0; h < NHIGH; ++h) {
        if ((r >> h) & 1U) {
            off |= g.highPow[h];
        }
    }
    return off;
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint4 ld_cs(const uint4* p)
{
    uint4 v;
    asm volatile("ld.global.cs.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_cs(uint4* p, uint4 v)
{
    asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void touch(unsigned char* tile, int tid, float f)
{
    // every thread scales its 16 chunks (what a one-op pass does at least): conflict-free linear addressing
    for (int i = 0; i < TILE_BYTES / 16 / NT; ++i) {
        float4* p = reinterpret_cast<float4*>(tile) + (i * NT + tid);
        float4 v = *p;
        v.x *= f; v.y *= f; v.z *= f; v.w *= f;
        *p = v;
    }
}

// ---- variant 0: per-thread 16-byte loads/stores ---------------------------------------------------------------------
__global__ void __launch_bounds__(NT, 2) k_threads(float2* psi, const __grid_constant__ Geo g, int work, float f)
{
    extern __shared__ __align__(1024) unsigned char tile[];
    __shared__ uint64_t rowOff[ROWS];
    const int tid = threadIdx.x;
    for (int r = tid; r < ROWS; r += NT) {
        rowOff[r] = row_off(g, r);
    }
    __syncthreads();
    constexpr int CPR = ROW_BYTES / 16; // chunks per row
    for (uint64_t t = blockIdx.x; t < g.nTiles; t += gridDim.x) {
        float2* tp = psi + tile_base(g, t);
        for (uint32_t c0 = tid; c0 < TILE_BYTES / 16; c0 += 8 * NT) {
            uint4 v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const uint32_t c = c0 + u * NT;
                v[u] = ld_cs(reinterpret_cast<const uint4*>(tp + rowOff[c / CPR]) + (c % CPR));
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                reinterpret_cast<uint4*>(tile)[c0 + u * NT] = v[u];
            }
        }
        __syncthreads();
        if (work) {
            touch(tile, tid, f);
            __syncthreads();
        }
#pragma unroll 4
        for (uint32_t c = tid; c < TILE_BYTES / 16; c += NT) {
            st_cs(reinterpret_cast<uint4*>(tp + rowOff[c / CPR]) + (c % CPR), reinterpret_cast<const uint4*>(tile)[c]);
        }
        __syncthreads();
    }
}

// ---- mbarrier / bulk-copy helpers -------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity)
{
    // bounded: a lost transaction must not hang the box
    for (int spin = 0; spin < (1 << 22); ++spin) {
        uint32_t ok;
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (ok) {
            return true;
        }
    }
    return false;
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- variant 1: one bulk copy per row, issued by the first ROWS threads (one row each) -----------------------------------------
__global__ void __launch_bounds__(NT, 2) k_bulk(float2* psi, const __grid_constant__ Geo g, int work, float f, int* fail)
{
    extern __shared__ __align__(1024) unsigned char tile[];
    __shared__ __align__(8) uint64_t bar;
    const int tid = threadIdx.x;
    const uint64_t myRow = (tid < ROWS) ? row_off(g, tid) : 0;
    if (tid == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    uint32_t parity = 0;
    for (uint64_t t = blockIdx.x; t < g.nTiles; t += gridDim.x) {
        float2* tp = psi + tile_base(g, t);
        if (tid == 0) {
            mbar_expect_tx(&bar, TILE_BYTES);
        }
        __syncthreads(); // expect_tx is posted before any complete_tx can arrive
        if (tid < ROWS) {
            bulk_g2s(tile + tid * ROW_BYTES, tp + myRow, ROW_BYTES, &bar);
        }
        if (!mbar_wait(&bar, parity)) {
            if (tid == 0) {
                atomicExch(fail, 1);
            }
            return;
        }
        parity ^= 1U;
        if (work) {
            touch(tile, tid, f);
            fence_async_smem(); // generic-proxy writes -> visible to the bulk store
        }
        __syncthreads();
        if (tid < ROWS) {
            bulk_s2g(tp + myRow, tile + tid * ROW_BYTES, ROW_BYTES);
            bulk_commit();
            bulk_wait_read0(); // smem may be overwritten once the store has READ it
        }
        __syncthreads();
    }
}

// ---- variant 2: two half tiles in flight (load of half B overlaps touch/store of half A) ---------------------------------------
__global__ void __launch_bounds__(NT, 2) k_bulk2(float2* psi, const __grid_constant__ Geo g, int work, float f, int* fail)
{
    extern __shared__ __align__(1024) unsigned char tile[];
    __shared__ __align__(8) uint64_t bar[2];
    const int tid = threadIdx.x;
    constexpr int HROWS = ROWS / 2, HBYTES = TILE_BYTES / 2;
    const uint64_t myRow0 = (tid < HROWS) ? row_off(g, tid) : 0, myRow1 = (tid < HROWS) ? row_off(g, tid + HROWS) : 0;
    if (tid == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    uint32_t parity = 0;
    for (uint64_t t = blockIdx.x; t < g.nTiles; t += gridDim.x) {
        float2* tp = psi + tile_base(g, t);
        if (tid == 0) {
            mbar_expect_tx(&bar[0], HBYTES);
            mbar_expect_tx(&bar[1], HBYTES);
        }
        __syncthreads();
        if (tid < HROWS) {
            bulk_g2s(tile + tid * ROW_BYTES, tp + myRow0, ROW_BYTES, &bar[0]);
            bulk_g2s(tile + HBYTES + tid * ROW_BYTES, tp + myRow1, ROW_BYTES, &bar[1]);
        }
        for (int h = 0; h < 2; ++h) {
            if (!mbar_wait(&bar[h], parity)) {
                if (tid == 0) {
                    atomicExch(fail, 1);
                }
                return;
            }
            if (work) {
                for (int i = 0; i < HBYTES / 16 / NT; ++i) {
                    float4* p = reinterpret_cast<float4*>(tile + h * HBYTES) + (i * NT + tid);
                    float4 v = *p;
                    v.x *= f; v.y *= f; v.z *= f; v.w *= f;
                    *p = v;
                }
                fence_async_smem();
            }
            __syncthreads();
            if (tid < HROWS) {
                bulk_s2g(tp + (h ? myRow1 : myRow0), tile + h * HBYTES + tid * ROW_BYTES, ROW_BYTES);
                bulk_commit();
            }
        }
        parity ^= 1U;
        if (tid < HROWS) {
            bulk_wait_read0();
        }
        __syncthreads();
    }
}

__global__ void k_fill(float2* p, uint64_t n)
{
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        p[i] = make_float2((float)(i & 1023U), 1.0f);
    }
}
__global__ void k_check(const float2* p, uint64_t n, float expectScale, unsigned long long* bad)
{
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const float2 v = p[i];
        if (v.x != (float)(i & 1023U) * expectScale || v.y != expectScale) {
            atomicAdd(bad, 1ULL);
        }
    }
}

int main(int argc, char** argv)
{
    const int nq = argc > 1 ? atoi(argv[1]) : 30;
    const int reps = argc > 2 ? atoi(argv[2]) : 5;
    const int highs[NHIGH] = { 8, 11, 14, 17, 20, 23, 26 };
    Geo g;
    for (int h = 0; h < NHIGH; ++h) {
        const int q = highs[h] < nq ? highs[h] : nq - NHIGH + h;
        g.highPow[h] = 1ULL << q;
        g.highLow[h] = (1ULL << q) - 1ULL;
    }
    const uint64_t dim = 1ULL << nq;
    g.nTiles = dim >> (LOW + NHIGH);
    float2* psi = nullptr;
    CK(cudaMalloc(&psi, dim * sizeof(float2)));
    int* fail = nullptr;
    unsigned long long* bad = nullptr;
    CK(cudaMalloc(&fail, sizeof(int)));
    CK(cudaMalloc(&bad, sizeof(unsigned long long)));
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const unsigned grid = (unsigned)(2 * sms);
    CK(cudaFuncSetAttribute(k_threads, cudaFuncAttributeMaxDynamicSharedMemorySize, TILE_BYTES));
    CK(cudaFuncSetAttribute(k_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, TILE_BYTES));
    CK(cudaFuncSetAttribute(k_bulk2, cudaFuncAttributeMaxDynamicSharedMemorySize, TILE_BYTES));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    const char* names[3] = { "threads_ld_st_v4", "tma_bulk_row_copies", "tma_bulk_row_copies_2_halves" };
    for (int work = 0; work < 2; ++work) {
        for (int var = 0; var < 3; ++var) {
            k_fill<<<4 * sms, 256>>>(psi, dim);
            CK(cudaMemset(fail, 0, sizeof(int)));
            CK(cudaMemset(bad, 0, sizeof(unsigned long long)));
            const float f = 1.0f; // the touch multiplies by 1: the data check stays exact for any number of repetitions
            auto launch = [&]() {
                if (var == 0) {
                    k_threads<<<grid, NT, TILE_BYTES>>>(psi, g, work, f);
                } else if (var == 1) {
                    k_bulk<<<grid, NT, TILE_BYTES>>>(psi, g, work, f, fail);
                } else {
                    k_bulk2<<<grid, NT, TILE_BYTES>>>(psi, g, work, f, fail);
                }
            };
            launch(); // warm-up
            CK(cudaDeviceSynchronize());
            CK(cudaEventRecord(e0));
            for (int r = 0; r < reps; ++r) {
                launch();
            }
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            CK(cudaGetLastError());
            float ms = 0;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            ms /= reps;
            k_check<<<4 * sms, 256>>>(psi, dim, 1.0f, bad);
            int hf = 0;
            unsigned long long hb = 0;
            CK(cudaMemcpy(&hf, fail, sizeof(int), cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(&hb, bad, sizeof(hb), cudaMemcpyDeviceToHost));
            const double gb = 2.0 * (double)dim * 8.0 / 1e9;
            printf("{\"probe\": \"tile staging through smem, %d q fp32, 64 KB tiles (2^7 rows x 512 B), 2 CTAs x 256 thr per SM\", \"variant\": \"%s\", "
                   "\"touch_every_chunk\": %d, \"ms\": %.4f, \"GBps\": %.1f, \"mbarrier_timeout\": %d, \"bad_amplitudes\": %llu}\n",
                nq, names[var], work, ms, gb / (ms / 1e3), hf, hb);
            fflush(stdout);
        }
    }
    return 0;
}
