// Probe: is the weight-stationary form of the tensor-core instruction (tcgen05.mma.ws, B held in a collector buffer
// across the M tiles of one K step) faster than the plain form for the 2x64 tower's shape (M = 128, N = 64, K = 16,
// four M tiles per weight block)?  Standalone; build + run:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I kami_b200/csrc -o tools/_mma_ws_probe tools/mma_ws_probe.cu
//   tools/_mma_ws_probe
// Prints, per mode, the cycles for 144 MMAs (one 3x3 layer of a 7-board item) and whether the accumulators equal a
// CUDA-core reference in the plain M = 128 layout (TMEM lane = row, column = n).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include "ptx.cuh"

using namespace kb;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 1; } } while (0)

__device__ __forceinline__ void mbar_test_wait_loop(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void mbar_test_wait_relaxed_loop(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.relaxed.cta.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
}
template <int MODE>
__device__ __forceinline__ void mma_issue(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc, int mt) {
    if constexpr (MODE == 0) {
        ptx::mma_bf16(d, a, b, idesc, acc);
    } else if constexpr (MODE == 1) {  // ws, no collector hint
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.ws.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b),
                     "r"(idesc), "r"(acc)
                     : "memory");
    } else if constexpr (MODE == 2) {  // ws, B in collector b0 across the four tiles
        if (mt == 0)
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.ws.cta_group::1.kind::f16.collector::b0::fill [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
                         "l"(a), "l"(b), "r"(idesc), "r"(acc)
                         : "memory");
        else if (mt == 3)
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.ws.cta_group::1.kind::f16.collector::b0::lastuse [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
                         "l"(a), "l"(b), "r"(idesc), "r"(acc)
                         : "memory");
        else
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.ws.cta_group::1.kind::f16.collector::b0::use [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
                         "l"(a), "l"(b), "r"(idesc), "r"(acc)
                         : "memory");
    } else {  // MODE 3: ws, alternating collector buffers b0 / b1 per K step (fill of the next overlaps use of the last)
        // (the buffer is chosen by the caller through mt's high bit)
        const int buf = mt >> 4, m = mt & 15;
        if (buf == 0) {
            if (m == 0)
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.ws.cta_group::1.kind::f16.collector::b0::fill [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
                             "l"(a), "l"(b), "r"(idesc), "r"(acc)
                             : "memory");
            else if (m == 3)
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.ws.cta_group::1.kind::f16.collector::b0::lastuse [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
                             "l"(a), "l"(b), "r"(idesc), "r"(acc)
                             : "memory");
            else
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.ws.cta_group::1.kind::f16.collector::b0::use [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
                             "l"(a), "l"(b), "r"(idesc), "r"(acc)
                             : "memory");
        } else {
            if (m == 0)
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.ws.cta_group::1.kind::f16.collector::b1::fill [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
                             "l"(a), "l"(b), "r"(idesc), "r"(acc)
                             : "memory");
            else if (m == 3)
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.ws.cta_group::1.kind::f16.collector::b1::lastuse [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
                             "l"(a), "l"(b), "r"(idesc), "r"(acc)
                             : "memory");
            else
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.ws.cta_group::1.kind::f16.collector::b1::use [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
                             "l"(a), "l"(b), "r"(idesc), "r"(acc)
                             : "memory");
        }
    }
}

// logical element (row, k) of a K-major 128B-swizzled operand whose rows are 128-byte lines
__device__ __forceinline__ uint32_t sw_off(int row, int k) {
    const int chunk = (k >> 3) ^ (row & 7);
    return (uint32_t)row * 128u + (uint32_t)chunk * 16u + (uint32_t)(k & 7) * 2u;
}
__device__ __forceinline__ float a_val(int row, int k) { return (float)(((row * 5 + k * 3) % 9) - 4); }
__device__ __forceinline__ float b_val(int n, int k) { return (float)(((n * 7 + k) % 5) - 2); }

constexpr int TILES = 4, NT = 64, KB = 4 /* K steps of 16 per 64-wide block */, REPS = 9;

// LAYOUT 0: dense tiles (8-row groups 1024 B apart, 1024-byte aligned).  LAYOUT 1: the tower's tall image -- a tile is 16
// image rows of 8 pixels, rows 1280 B apart (pitch 10), first pixel one line in: no 8-row group is 1024-byte aligned.
// EXTRA bit 0: a commit to a second barrier after every weight block (16 MMAs), as the tower frees its ring stages;
// bit 1: 16 more warps wait on the completion barrier while the MMAs run, as the tower's epilogue warps do.
template <int MODE, int LAYOUT, int EXTRA>
// bit 2: the tower's ring protocol -- a producer warp streams one 8 KB block per 16 MMAs from global memory into a 5-stage
// ring with full / empty barriers; the MMA warp waits for the stage, fences, elects, issues the block's 16 MMAs (operands
// still read from the prefilled copy, the ring only carries the traffic and the hand-shakes) and commits the stage free.
__global__ void __launch_bounds__(640) probe(long long* cycles, int* mismatches, float* sample, const uint4* gsrc) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    unsigned char* A = smem;                     // 4 tiles x 128 rows x 128 B
    unsigned char* B = smem + 90112;  // REPS blocks x 64 rows x 128 B (behind the larger of the two A layouts)
    __shared__ uint64_t bar, bar2, b_full[9], b_empty[9];
    // bit 4: the producer arrives on the full barriers without copying anything; bit 5: nine full barriers, nothing waits
    // for a free stage (every block is requested at once: no refill dependency)
    constexpr int NBAR = (EXTRA & 32) ? 9 : 5;
    unsigned char* ring = smem + 90112 + REPS * 8192;
    __shared__ uint32_t tmem_slot, flag[9];
    if (threadIdx.x < 9) flag[threadIdx.x] = 1;
    const int tid = threadIdx.x, warp = tid >> 5;
    constexpr uint32_t A_SBO = LAYOUT ? 1280 : 1024, A_TILE = 16 * A_SBO, A_START = LAYOUT ? 1280 + 128 : 0;
    for (int i = tid; i < TILES * 128 * 64; i += blockDim.x) {
        const int row = i / 64, k = i % 64;
        if (LAYOUT == 0) *reinterpret_cast<__nv_bfloat16*>(A + sw_off(row, k)) = __float2bfloat16(a_val(row, k));
        else {
            const uint32_t line = A_START + (uint32_t)(row / 128) * A_TILE + (uint32_t)((row % 128) / 8) * A_SBO + (uint32_t)(row % 8) * 128;
            const uint32_t chunk = (uint32_t)(k >> 3) ^ ((line >> 7) & 7);
            *reinterpret_cast<__nv_bfloat16*>(A + line + chunk * 16 + (k & 7) * 2) = __float2bfloat16(a_val(row, k));
        }
    }
    for (int i = tid; i < REPS * 64 * 64; i += blockDim.x) {
        const int blk = i / 4096, n = (i / 64) % 64, k = i % 64;
        *reinterpret_cast<__nv_bfloat16*>(B + blk * 8192 + sw_off(n, k)) = __float2bfloat16(b_val(n + blk, k));
    }
    if (tid == 0) {
        ptx::mbar_init(ptx::smem_u32(&bar), 1);
        ptx::mbar_init(ptx::smem_u32(&bar2), 1000);
        for (int i = 0; i < 9; ++i) {
            ptx::mbar_init(ptx::smem_u32(&b_full[i]), 1);
            ptx::mbar_init(ptx::smem_u32(&b_empty[i]), 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        ptx::tmem_alloc(ptx::smem_u32(&tmem_slot), 256);
        ptx::tmem_relinquish();
    }
    ptx::fence_proxy_async_smem();
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = tmem_slot;
    constexpr uint32_t idesc = ptx::idesc_bf16(128, NT);
    long long t0 = 0, t1 = 0;
    if ((EXTRA & 4) && warp == 0) {
        int stage = 0, ph = 0;
        for (int blk = 0; blk < REPS; ++blk) {
            if (!(EXTRA & 32)) ptx::mbar_wait(ptx::smem_u32(&b_empty[stage]), ph ^ 1);
            if (ptx::elect_one()) {
                if (EXTRA & 16) ptx::mbar_arrive(ptx::smem_u32(&b_full[stage]));
                else {
                    ptx::mbar_arrive_expect_tx(ptx::smem_u32(&b_full[stage]), 8192);
                    ptx::bulk_g2s(ptx::smem_u32(ring + (stage % 5) * 8192), gsrc + blk * 512, 8192, ptx::smem_u32(&b_full[stage]));
                }
            }
            __syncwarp();
            if (++stage == NBAR) { stage = 0; ph ^= 1; }
        }
    }
    if ((EXTRA & 4) && warp == 1) {
        const uint32_t a_hi = ptx::sw128_hi(A_SBO), b_hi = ptx::sw128_hi(1024);
        const long long w0 = clock64();
        while (clock64() - w0 < 6000) {}  // (the ring is full when the layer starts, as in the tower)
        int stage = 0, ph = 0;
        t0 = clock64();
        if (EXTRA & 64) {  // bit 6: one election for the whole layer; the elected lane waits, issues and commits
            if (ptx::elect_one()) {
#pragma unroll 1
                for (int rep = 0; rep < REPS; ++rep) {
                    // bit 7: no waits; bit 9: test_wait; bit 10: relaxed test_wait; bit 11: a volatile shared-memory flag
                    if (EXTRA & 512) mbar_test_wait_loop(ptx::smem_u32(&b_full[stage]), ph);
                    else if (EXTRA & 1024) mbar_test_wait_relaxed_loop(ptx::smem_u32(&b_full[stage]), ph);
                    else if (EXTRA & 2048) { while (*(volatile uint32_t*)&flag[stage] == 0) {} }
                    else if (!(EXTRA & 128)) ptx::mbar_wait(ptx::smem_u32(&b_full[stage]), ph);
#pragma unroll
                    for (int kk = 0; kk < KB; ++kk) {
                        const uint64_t bdesc = ptx::desc_pack(ptx::sw128_lo(ptx::smem_u32(B + rep * 8192)) + kk * 2, b_hi);
#pragma unroll
                        for (int mt = 0; mt < TILES; ++mt) {
                            const uint64_t adesc = ptx::desc_pack(ptx::sw128_lo(ptx::smem_u32(A + A_START + mt * A_TILE)) + kk * 2, a_hi);
                            const int tag = MODE == 3 ? ((kk & 1) << 4) | mt : mt;
                            mma_issue<MODE>(tmem + mt * NT, adesc, bdesc, idesc, (rep | kk) != 0, tag);
                        }
                    }
                    ptx::mma_commit(ptx::smem_u32((EXTRA & 256) ? &bar2 : &b_empty[stage]));  // bit 8: commits never complete a phase
                    if (++stage == NBAR) { stage = 0; ph ^= 1; }
                }
            }
            __syncwarp();
        } else
#pragma unroll 1
        for (int rep = 0; rep < REPS; ++rep) {
            if (!(EXTRA & 128)) ptx::mbar_wait(ptx::smem_u32(&b_full[stage]), ph);
            if (EXTRA & 8) ptx::tc_fence_after();
            if (ptx::elect_one()) {
#pragma unroll
                for (int kk = 0; kk < KB; ++kk) {
                    const uint64_t bdesc = ptx::desc_pack(ptx::sw128_lo(ptx::smem_u32(B + rep * 8192)) + kk * 2, b_hi);
#pragma unroll
                    for (int mt = 0; mt < TILES; ++mt) {
                        const uint64_t adesc = ptx::desc_pack(ptx::sw128_lo(ptx::smem_u32(A + A_START + mt * A_TILE)) + kk * 2, a_hi);
                        const int tag = MODE == 3 ? ((kk & 1) << 4) | mt : mt;
                        mma_issue<MODE>(tmem + mt * NT, adesc, bdesc, idesc, (rep | kk) != 0, tag);
                    }
                }
                ptx::mma_commit(ptx::smem_u32(&b_empty[stage]));
            }
            __syncwarp();
            if (++stage == NBAR) { stage = 0; ph ^= 1; }
        }
        if (ptx::elect_one()) ptx::mma_commit(ptx::smem_u32(&bar));
        __syncwarp();
        const long long t_issue = clock64();
        ptx::mbar_wait(ptx::smem_u32(&bar), 0);
        t1 = clock64();
        if (tid == 32) cycles[blockIdx.x] = t1 - t0, cycles[148 + blockIdx.x] = t_issue - t0;
    }
    if (!(EXTRA & 4) && warp == 1) {
        const uint32_t a_hi = ptx::sw128_hi(A_SBO), b_hi = ptx::sw128_hi(1024);
        t0 = clock64();
        if (ptx::elect_one()) {
            for (int rep = 0; rep < REPS; ++rep)
                for (int kk = 0; kk < KB; ++kk) {
                    const uint64_t bdesc = ptx::desc_pack(ptx::sw128_lo(ptx::smem_u32(B + rep * 8192)) + kk * 2, b_hi);
                    for (int mt = 0; mt < TILES; ++mt) {
                        const uint64_t adesc = ptx::desc_pack(ptx::sw128_lo(ptx::smem_u32(A + A_START + mt * A_TILE)) + kk * 2, a_hi);
                        const int tag = MODE == 3 ? (((rep * KB + kk) & 1) << 4) | mt : mt;
                        mma_issue<MODE>(tmem + mt * NT, adesc, bdesc, idesc, (rep | kk) != 0, tag);
                    }
                    if ((EXTRA & 1) && kk == KB - 1) ptx::mma_commit(ptx::smem_u32(&bar2));
                }
            ptx::mma_commit(ptx::smem_u32(&bar));
        }
        __syncwarp();
        ptx::mbar_wait(ptx::smem_u32(&bar), 0);
        t1 = clock64();
        if (tid == 32) cycles[blockIdx.x] = t1 - t0;
    }
    if (warp >= 4) ptx::mbar_wait(ptx::smem_u32(&bar), 0);  // (EXTRA bit 1: the launch has 20 warps)
    __syncthreads();
    ptx::mbar_wait(ptx::smem_u32(&bar), 0);
    ptx::tc_fence_after();
    if (warp >= 4) return;
    // check: warp w reads lanes 32w..32w+31 of every tile
    int bad = 0;
    for (int mt = 0; mt < TILES; ++mt) {
        const int row = mt * 128 + warp * 32 + (tid & 31);
        for (int c0 = 0; c0 < NT; c0 += 16) {
            uint32_t v[16];
            ptx::tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + mt * NT + c0, v);
            ptx::tmem_ld_wait();
            for (int j = 0; j < 16; ++j) {
                const int n = c0 + j;
                float ref = 0.0f;
                for (int rep = 0; rep < REPS; ++rep)
                    for (int k = 0; k < 64; ++k) ref += a_val(row, k) * b_val(n + rep, k);
                const float got = __uint_as_float(v[j]);
                if (got != ref) ++bad;
                if (blockIdx.x == 0 && mt == 0 && warp == 0 && (tid & 31) < 2 && j < 4 && c0 == 0) sample[(tid & 31) * 8 + j * 2] = got, sample[(tid & 31) * 8 + j * 2 + 1] = ref;
            }
        }
    }
    atomicAdd(mismatches, bad);
    ptx::tc_fence_before();
    asm volatile("bar.sync 1, 128;" ::: "memory");
    if (warp == 0) ptx::tmem_dealloc(tmem, 256);
}

template <int MODE, int LAYOUT, int EXTRA = 0>
int run(const char* name, int grid) {
    long long* cyc;
    int* mis;
    float* sample;
    CK(cudaMalloc(&cyc, 2 * 148 * 8));
    uint4* gsrc;
    CK(cudaMalloc(&gsrc, REPS * 8192));
    CK(cudaMemset(gsrc, 0, REPS * 8192));
    CK(cudaMalloc(&mis, 4));
    CK(cudaMalloc(&sample, 64));
    CK(cudaMemset(mis, 0, 4));
    CK(cudaMemset(sample, 0, 64));
    const int smem = 90112 + REPS * 8192 + 5 * 8192 + 1024;
    CK(cudaFuncSetAttribute(probe<MODE, LAYOUT, EXTRA>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    for (int it = 0; it < 3; ++it) {
        CK(cudaMemset(mis, 0, 4));
        probe<MODE, LAYOUT, EXTRA><<<grid, (EXTRA & 2) ? 640 : 128, smem>>>(cyc, mis, sample, gsrc);
        CK(cudaDeviceSynchronize());
    }
    long long h[148];
    int hm;
    float hs[16];
    CK(cudaMemcpy(h, cyc, grid * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&hm, mis, 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hs, sample, 64, cudaMemcpyDeviceToHost));
    long long mn = h[0], mx = h[0];
    for (int i = 1; i < grid; ++i) { mn = h[i] < mn ? h[i] : mn; mx = h[i] > mx ? h[i] : mx; }
    printf("%-28s grid %3d: %lld..%lld cycles for %d MMAs = %.1f per MMA | mismatches %d of %d | d[0][0..3] got/ref %g/%g %g/%g %g/%g %g/%g\n", name, grid, mn, mx,
           REPS * KB * TILES, (double)mn / (REPS * KB * TILES), hm, grid * TILES * 128 * NT, hs[0], hs[1], hs[2], hs[3], hs[4], hs[5], hs[6], hs[7]);
    cudaFree(cyc); cudaFree(mis); cudaFree(sample);
    return 0;
}

int main() {
    for (int grid : {1, 148}) {
        if (run<0, 0>("plain", grid)) return 1;
        if (run<1, 0>("ws", grid)) return 1;
        if (run<2, 0>("ws + collector b0", grid)) return 1;
        if (run<3, 0>("ws + collectors b0/b1", grid)) return 1;
        if (run<0, 1>("tall A: plain", grid)) return 1;
        if (run<3, 1>("tall A: ws + collectors", grid)) return 1;
        if (run<3, 1, 1>("  + commit per block", grid)) return 1;
        if (run<3, 1, 2>("  + 16 waiting warps", grid)) return 1;
        if (run<3, 1, 3>("  + both", grid)) return 1;
        if (run<3, 1, 4>("  ring protocol", grid)) return 1;
        if (run<3, 1, 12>("  ring protocol + fence", grid)) return 1;
        if (run<3, 1, 14>("  ring + fence + waiters", grid)) return 1;
        if (run<0, 1, 14>("  same, plain MMAs", grid)) return 1;
        if (run<3, 1, 4 + 16>("  ring, arrive instead of TMA", grid)) return 1;
        if (run<3, 1, 4 + 32>("  ring, all blocks requested at once", grid)) return 1;
        if (run<3, 1, 4 + 16 + 32>("  ring, both", grid)) return 1;
        if (run<3, 1, 4 + 64>("  ring, one election per layer", grid)) return 1;
        if (run<3, 1, 4 + 64 + 2>("  same + 16 waiting warps", grid)) return 1;
        if (run<3, 1, 4 + 16 + 32 + 64 + 128>("  one election, no waits", grid)) return 1;
        if (run<3, 1, 4 + 16 + 32 + 64 + 256>("  one election, waits, commits elsewhere", grid)) return 1;
        if (run<3, 1, 4 + 16 + 32 + 64 + 128 + 256>("  one election, neither", grid)) return 1;
        if (run<3, 1, 4 + 16 + 64 + 128>("  one election, no waits, 5-stage wrap", grid)) return 1;
        if (run<3, 1, 4 + 16 + 128>("  election + syncwarp per block, no waits", grid)) return 1;
        if (run<3, 1, 4 + 16 + 128 + 8>("  same + tcgen05 fence per block", grid)) return 1;
    }
    return 0;
}
