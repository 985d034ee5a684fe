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

template <int MODE>
__global__ void __launch_bounds__(128) probe(long long* cycles, int* mismatches, float* sample) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    unsigned char* A = smem;                     // 4 tiles x 128 rows x 128 B
    unsigned char* B = smem + TILES * 128 * 128;  // REPS blocks x 64 rows x 128 B
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < TILES * 128 * 64; i += 128) {
        const int row = i / 64, k = i % 64;
        *reinterpret_cast<__nv_bfloat16*>(A + sw_off(row, k)) = __float2bfloat16(a_val(row, k));
    }
    for (int i = tid; i < REPS * 64 * 64; i += 128) {
        const int blk = i / 4096, n = (i / 64) % 64, k = i % 64;
        *reinterpret_cast<__nv_bfloat16*>(B + blk * 8192 + sw_off(n, k)) = __float2bfloat16(b_val(n + blk, k));
    }
    if (tid == 0) ptx::mbar_init(ptx::smem_u32(&bar), 1);
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
    if (warp == 1) {
        const uint32_t a_hi = ptx::sw128_hi(1024), b_hi = ptx::sw128_hi(1024);
        t0 = clock64();
        if (ptx::elect_one()) {
            for (int rep = 0; rep < REPS; ++rep)
                for (int kk = 0; kk < KB; ++kk) {
                    const uint64_t bdesc = ptx::desc_pack(ptx::sw128_lo(ptx::smem_u32(B + rep * 8192)) + kk * 2, b_hi);
                    for (int mt = 0; mt < TILES; ++mt) {
                        const uint64_t adesc = ptx::desc_pack(ptx::sw128_lo(ptx::smem_u32(A + mt * 16384)) + kk * 2, a_hi);
                        const int tag = MODE == 3 ? (((rep * KB + kk) & 1) << 4) | mt : mt;
                        mma_issue<MODE>(tmem + mt * NT, adesc, bdesc, idesc, (rep | kk) != 0, tag);
                    }
                }
            ptx::mma_commit(ptx::smem_u32(&bar));
        }
        __syncwarp();
        ptx::mbar_wait(ptx::smem_u32(&bar), 0);
        t1 = clock64();
        if (tid == 32) cycles[blockIdx.x] = t1 - t0;
    }
    __syncthreads();
    ptx::mbar_wait(ptx::smem_u32(&bar), 0);
    ptx::tc_fence_after();
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
    __syncthreads();
    if (warp == 0) ptx::tmem_dealloc(tmem, 256);
}

template <int MODE>
int run(const char* name, int grid) {
    long long* cyc;
    int* mis;
    float* sample;
    CK(cudaMalloc(&cyc, 148 * 8));
    CK(cudaMalloc(&mis, 4));
    CK(cudaMalloc(&sample, 64));
    CK(cudaMemset(mis, 0, 4));
    CK(cudaMemset(sample, 0, 64));
    const int smem = TILES * 128 * 128 + REPS * 8192 + 1024;
    CK(cudaFuncSetAttribute(probe<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    for (int it = 0; it < 3; ++it) {
        CK(cudaMemset(mis, 0, 4));
        probe<MODE><<<grid, 128, smem>>>(cyc, mis, sample);
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
        if (run<0>("plain", grid)) return 1;
        if (run<1>("ws", grid)) return 1;
        if (run<2>("ws + collector b0", grid)) return 1;
        if (run<3>("ws + collectors b0/b1", grid)) return 1;
    }
    return 0;
}
