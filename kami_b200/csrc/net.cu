// Policy/value network forward on sm_100a.
//
//   k_conv<N_TILE, SLICE>  implicit-GEMM 3x3 / 1x1 convolution: tcgen05.mma (bf16 x bf16 -> fp32
//                          in TMEM), operands staged by 1-D bulk TMA copies, BN folded into
//                          weights/bias, bias + ReLU (+ skip) fused in the TMEM epilogue.
//   k_value_head           valueconv(1x1) + BN + ReLU + Linear(64,256) + tanh
//   k_softmax              exp(log_softmax) over all 4672 logits
//
// Reference: kami/nn/nn.cpp:26-34 (residual block), :59-91 (NNModule::forward), :155-187
// (NN::infer).  Activations use the tall-image layout of layout.cuh so that every 3x3 tap is
// a descriptor start-address offset and no im2col copy or swizzle is needed.
#include "net.cuh"

#include <cuda_bf16.h>
#include <math.h>
#include <string.h>

#include <atomic>
#include <condition_variable>
#include <mutex>
#include <new>
#include <vector>

#include "common.cuh"
#include "layout.cuh"
#include "ptx.cuh"

namespace kb {

struct ConvParams {
    const uint4* in;    // input activations, slabs_in 64-channel slabs per item
    uint4* out;         // bf16 output activations (swizzled tall image) or nullptr
    const uint4* skip;  // residual input added after the ReLU (nn.cpp:31), or nullptr
    float* out_f32;     // dense fp32 [boards][64][n_valid] output (policy logits) or nullptr
    const uint4* w;     // weights, one contiguous block per (pass, slab, tap)
    const float* bias;  // folded bias, n_total entries
    int items, boards;
    int slabs_in, slabs_out;
    int ksteps;   // 16-channel MMA steps per slab: 4, or 2 for the 30(32)-channel input slab
    int n_total;  // padded output channels (multiple of N_TILE)
    int n_valid;  // real output channels
    int ntaps;    // 9 (3x3, pad 1) or 1 (1x1)
    int relu;
    long long* ts;  // optional profiling hook: wait-cycle counters of CTA 0's MMA thread
    int skew;       // k_conv2: weight blocks issued tile-major at each end of a pass (0..3)
};

template <int N_TILE>
struct ConvCfg {
    static constexpr int MT = 4;                       // 4 M tiles x 16 tall rows = the 64 rows of an item
    static constexpr int A_BYTES = SLAB_BYTES;         // one 64-channel slab of an item
    static constexpr int B_BYTES = N_TILE * LINE_BYTES;  // [n][64 k] swizzled weight block
    static constexpr int NSTAGE = N_TILE > 80 ? 3 : 4;
    static constexpr int ACC_COLS = MT * N_TILE;
    static constexpr int NACC = (2 * ACC_COLS <= 512) ? 2 : 1;
    static constexpr int BAR_BYTES = 2048;             // mbarriers, TMEM slot, folded biases (<= 256 floats at +1024)
    static constexpr int B_STRIDE = (B_BYTES + 1023) / 1024 * 1024;  // stages stay 1024-byte aligned
    static constexpr int STG_BYTES = 4 * 4096;         // epilogue staging: one 32-pixel x 128-byte tile per warp
    static constexpr int SMEM = BAR_BYTES + 2 * A_BYTES + NSTAGE * B_STRIDE + STG_BYTES;
    static_assert(ACC_COLS <= 512, "accumulators must fit TMEM");
    static_assert(SMEM <= 232448, "shared memory budget");
    static_assert(N_TILE % 16 == 0 && N_TILE <= 256, "UMMA N for M=128");
};

// byte offset of tap (dy,dx) for M tile mt inside a slab: first pixel of the tile's first row group
__device__ __forceinline__ int tap_offset(int mt, int dy, int dx) { return ((16 * mt + dy) * TALL_PITCH + 1 + dx) * LINE_BYTES; }

// Warp roles: warp 0 weight (B) producer, warp 1 tcgen05.mma issuer, warp 2 TMEM allocator,
// warp 3 activation (A) producer, warps 4-7 epilogue (one per TMEM lane quarter).
//
// Epilogue data path (bf16 output): a warp owns 32 rows of an M tile = 4 tall rows x 8 pixels.
// For each 64-channel half of the accumulator it builds the 4 row segments (8 pixels x 128 B =
// 1024 contiguous bytes of the output slab, already swizzled) in a 4 KB staging tile and hands
// them to the TMA as bulk shared->global stores; the residual skip input arrives the same way
// (bulk global->shared into the staging tile, added in fp32).  No scattered 16-byte global
// accesses: every global transaction of the kernel is a >= 1 KB bulk copy.
template <int N_TILE>
__global__ void __launch_bounds__(256, 1) k_conv(const ConvParams P) {
    using C = ConvCfg<N_TILE>;
    extern __shared__ __align__(1024) uint8_t smem[];
    const int warp = ptx::uniform_warp_id(), lane = threadIdx.x & 31;
    const uint32_t bar0 = ptx::smem_u32(smem);
    auto b_full = [&](int s) { return bar0 + 8u * s; };
    auto b_empty = [&](int s) { return bar0 + 8u * (4 + s); };
    auto a_full = [&](int s) { return bar0 + 8u * (8 + s); };
    auto a_empty = [&](int s) { return bar0 + 8u * (10 + s); };
    auto t_full = [&](int s) { return bar0 + 8u * (12 + s); };
    auto t_empty = [&](int s) { return bar0 + 8u * (14 + s); };
    auto skip_full = [&](int q) { return bar0 + 8u * (16 + q); };
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 256);
    float* sbias = reinterpret_cast<float*>(smem + 1024);
    const uint32_t a_smem = bar0 + C::BAR_BYTES;
    const uint32_t b_smem = a_smem + 2 * C::A_BYTES;
    constexpr int STG_OFF = C::BAR_BYTES + 2 * C::A_BYTES + C::NSTAGE * C::B_STRIDE;

    const int kslices = P.slabs_in;
    const int npass = P.n_total / N_TILE;
    const int my_items = P.items > (int)blockIdx.x ? (P.items - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    const int jobs_per_item = npass * kslices;
    const int total_jobs = my_items * jobs_per_item;

    if (threadIdx.x == 0) {
        for (int s = 0; s < 4; ++s) {
            ptx::mbar_init(b_full(s), 1);
            ptx::mbar_init(b_empty(s), 1);
            ptx::mbar_init(skip_full(s), 1);
        }
        for (int s = 0; s < 2; ++s) {
            ptx::mbar_init(a_full(s), 1);
            ptx::mbar_init(a_empty(s), 1);
            ptx::mbar_init(t_full(s), 1);
            ptx::mbar_init(t_empty(s), 4);
        }
        ptx::fence_barrier_init();
    }
    if (warp == 2) {
        ptx::tmem_alloc(ptx::smem_u32(tmem_slot), 512);
        ptx::tmem_relinquish();
    }
    for (int i = threadIdx.x; i < P.n_total && i < 256; i += blockDim.x) sbias[i] = P.bias[i];
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 3) {
        // ===== A producer: one bulk copy per (item, pass, slab) job, double buffered =====
        for (int job = 0; job < total_jobs; ++job) {
            const int ii = job / jobs_per_item, rem = job - ii * jobs_per_item;
            const int ks = rem % kslices;
            const int item = (int)blockIdx.x + ii * (int)gridDim.x;
            const int buf = job & 1;
            ptx::mbar_wait(a_empty(buf), ((job >> 1) & 1) ^ 1);
            const uint4* src = P.in + ((size_t)item * P.slabs_in + ks) * SLAB_U4;
            if (ptx::elect_one()) {
                ptx::mbar_arrive_expect_tx(a_full(buf), C::A_BYTES);
                ptx::bulk_g2s(a_smem + buf * C::A_BYTES, src, C::A_BYTES, a_full(buf));
            }
            __syncwarp();
        }
    } else if (warp == 0) {
        // ===== B producer: weight blocks in consumption order through the stage ring =====
        int stage = 0, sphase = 0;
        for (int job = 0; job < total_jobs; ++job) {
            const int rem = job % jobs_per_item;
            const int pass = rem / kslices, ks = rem - pass * kslices;
            const uint4* src = P.w + (size_t)(pass * kslices + ks) * P.ntaps * (C::B_BYTES / 16);
            for (int tap = 0; tap < P.ntaps; ++tap, src += C::B_BYTES / 16) {
                ptx::mbar_wait(b_empty(stage), sphase ^ 1);
                if (ptx::elect_one()) {
                    ptx::mbar_arrive_expect_tx(b_full(stage), C::B_BYTES);
                    ptx::bulk_g2s(b_smem + stage * C::B_STRIDE, src, C::B_BYTES, b_full(stage));
                }
                __syncwarp();
                if (++stage == C::NSTAGE) {
                    stage = 0;
                    sphase ^= 1;
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: converged warp, one elected thread drives the tensor core =====
        constexpr uint32_t idesc = ptx::idesc_bf16(128, N_TILE);
        int stage = 0, sphase = 0, acc_count = 0;
        const bool prof = P.ts && blockIdx.x == 0;
        long long w_t = 0, w_a = 0, w_b = 0, t_begin = prof ? clock64() : 0, c0 = 0;
        for (int job = 0; job < total_jobs; ++job) {
            const int rem = job % jobs_per_item;
            const int ks = rem % kslices;
            const int buf = job & 1;
            const int acc = acc_count % C::NACC;
            if (ks == 0) {
                if (prof) c0 = clock64();
                ptx::mbar_wait(t_empty(acc), ((acc_count / C::NACC) & 1) ^ 1);
                ptx::tc_fence_after();
                if (prof) w_t += clock64() - c0;
            }
            if (prof) c0 = clock64();
            ptx::mbar_wait(a_full(buf), (job >> 1) & 1);
            if (prof) w_a += clock64() - c0;
            // descriptor words: start addresses advance by plain adds (16-byte units)
            const uint32_t a_lo0 = ptx::sw128_lo(a_smem + buf * C::A_BYTES);
            const uint32_t a_hi = ptx::sw128_hi(TALL_PITCH * LINE_BYTES), b_hi = ptx::sw128_hi(1024);
            const uint32_t d_base = tmem_base + acc * C::ACC_COLS;
            const int ksteps = P.ksteps;
            for (int tap = 0; tap < P.ntaps; ++tap) {
                const int dy = P.ntaps == 9 ? tap / 3 - 1 : 0, dx = P.ntaps == 9 ? tap % 3 - 1 : 0;
                const uint32_t a_tap = a_lo0 + (uint32_t)((dy * TALL_PITCH + dx + 1) * (LINE_BYTES / 16));
                if (prof) c0 = clock64();
                ptx::mbar_wait(b_full(stage), sphase);
                ptx::tc_fence_after();
                if (prof) w_b += clock64() - c0;
                const uint32_t b_lo0 = ptx::sw128_lo(b_smem + stage * C::B_STRIDE);
                if (ptx::elect_one()) {
                    uint32_t first = (ks | tap) == 0 ? 0u : 1u;
                    for (int kk = 0; kk < ksteps; ++kk) {
                        const uint64_t bdesc = ptx::desc_pack(b_lo0 + kk * 2, b_hi);
#pragma unroll
                        for (int mt = 0; mt < C::MT; ++mt) {
                            const uint64_t adesc = ptx::desc_pack(a_tap + kk * 2 + mt * (16 * TALL_PITCH * LINE_BYTES / 16), a_hi);
                            ptx::mma_bf16(d_base + mt * N_TILE, adesc, bdesc, idesc, first);
                        }
                        first = 1u;
                    }
                    ptx::mma_commit(b_empty(stage));
                }
                __syncwarp();
                if (++stage == C::NSTAGE) {
                    stage = 0;
                    sphase ^= 1;
                }
            }
            if (ptx::elect_one()) {
                ptx::mma_commit(a_empty(buf));
                if (ks == kslices - 1) ptx::mma_commit(t_full(acc));
            }
            __syncwarp();
            if (ks == kslices - 1) ++acc_count;
        }
        if (prof && lane == 0) {
            P.ts[0] = clock64() - t_begin;
            P.ts[1] = w_t;
            P.ts[2] = w_a;
            P.ts[3] = w_b;
        }
    } else if (warp >= 4) {
        // ===== epilogue: TMEM -> registers -> bias/ReLU/skip -> bf16 (or fp32 logits) -> global =====
        const int q = warp & 3;  // TMEM lane quarter this warp may access
        int acc_count = 0;
        uint8_t* stg = smem + STG_OFF + q * 4096;
        const uint32_t stg_s = ptx::smem_u32(stg);
        uint32_t skip_phase = 0;
        bool store_pending = false;
        const bool eprof = P.ts && blockIdx.x == 0 && q == 0;
        long long e_full = 0, e_store = 0, e_tmem = 0, e_skip = 0, e_math = 0, e_issue = 0, ec = 0;
#define KB_EP(var) do { if (eprof) { const long long now_ = clock64(); var += now_ - ec; ec = now_; } } while (0)
        if (eprof) ec = clock64();
        for (int ii = 0; ii < my_items; ++ii) {
            const int item = (int)blockIdx.x + ii * (int)gridDim.x;
            for (int pass = 0; pass < npass; ++pass, ++acc_count) {
                const int acc = acc_count % C::NACC;
                ptx::mbar_wait(t_full(acc), (acc_count / C::NACC) & 1);
                ptx::tc_fence_after();
                KB_EP(e_full);
#pragma unroll 1
                for (int mt = 0; mt < C::MT; ++mt) {
                    const int r = 32 * q + lane;          // row of the M tile == TMEM lane
                    const int R = 16 * mt + (r >> 3);     // tall row
                    const int x = r & 7;
                    const int slot = (R - 1) / 9, y = (R - 1) - slot * 9;
                    const bool valid = R >= 1 && y < 8 && slot < NB;
                    const int px = R * TALL_PITCH + 1 + x;
                    const int board = item * NB + slot;
                    const uint32_t taddr = tmem_base + ((uint32_t)(32 * q) << 16) + acc * C::ACC_COLS + mt * N_TILE;
                    if (P.out_f32) {
#pragma unroll 1
                        for (int cg = 0; cg < N_TILE / 16; ++cg) {
                            uint32_t v[16];
                            ptx::tmem_ld16(taddr + cg * 16, v);
                            ptx::tmem_ld_wait();
                            if (!valid || board >= P.boards) continue;
                            const int ch0 = pass * N_TILE + cg * 16;
                            float* dst = P.out_f32 + ((size_t)board * 64 + (y * 8 + x)) * P.n_valid;
#pragma unroll
                            for (int j = 0; j < 16; ++j) {
                                float f = __uint_as_float(v[j]) + sbias[ch0 + j];
                                if (P.relu) f = fmaxf(f, 0.0f);
                                if (ch0 + j < P.n_valid) dst[ch0 + j] = f;
                            }
                        }
                        continue;
                    }
                    if constexpr (N_TILE % 64 == 0) {
                        // row segments of this warp: tall rows 16*mt + 4*q + g, g = 0..3 (lanes 8g..8g+7)
                        const int R0 = 16 * mt + 4 * q;
                        auto seg_valid = [&](int g) { const int Rg = R0 + g; return Rg >= 1 && (Rg - 1) % 9 < 8; };
#pragma unroll 1
                        for (int half = 0; half < N_TILE / 64; ++half) {
                            const int ch0 = pass * N_TILE + half * 64;
                            const size_t slab_u4 = ((size_t)item * P.slabs_out + (ch0 >> 6)) * SLAB_U4;
                            if (store_pending) {  // the previous tile's bulk stores must have read the staging tile
                                if (lane == 0) ptx::bulk_wait_read0();
                                __syncwarp();
                                store_pending = false;
                            }
                            KB_EP(e_store);
                            if (P.skip && lane == 0) {
                                int nv = 0;
                                for (int g = 0; g < 4; ++g) nv += seg_valid(g) ? 1 : 0;
                                ptx::mbar_arrive_expect_tx(skip_full(q), 1024u * nv);
                                for (int g = 0; g < 4; ++g)
                                    if (seg_valid(g))
                                        ptx::bulk_g2s(stg_s + g * 1024, P.skip + slab_u4 + (size_t)((R0 + g) * TALL_PITCH + 1) * 8, 1024, skip_full(q));
                            }
                            uint32_t v[4][16];
#pragma unroll
                            for (int g = 0; g < 4; ++g) ptx::tmem_ld16(taddr + half * 64 + g * 16, v[g]);
                            ptx::tmem_ld_wait();
                            if (mt == C::MT - 1 && half == N_TILE / 64 - 1) {  // accumulator drained: release it early
                                ptx::tc_fence_before();
                                __syncwarp();
                                if (lane == 0) ptx::mbar_arrive(t_empty(acc));
                            }
                            KB_EP(e_tmem);
                            if (P.skip) {
                                ptx::mbar_wait(skip_full(q), skip_phase);
                                skip_phase ^= 1;
                            }
                            KB_EP(e_skip);
                            if (valid) {
                                uint4* line = reinterpret_cast<uint4*>(stg + lane * 128);
#pragma unroll
                                for (int j = 0; j < 8; ++j) {  // 8 channels = one 16-byte chunk
                                    float f[8];
                                    const float4 b0 = *reinterpret_cast<const float4*>(sbias + ch0 + 8 * j);
                                    const float4 b1 = *reinterpret_cast<const float4*>(sbias + ch0 + 8 * j + 4);
                                    const uint32_t* vv = &v[j >> 1][(j & 1) * 8];
                                    f[0] = __uint_as_float(vv[0]) + b0.x; f[1] = __uint_as_float(vv[1]) + b0.y;
                                    f[2] = __uint_as_float(vv[2]) + b0.z; f[3] = __uint_as_float(vv[3]) + b0.w;
                                    f[4] = __uint_as_float(vv[4]) + b1.x; f[5] = __uint_as_float(vv[5]) + b1.y;
                                    f[6] = __uint_as_float(vv[6]) + b1.z; f[7] = __uint_as_float(vv[7]) + b1.w;
                                    if (P.relu) {
#pragma unroll
                                        for (int k = 0; k < 8; ++k) f[k] = fmaxf(f[k], 0.0f);
                                    }
                                    uint4* cp = line + (j ^ (px & 7));
                                    if (P.skip) {
                                        const uint4 s4 = *cp;
                                        const __nv_bfloat162* sb = reinterpret_cast<const __nv_bfloat162*>(&s4);
#pragma unroll
                                        for (int k = 0; k < 4; ++k) {
                                            const float2 sv = __bfloat1622float2(sb[k]);
                                            f[2 * k] += sv.x;
                                            f[2 * k + 1] += sv.y;
                                        }
                                    }
                                    uint32_t w4[4];
#pragma unroll
                                    for (int k = 0; k < 4; ++k) {
                                        const __nv_bfloat162 b = __floats2bfloat162_rn(f[2 * k], f[2 * k + 1]);
                                        w4[k] = *reinterpret_cast<const uint32_t*>(&b);
                                    }
                                    *cp = make_uint4(w4[0], w4[1], w4[2], w4[3]);
                                }
                            }
                            // every lane fences, also those without a pixel of their own: the lane that issues the bulk
                            // store must have ordered the tile's generic-proxy writes before the async proxy itself
                            ptx::fence_proxy_async_smem();
                            __syncwarp();
                            KB_EP(e_math);
                            if (lane == 0) {
                                for (int g = 0; g < 4; ++g)
                                    if (seg_valid(g))
                                        ptx::bulk_s2g(P.out + slab_u4 + (size_t)((R0 + g) * TALL_PITCH + 1) * 8, stg_s + g * 1024, 1024);
                                ptx::bulk_commit();
                            }
                            store_pending = true;
                            KB_EP(e_issue);
                        }
                    }
                }
                if (P.out_f32 || N_TILE % 64 != 0) {
                    ptx::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive(t_empty(acc));
                }
            }
        }
        if (lane == 0) ptx::bulk_wait0();  // shared memory must outlive the last bulk stores
        if (eprof && lane == 0) {
            P.ts[4] = e_full; P.ts[5] = e_store; P.ts[6] = e_tmem; P.ts[7] = e_skip; P.ts[8] = e_math; P.ts[9] = e_issue;
        }
#undef KB_EP
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 2) ptx::tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------
// k_conv2: the same implicit-GEMM convolution on CTA PAIRS (cluster of 2, tcgen05 cta_group::2).
// Each CTA of a pair owns one item (its A slabs and its 128-lane accumulators); the even CTA
// issues M = 256 MMAs that read A from both CTAs and HALF of every weight block from each CTA,
// so an SM streams and re-reads only N_TILE/2 weight rows per tap.  The 128x128x16 single-CTA
// MMA needs 128 B/clk of shared-memory operand reads (the whole port); the paired form needs 96
// and halves the L2->SM weight traffic.  Eight epilogue warps (two per TMEM lane quarter).
//
// Barrier protocol: full barriers live in the even (leader) CTA with two arrivals -- its own
// arrive.expect_tx and a forwarded arrive from the odd CTA (warp 1 of the odd CTA waits for the
// local TMA completion and arrives remotely); empty / accumulator-full barriers are signalled in
// both CTAs by multicast tcgen05.commit; accumulator-empty collects all 16 epilogue warps.
// ------------------------------------------------------------------------------------------
template <int N_TILE>
struct Conv2Cfg {
    static constexpr int MT = 4;
    static constexpr int HALVES = N_TILE / 64;
    static constexpr int A_BYTES = SLAB_BYTES;
    static constexpr int B_HALF = N_TILE * LINE_BYTES / 2;  // this CTA's rows of a [n][64 k] weight block
    static constexpr int NSTAGE = 4;
    static constexpr int ACC_COLS = MT * N_TILE;
    static constexpr int NACC = (2 * ACC_COLS <= 512) ? 2 : 1;
    static constexpr int BAR_BYTES = 2048;
    static constexpr int B_STRIDE = (B_HALF + 1023) / 1024 * 1024;
    static constexpr int STG_BYTES = 8 * 4096;
    static constexpr int SMEM = BAR_BYTES + 2 * A_BYTES + NSTAGE * B_STRIDE + STG_BYTES;
    static_assert(N_TILE == 128 && ACC_COLS == 512, "one pass fills TMEM; two channel halves per tile");
    static_assert(SMEM <= 232448, "shared memory budget");
};

template <int N_TILE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(384, 1) k_conv2(const ConvParams P) {
    using C = Conv2Cfg<N_TILE>;
    extern __shared__ __align__(1024) uint8_t smem[];
    const int warp = ptx::uniform_warp_id(), lane = threadIdx.x & 31;
    const uint32_t rank = ptx::cluster_ctarank();
    const uint32_t bar0 = ptx::smem_u32(smem);
    auto b_full = [&](int s) { return bar0 + 8u * s; };
    auto b_empty = [&](int s) { return bar0 + 8u * (4 + s); };
    auto a_full = [&](int s) { return bar0 + 8u * (8 + s); };
    auto a_empty = [&](int s) { return bar0 + 8u * (10 + s); };
    auto t_full = [&](int mt) { return bar0 + 8u * (12 + mt); };   // accumulator of M tile mt complete (both CTAs)
    auto t_empty = [&](int mt) { return bar0 + 8u * (16 + mt); };  // ... drained by all 16 epilogue warps of the pair
    auto skip_full = [&](int e) { return bar0 + 8u * (20 + e); };
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 256);
    float* sbias = reinterpret_cast<float*>(smem + 1024);
    const uint32_t a_smem = bar0 + C::BAR_BYTES;
    const uint32_t b_smem = a_smem + 2 * C::A_BYTES;
    constexpr int STG_OFF = C::BAR_BYTES + 2 * C::A_BYTES + C::NSTAGE * C::B_STRIDE;

    const int kslices = P.slabs_in;
    const int npass = P.n_total / N_TILE;
    const int pair0 = (int)blockIdx.x >> 1, pair_stride = (int)gridDim.x >> 1;
    const int total_pairs = (P.items + 1) >> 1;
    const int my_pairs = total_pairs > pair0 ? (total_pairs - 1 - pair0) / pair_stride + 1 : 0;
    const int jobs_per_item = npass * kslices;
    const int total_jobs = my_pairs * jobs_per_item;

    if (threadIdx.x == 0) {
        const uint32_t full_count = rank == 0 ? 2u : 1u;
        for (int s = 0; s < 4; ++s) {
            ptx::mbar_init(b_full(s), full_count);
            ptx::mbar_init(b_empty(s), 1);
        }
        for (int s = 0; s < 2; ++s) {
            ptx::mbar_init(a_full(s), full_count);
            ptx::mbar_init(a_empty(s), 1);
        }
        for (int mt = 0; mt < C::MT; ++mt) {
            ptx::mbar_init(t_full(mt), 1);
            ptx::mbar_init(t_empty(mt), 16);
        }
        for (int e = 0; e < 8; ++e) ptx::mbar_init(skip_full(e), 1);
        ptx::fence_barrier_init();
    }
    if (warp == 2) {
        ptx::tmem_alloc2(ptx::smem_u32(tmem_slot), 512);
        ptx::tmem_relinquish2();
    }
    for (int i = threadIdx.x; i < P.n_total && i < 256; i += blockDim.x) sbias[i] = P.bias[i];
    ptx::tc_fence_before();
    __syncthreads();
    ptx::cluster_sync();  // both CTAs' barriers are initialised before any remote arrive / multicast commit
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 3) {
        // ===== A producer: this CTA's item, one bulk copy per (pair, pass, slab) job =====
        for (int job = 0; job < total_jobs; ++job) {
            const int ii = job / jobs_per_item, rem = job - ii * jobs_per_item;
            const int ks = rem % kslices;
            int item = 2 * (pair0 + ii * pair_stride) + (int)rank;
            if (item >= P.items) item = P.items - 1;  // odd item count: the last pair's odd CTA recomputes a live item, writes nothing
            const int buf = job & 1;
            ptx::mbar_wait(a_empty(buf), ((job >> 1) & 1) ^ 1);
            const uint4* src = P.in + ((size_t)item * P.slabs_in + ks) * SLAB_U4;
            if (ptx::elect_one()) {
                ptx::mbar_arrive_expect_tx(a_full(buf), C::A_BYTES);
                ptx::bulk_g2s(a_smem + buf * C::A_BYTES, src, C::A_BYTES, a_full(buf));
            }
            __syncwarp();
        }
    } else if (warp == 0) {
        // ===== B producer: this CTA's half (N_TILE/2 rows) of every weight block =====
        int stage = 0, sphase = 0;
        for (int job = 0; job < total_jobs; ++job) {
            const int rem = job % jobs_per_item;
            const int pass = rem / kslices, ks = rem - pass * kslices;
            const uint4* src = P.w + (size_t)(pass * kslices + ks) * P.ntaps * (N_TILE * LINE_BYTES / 16) + rank * (C::B_HALF / 16);
            for (int tap = 0; tap < P.ntaps; ++tap, src += N_TILE * LINE_BYTES / 16) {
                ptx::mbar_wait(b_empty(stage), sphase ^ 1);
                if (ptx::elect_one()) {
                    ptx::mbar_arrive_expect_tx(b_full(stage), C::B_HALF);
                    ptx::bulk_g2s(b_smem + stage * C::B_STRIDE, src, C::B_HALF, b_full(stage));
                }
                __syncwarp();
                if (++stage == C::NSTAGE) {
                    stage = 0;
                    sphase ^= 1;
                }
            }
        }
    } else if (warp == 1 && rank == 1) {
        // ===== odd CTA: forward local TMA completions to the leader's full barriers =====
        int stage = 0, sphase = 0;
        for (int job = 0; job < total_jobs; ++job) {
            const int buf = job & 1;
            ptx::mbar_wait(a_full(buf), (job >> 1) & 1);
            if (ptx::elect_one()) ptx::mbar_arrive_cluster(ptx::mapa(a_full(buf), 0));
            __syncwarp();
            for (int tap = 0; tap < P.ntaps; ++tap) {
                ptx::mbar_wait(b_full(stage), sphase);
                if (ptx::elect_one()) ptx::mbar_arrive_cluster(ptx::mapa(b_full(stage), 0));
                __syncwarp();
                if (++stage == C::NSTAGE) {
                    stage = 0;
                    sphase ^= 1;
                }
            }
        }
    } else if (warp == 1) {
        // ===== leader MMA issuer: one elected thread drives both SMs' tensor cores =====
        // A pass is U = slabs x taps weight blocks of 4 k-steps for each of the 4 M tiles.  TMEM holds
        // exactly one pass (4 x 128 columns), so the accumulators are handed over tile by tile: the
        // first and the last SKEW blocks of a pass are issued tile-major (all blocks for tile 0, then
        // tile 1, ...), the middle block-major.  Tile mt then completes 3*(3-mt) block-times before the
        // pass ends and is first needed 3*mt block-times after the next pass starts, which is the
        // window the epilogue warps drain it in.
        constexpr uint32_t idesc = ptx::idesc_bf16(256, N_TILE);
        constexpr int SKEW = 3;  // <= NSTAGE - 1 weight stages stay resident during a tile-major group
        const int ntaps = P.ntaps, ksteps = P.ksteps;
        const int U = kslices * ntaps;
        const int G = U / 2 < P.skew ? U / 2 : (P.skew < SKEW ? P.skew : SKEW);
        const int total_passes = my_pairs * npass;
        const uint32_t a_hi = ptx::sw128_hi(TALL_PITCH * LINE_BYTES), b_hi = ptx::sw128_hi(1024);
        int gb_ready = 0, gj_ready = 0;  // weight blocks / activation slabs already waited for (global counters)
        const bool prof = P.ts && blockIdx.x == 0;
        long long w_t = 0, w_a = 0, w_b = 0, t_begin = prof ? clock64() : 0, c0 = 0;
        uint32_t a_tap = 0, b_lo0 = 0;
        auto open_block = [&](int pi, int u) {  // wait for block u's operands, set up its descriptor bases
            const int ks = ntaps == 9 ? u / 9 : u, tap = u - ks * ntaps;
            const int gj = pi * kslices + ks, gb = pi * U + u;
            if (gj_ready <= gj) {
                if (prof) c0 = clock64();
                for (; gj_ready <= gj; ++gj_ready) ptx::mbar_wait(a_full(gj_ready & 1), (gj_ready >> 1) & 1);
                if (prof) w_a += clock64() - c0;
            }
            if (gb_ready <= gb) {
                if (prof) c0 = clock64();
                for (; gb_ready <= gb; ++gb_ready) ptx::mbar_wait(b_full(gb_ready % C::NSTAGE), (gb_ready / C::NSTAGE) & 1);
                ptx::tc_fence_after();
                if (prof) w_b += clock64() - c0;
            }
            const int dy = ntaps == 9 ? tap / 3 - 1 : 0, dx = ntaps == 9 ? tap - (tap / 3) * 3 - 1 : 0;
            a_tap = ptx::sw128_lo(a_smem + (gj & 1) * C::A_BYTES) + (uint32_t)((dy * TALL_PITCH + dx + 1) * (LINE_BYTES / 16));
            b_lo0 = ptx::sw128_lo(b_smem + (gb % C::NSTAGE) * C::B_STRIDE);
        };
        constexpr uint32_t MT_STEP = 16 * TALL_PITCH * LINE_BYTES / 16;  // descriptor start-address units between M tiles
        auto mma_block_all = [&](int pi, int u) {  // block-major: every k-step feeds the four tiles
            open_block(pi, u);
            if (ptx::elect_one()) {
                uint32_t acc = u == 0 ? 0u : 1u;
                for (int kk = 0; kk < ksteps; ++kk) {
                    const uint64_t bdesc = ptx::desc_pack(b_lo0 + kk * 2, b_hi);
#pragma unroll
                    for (int mt = 0; mt < C::MT; ++mt)
                        ptx::mma_bf16_2cta(tmem_base + mt * N_TILE, ptx::desc_pack(a_tap + kk * 2 + mt * MT_STEP, a_hi), bdesc, idesc, acc);
                    acc = 1u;
                }
            }
            __syncwarp();
        };
        auto mma_block_one = [&](int pi, int u, int mt) {  // tile-major: one tile's k-steps of block u
            open_block(pi, u);
            const uint32_t d = tmem_base + mt * N_TILE, a0 = a_tap + mt * MT_STEP;
            if (ptx::elect_one()) {
                uint32_t acc = u == 0 ? 0u : 1u;
#pragma unroll 2
                for (int kk = 0; kk < ksteps; ++kk) {
                    ptx::mma_bf16_2cta(d, ptx::desc_pack(a0 + kk * 2, a_hi), ptx::desc_pack(b_lo0 + kk * 2, b_hi), idesc, acc);
                    acc = 1u;
                }
            }
            __syncwarp();
        };
        auto release_block = [&](int pi, int u) {  // every tile has consumed block u: free its stage (and slab)
            const int ks = ntaps == 9 ? u / 9 : u;
            if (ptx::elect_one()) {
                ptx::mma_commit_2cta(b_empty((pi * U + u) % C::NSTAGE));
                if (u + 1 == (ks + 1) * ntaps) ptx::mma_commit_2cta(a_empty((pi * kslices + ks) & 1));
            }
            __syncwarp();
        };
        for (int pi = 0; pi < total_passes; ++pi) {
            for (int mt = 0; mt < C::MT; ++mt) {  // head: tile-major, each tile waits for its own drain
                if (prof) c0 = clock64();
                ptx::mbar_wait(t_empty(mt), (pi & 1) ^ 1);
                ptx::tc_fence_after();
                if (prof) w_t += clock64() - c0;
                for (int u = 0; u < G; ++u) mma_block_one(pi, u, mt);
            }
            for (int u = 0; u < G; ++u) release_block(pi, u);
            for (int u = G; u < U - G; ++u) {     // middle: block-major
                mma_block_all(pi, u);
                release_block(pi, u);
            }
            for (int mt = 0; mt < C::MT; ++mt) {  // tail: tile-major, each tile is handed over as it completes
                for (int u = U - G; u < U; ++u) mma_block_one(pi, u, mt);
                if (ptx::elect_one()) ptx::mma_commit_2cta(t_full(mt));
                __syncwarp();
            }
            for (int u = U - G; u < U; ++u) release_block(pi, u);
        }
        if (prof && lane == 0) {
            P.ts[0] = clock64() - t_begin;
            P.ts[1] = w_t;
            P.ts[2] = w_a;
            P.ts[3] = w_b;
        }
    } else if (warp >= 4) {
        // ===== epilogue: 8 warps; warp (q, h) takes the steps (mt, half) with (mt*HALVES + half) % 2 == h =====
        const int e = warp - 4, q = warp & 3, h = e >> 2;
        int acc_count = 0;
        uint8_t* stg = smem + STG_OFF + e * 4096;
        const uint32_t stg_s = ptx::smem_u32(stg);
        uint32_t skip_phase = 0;
        const bool eprof = P.ts && blockIdx.x == 0 && e == 0;
        long long e_full = 0, e_store = 0, e_tmem = 0, e_skip = 0, e_math = 0, e_issue = 0, ec = 0;
#define KB_EP(var) do { if (eprof) { const long long now_ = clock64(); var += now_ - ec; ec = now_; } } while (0)
        if (eprof) ec = clock64();
        for (int ii = 0; ii < my_pairs; ++ii) {
            const int item = 2 * (pair0 + ii * pair_stride) + (int)rank;
            const bool live = item < P.items;
            for (int pass = 0; pass < npass; ++pass, ++acc_count) {
#pragma unroll 1
                for (int mt = 0; mt < C::MT; ++mt) {  // this warp: channel half h of every M tile
                    const int half = h;
                    ptx::mbar_wait(t_full(mt), acc_count & 1);
                    ptx::tc_fence_after();
                    KB_EP(e_full);
                    const int r = 32 * q + lane;          // row of the M tile == TMEM lane
                    const int R = 16 * mt + (r >> 3);     // tall row
                    const int x = r & 7;
                    const bool valid = live && R >= 1 && (R - 1) % 9 < 8;
                    const int px = R * TALL_PITCH + 1 + x;
                    // row segment g (lanes 8g..8g+7) = tall row 16*mt + 4*q + g; lane g < 4 moves it
                    const int Rg = 16 * mt + 4 * q + lane;
                    const bool mover = lane < 4 && live && Rg >= 1 && (Rg - 1) % 9 < 8;
                    const int nv = __popc(__ballot_sync(0xffffffffu, mover));
                    const int ch0 = pass * N_TILE + half * 64;
                    const size_t seg_u4 = ((size_t)item * P.slabs_out + (ch0 >> 6)) * SLAB_U4 + (size_t)(Rg * TALL_PITCH + 1) * 8;
                    const uint32_t taddr = tmem_base + ((uint32_t)(32 * q) << 16) + mt * N_TILE + half * 64;
                    if (lane < 4) ptx::bulk_wait_read0();  // the previous step's stores have read the staging tile
                    __syncwarp();
                    KB_EP(e_store);
                    if (P.skip && nv) {
                        if (lane == 0) ptx::mbar_arrive_expect_tx(skip_full(e), 1024u * nv);
                        __syncwarp();
                        if (mover) ptx::bulk_g2s(stg_s + lane * 1024, P.skip + seg_u4, 1024, skip_full(e));
                    }
                    uint32_t v[4][16];
#pragma unroll
                    for (int g = 0; g < 4; ++g) ptx::tmem_ld16(taddr + g * 16, v[g]);
                    ptx::tmem_ld_wait();
                    ptx::tc_fence_before();  // this warp's share of tile mt is in registers: hand it back
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive_cluster(ptx::mapa(t_empty(mt), 0));
                    KB_EP(e_tmem);
                    if (P.skip && nv) {
                        ptx::mbar_wait(skip_full(e), skip_phase);
                        skip_phase ^= 1;
                    }
                    KB_EP(e_skip);
                    if (valid) {
                        uint4* line = reinterpret_cast<uint4*>(stg + lane * 128);
                        const int sw = px & 7;
#pragma unroll
                        for (int g = 0; g < 4; ++g) {  // 16 channels = two 16-byte chunks per TMEM column group
                            float4 bb[4];
#pragma unroll
                            for (int k = 0; k < 4; ++k) bb[k] = *reinterpret_cast<const float4*>(sbias + ch0 + 16 * g + 4 * k);
                            const float* bf = reinterpret_cast<const float*>(bb);
#pragma unroll
                            for (int c = 0; c < 2; ++c) {
                                uint4* cp = line + ((2 * g + c) ^ sw);
                                uint32_t w4[4];
                                if (P.skip) {  // x = skip + relu(conv + bias): ReLU before the add, none after (nn.cpp:31)
                                    const uint4 s4 = *cp;
                                    const uint32_t sk[4] = {s4.x, s4.y, s4.z, s4.w};
                                    const float floor_ = P.relu ? 0.0f : -INFINITY;  // relu == 0: plain conv + skip (training passes)
#pragma unroll
                                    for (int k = 0; k < 4; ++k) {
                                        const int i0 = 8 * c + 2 * k;
                                        const float lo = fmaxf(__uint_as_float(v[g][i0]) + bf[i0], floor_) + __uint_as_float(sk[k] << 16);
                                        const float hi = fmaxf(__uint_as_float(v[g][i0 + 1]) + bf[i0 + 1], floor_) + __uint_as_float(sk[k] & 0xffff0000u);
                                        w4[k] = ptx::pack_bf16x2(lo, hi);
                                    }
                                } else if (!P.relu) {
#pragma unroll
                                    for (int k = 0; k < 4; ++k) {
                                        const int i0 = 8 * c + 2 * k;
                                        w4[k] = ptx::pack_bf16x2(__uint_as_float(v[g][i0]) + bf[i0], __uint_as_float(v[g][i0 + 1]) + bf[i0 + 1]);
                                    }
                                } else {
#pragma unroll
                                    for (int k = 0; k < 4; ++k) {
                                        const int i0 = 8 * c + 2 * k;
                                        w4[k] = ptx::pack_relu_bf16x2(__uint_as_float(v[g][i0]) + bf[i0], __uint_as_float(v[g][i0 + 1]) + bf[i0 + 1]);
                                    }
                                }
                                *cp = make_uint4(w4[0], w4[1], w4[2], w4[3]);
                            }
                        }
                    }
                    // every lane fences, also those without a pixel of their own (pad rows): a mover lane must have
                    // ordered the tile's generic-proxy writes before the async proxy itself (lanes 1-3 move row
                    // segments 1-3 but own pixels of segment 0)
                    ptx::fence_proxy_async_smem();
                    __syncwarp();
                    ptx::fence_proxy_async_smem();
                    KB_EP(e_math);
                    if (mover) {
                        ptx::bulk_s2g(P.out + seg_u4, stg_s + lane * 1024, 1024);
                        ptx::bulk_commit();
                    }
                    KB_EP(e_issue);
                }
            }
        }
        if (lane < 4) ptx::bulk_wait0();  // shared memory must outlive the last bulk stores
        if (eprof && lane == 0) {
            P.ts[4] = e_full; P.ts[5] = e_store; P.ts[6] = e_tmem; P.ts[7] = e_skip; P.ts[8] = e_math; P.ts[9] = e_issue;
        }
#undef KB_EP
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::cluster_sync();  // the peer may still read this CTA's shared memory / signal its barriers
    if (warp == 2) ptx::tmem_dealloc2(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------
// Fused forward for the 64-filter network (options.def.yml: filters 64): ONE kernel per batch.
// A CTA owns an item (7 boards).  Activations never leave shared memory between layers:
//
//   region  = [ X : slab 0 | Y : slab 1 ]   (2 x 80 KB swizzled tall-image slabs, pads zero)
//   P (input slab) lands in Y; conv1: P -> X; residual r: X -> Y -> X (+skip from X);
//   value head (CUDA cores) reads X; policy conv 1x1: X -> H (128 ch = both slabs);
//   policy conv2 1x1: H -> fp32 logits (overlay), softmax over 4672 per board -> global.
//
// Warp roles: warp 0 bulk-TMA producer (input slab + all weight blocks in consumption order
// through a 5-stage ring); warp 1 tcgen05.mma issuer -- ONE lane issues a whole layer as a straight
// line of weight-stationary MMAs (tower_issue_3x3_n64 / tower_issue_1x1 below; the general loop
// serves any other layer shape); warp 2 TMEM allocator, warps 2-3 the value head's Linear + tanh;
// warps 4.. the EW epilogue warps (EW = 16: one M tile of a layer per warp and TMEM lane quarter;
// EW = 8: two tiles per warp).
// ------------------------------------------------------------------------------------------
struct FusedLayer {
    int src_off, dst_off;  // byte offsets of the source / destination slabs inside the region
    int slabs_in;          // 64-channel slabs read
    int ksteps;            // 16-channel MMA steps per slab
    int n;                 // output channels of the layer (accumulator columns per M tile)
    int n_sub;             // UMMA N of one weight block (n = nsub * n_sub)
    int ntaps, relu, skip, bias_off;
    int kind;              // 0 bf16 activation to smem, 1 fp32 logits to smem
    uint32_t idesc;
};
struct FusedParams {
    const uint4* planes;
    const uint4* w;
    const float* bias;
    float* policy;
    float* value256;
    const float* wv;
    const float* fct;
    const float* fcb;
    int* nan_flag;
    long long* ts;  // optional per-phase clock64 stamps of CTA 0 (profiling hook), or nullptr
    // legal-gather mode (pool step): per-board action lists instead of the dense softmax
    const uint8_t* legal_act;
    const uint8_t* legal_n;
    unsigned long long legal_stride;
    float* prior;
    // split select (tree.cuh PoolDev::sync): the producing kernel is still running when this one starts; sync[0] reaches
    // sync_target when every board's planes are written, sync[1] when every legal-move list is.  nullptr: ordinary
    // programmatic dependent launch (griddepcontrol.wait = the producing kernel has completed).
    const unsigned* sync;
    unsigned sync_target;
    float bv;
    int n_bias, items, boards, n_layers, tower_layers;
    int wv_off;  // the value conv's 64 folded weights sit behind the biases in shared memory (float offset, multiple of 4)
    int mma_ws;  // 1: N = 64 / 128 layers issue tcgen05.mma.ws with the weight block held in a collector buffer (KB_TOWER_WS)
    int blocks_before_logits, blocks_per_item;  // weight-ring position of policyconv2's two blocks inside an item
    int gather;  // 1: in legal-move mode the logits of the legal moves are computed where they are needed, from the policy head's
                 //    features and policyconv2's rows in the weight ring, instead of all 4 672 by MMA (KB_TOWER_GATHER=1; measured
                 //    62.0 k vs 64.0 k cycles per item on young games, 69.9 vs 70.1 us per step on aged ones: off by default)
    int wait_group;  // 1: every layer through the general issue loop (one ring wait per weight block); otherwise the 3x3 / 1x1 layers
                     //    of the standard network take the straight-line routines and the general loop waits for up to this many
                     //    blocks at a time (KB_TOWER_WAIT_GROUP, default 3)
    FusedLayer layer[16];
};
#ifndef KB_TOWER_PIPE_DEFAULT
#define KB_TOWER_PIPE_DEFAULT 0
#endif
constexpr int FZ_HDR = 8192;                  // barriers, tmem slot, value scratch, biases
constexpr int FZ_REGION = 2 * SLAB_BYTES;     // 163840
constexpr int FZ_STAGE = 10240;               // one weight block: n_sub x 128 B <= 10 KB (80 rows)
constexpr int FZ_NSTAGE = 5;                  // 50 KB of weights in flight
constexpr int FZ_SMEM = FZ_HDR + FZ_REGION + FZ_NSTAGE * FZ_STAGE;
// EW epilogue warps (8 or 16).  With 16, every warp of a TMEM lane quarter owns ONE M tile of a layer instead of two:
// the epilogue phases, which alternate with the MMA phases (layer l + 1 needs all of layer l), halve.
static_assert(FZ_SMEM <= 232448, "fused tower shared memory budget");
static_assert(FZ_STAGE % 1024 == 0 && FZ_HDR % 1024 == 0, "swizzled operands need 1024-byte aligned bases");

// Straight-line issue of one 3x3 layer with 64 output channels (nine weight blocks of KSTEPS K steps, four M tiles),
// by ONE lane.  The tensor pipe's issue queue is short: anything but uniform-register arithmetic between two MMAs shows
// as idle tensor cycles (tools/mma_ws_probe.cu, cycles per 128x64x16 MMA: 43.8 issued straight; 51.0 with an election +
// __syncwarp per 16; 55.4 with a wait per 16 -- an mbarrier try_wait that succeeds at once, even a plain shared-memory
// load).  So every descriptor of the layer is a compile-time offset from values computed before the first MMA, and the
// lane looks at the weight ring three times per layer instead of nine: before block 0 (blocks 0..4: the whole ring, loaded
// during the previous epilogue), before block 5 (5, 6) and before block 7 (7, 8) -- block j + 5 is requested when block
// j's MMAs retire, about 1.3 k cycles before it is due here.
template <int KSTEPS, bool WS>
__device__ __forceinline__ void tower_issue_3x3_n64(uint32_t a_lo0, uint32_t a_hi, uint32_t b_hi, uint32_t d_base, uint32_t idesc, uint32_t ring_lo,
                                                    uint32_t bar_full0, uint32_t bar_empty0, int stage, int phase, uint32_t bar_t_full,
                                                    uint32_t bar_in, uint32_t par_in) {
    constexpr uint32_t MT_STEP = 16 * TALL_PITCH * LINE_BYTES / 16;
    constexpr int N = 64;
    static_assert(FZ_NSTAGE == 5, "wait schedule below assumes a 5-stage ring");
    uint32_t b_lo[9], full[9], empty[9], par[9];
#pragma unroll
    for (int j = 0; j < 9; ++j) {
        const int sj = stage + j;
        const int wrap = sj >= 2 * FZ_NSTAGE ? 2 : sj >= FZ_NSTAGE ? 1 : 0;  // (stage <= 4, j <= 8: up to two laps)
        const int sg = sj - wrap * FZ_NSTAGE;
        b_lo[j] = ring_lo + (uint32_t)sg * (FZ_STAGE / 16);
        full[j] = bar_full0 + 8u * sg;
        empty[j] = bar_empty0 + 8u * sg;
        par[j] = (uint32_t)(phase ^ (wrap & 1));
    }
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
        if (tap == 0) {
#pragma unroll
            for (int j = 0; j < 5; ++j) ptx::mbar_wait(full[j], par[j]);
            // the layer's input last: the previous layer written to shared memory and its accumulators drained (or the
            // planes landed) -- everything else this lane needs is already in registers when that happens
            ptx::mbar_wait(bar_in, par_in);
            ptx::tc_fence_after();
        } else if (tap == 5) {
            ptx::mbar_wait(full[5], par[5]);
            ptx::mbar_wait(full[6], par[6]);
        } else if (tap == 7) {
            ptx::mbar_wait(full[7], par[7]);
            ptx::mbar_wait(full[8], par[8]);
        }
        const int dy = tap / 3 - 1, dx = tap % 3 - 1;
        const uint32_t a_tap = a_lo0 + (uint32_t)((dy * TALL_PITCH + dx + 1) * (LINE_BYTES / 16));
#pragma unroll
        for (int kk = 0; kk < KSTEPS; ++kk) {
            const uint64_t bdesc = ptx::desc_pack(b_lo[tap] + kk * 2, b_hi);
            const uint32_t a0 = a_tap + kk * 2;
            const uint32_t acc = (tap | kk) != 0 ? 1u : 0u;
            if constexpr (WS) {
                // the weight block of a K step stays in a collector buffer for its four M tiles (buffers alternate so the
                // next fill does not wait for the last use)
                if (kk & 1) {
                    ptx::mma_bf16_ws<1, 0>(d_base, ptx::desc_pack(a0, a_hi), bdesc, idesc, acc);
                    ptx::mma_bf16_ws<1, 1>(d_base + N, ptx::desc_pack(a0 + MT_STEP, a_hi), bdesc, idesc, acc);
                    ptx::mma_bf16_ws<1, 1>(d_base + 2 * N, ptx::desc_pack(a0 + 2 * MT_STEP, a_hi), bdesc, idesc, acc);
                    ptx::mma_bf16_ws<1, 2>(d_base + 3 * N, ptx::desc_pack(a0 + 3 * MT_STEP, a_hi), bdesc, idesc, acc);
                } else {
                    ptx::mma_bf16_ws<0, 0>(d_base, ptx::desc_pack(a0, a_hi), bdesc, idesc, acc);
                    ptx::mma_bf16_ws<0, 1>(d_base + N, ptx::desc_pack(a0 + MT_STEP, a_hi), bdesc, idesc, acc);
                    ptx::mma_bf16_ws<0, 1>(d_base + 2 * N, ptx::desc_pack(a0 + 2 * MT_STEP, a_hi), bdesc, idesc, acc);
                    ptx::mma_bf16_ws<0, 2>(d_base + 3 * N, ptx::desc_pack(a0 + 3 * MT_STEP, a_hi), bdesc, idesc, acc);
                }
            } else {
#pragma unroll
                for (int mt = 0; mt < 4; ++mt) ptx::mma_bf16(d_base + mt * N, ptx::desc_pack(a0 + mt * MT_STEP, a_hi), bdesc, idesc, acc);
            }
        }
        ptx::mma_commit(empty[tap]);
    }
    ptx::mma_commit(bar_t_full);
}

// The same for the 1x1 head layers: NSUB x SLABS weight blocks (sub-block major), all waited for up front (they were
// loaded during the previous epilogue), 64-channel K slabs.
template <int NSUB, int SLABS, int N_SUB, bool WS>
__device__ __forceinline__ void tower_issue_1x1(uint32_t a_lo0, uint32_t a_hi, uint32_t b_hi, uint32_t d_base, uint32_t idesc, uint32_t ring_lo,
                                                uint32_t bar_full0, uint32_t bar_empty0, int stage, int phase, uint32_t bar_t_full, uint32_t bar_in,
                                                uint32_t par_in) {
    constexpr uint32_t MT_STEP = 16 * TALL_PITCH * LINE_BYTES / 16;
    constexpr int NBLK = NSUB * SLABS, N = NSUB * N_SUB;
    static_assert(NBLK <= FZ_NSTAGE, "all blocks of the layer must fit the ring");
    uint32_t b_lo[NBLK], empty[NBLK];
#pragma unroll
    for (int j = 0; j < NBLK; ++j) {
        const int sj = stage + j;
        const int wrap = sj >= FZ_NSTAGE ? 1 : 0;
        const int sg = sj - wrap * FZ_NSTAGE;
        b_lo[j] = ring_lo + (uint32_t)sg * (FZ_STAGE / 16);
        empty[j] = bar_empty0 + 8u * sg;
        ptx::mbar_wait(bar_full0 + 8u * sg, (uint32_t)(phase ^ wrap));
    }
    ptx::mbar_wait(bar_in, par_in);
    ptx::tc_fence_after();
#pragma unroll
    for (int sub = 0; sub < NSUB; ++sub)
#pragma unroll
        for (int ks = 0; ks < SLABS; ++ks) {
            const int j = sub * SLABS + ks;
            const uint32_t a_tap = a_lo0 + (uint32_t)ks * (SLAB_BYTES / 16) + (uint32_t)(LINE_BYTES / 16);  // centre tap: one pixel in
            const uint32_t d = d_base + sub * N_SUB;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                const uint64_t bdesc = ptx::desc_pack(b_lo[j] + kk * 2, b_hi);
                const uint32_t a0 = a_tap + kk * 2;
                const uint32_t acc = (ks | kk) != 0 ? 1u : 0u;
                if constexpr (WS) {
                    if (kk & 1) {
                        ptx::mma_bf16_ws<1, 0>(d, ptx::desc_pack(a0, a_hi), bdesc, idesc, acc);
                        ptx::mma_bf16_ws<1, 1>(d + N, ptx::desc_pack(a0 + MT_STEP, a_hi), bdesc, idesc, acc);
                        ptx::mma_bf16_ws<1, 1>(d + 2 * N, ptx::desc_pack(a0 + 2 * MT_STEP, a_hi), bdesc, idesc, acc);
                        ptx::mma_bf16_ws<1, 2>(d + 3 * N, ptx::desc_pack(a0 + 3 * MT_STEP, a_hi), bdesc, idesc, acc);
                    } else {
                        ptx::mma_bf16_ws<0, 0>(d, ptx::desc_pack(a0, a_hi), bdesc, idesc, acc);
                        ptx::mma_bf16_ws<0, 1>(d + N, ptx::desc_pack(a0 + MT_STEP, a_hi), bdesc, idesc, acc);
                        ptx::mma_bf16_ws<0, 1>(d + 2 * N, ptx::desc_pack(a0 + 2 * MT_STEP, a_hi), bdesc, idesc, acc);
                        ptx::mma_bf16_ws<0, 2>(d + 3 * N, ptx::desc_pack(a0 + 3 * MT_STEP, a_hi), bdesc, idesc, acc);
                    }
                } else {
#pragma unroll
                    for (int mt = 0; mt < 4; ++mt) ptx::mma_bf16(d + mt * N, ptx::desc_pack(a0 + mt * MT_STEP, a_hi), bdesc, idesc, acc);
                }
            }
            ptx::mma_commit(empty[j]);
        }
    ptx::mma_commit(bar_t_full);
}

template <int N>
__device__ __forceinline__ void named_bar_sync(int id) { asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(N) : "memory"); }
template <int N>
__device__ __forceinline__ void named_bar_arrive(int id) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "n"(N) : "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <int EW>
__global__ void __launch_bounds__(128 + 32 * EW, 1) k_tower64(const __grid_constant__ FusedParams P) {
    constexpr int FZ_EPI_THREADS = 32 * EW, FC_BAR = FZ_EPI_THREADS + 64, PARTS = EW / 4;
    auto epi_bar = [] { named_bar_sync<FZ_EPI_THREADS>(1); };
    extern __shared__ __align__(1024) uint8_t smem[];
    const int warp = ptx::uniform_warp_id(), lane = threadIdx.x & 31;
    const uint32_t s0 = ptx::smem_u32(smem);
    auto b_full = [&](int s) { return s0 + 8u * s; };
    auto b_empty = [&](int s) { return s0 + 8u * (8 + s); };
    const uint32_t p_full = s0 + 8u * 16, t_full = s0 + 8u * 17, act_ready = s0 + 8u * 18, region_clean = s0 + 8u * 19;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 256);
    float* vbuf = reinterpret_cast<float*>(smem + 512);     // [7][64] value-conv outputs
    float* sbias = reinterpret_cast<float*>(smem + 2816);   // all folded biases (n_bias <= 1340 floats)
    float* sred = reinterpret_cast<float*>(smem + 2304);    // [EW][7] block-reduction scratch of the dense softmax (<= 448 B)
    uint8_t* region = smem + FZ_HDR;
    const uint32_t region_s = s0 + FZ_HDR;
    const uint32_t ring_s = region_s + FZ_REGION;

    const int my_items = P.items > (int)blockIdx.x ? (P.items - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    if (P.ts && threadIdx.x == 0 && blockIdx.x < 160) {
        long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        P.ts[128 + 2 * blockIdx.x] = t;
    }
    // PDL: this CTA may have become resident while the kernel before it on the stream (the pool's select) is still
    // draining.  Everything up to each role's pdl_wait() touches only this CTA's shared / tensor memory and constant
    // weights; planes, legal-action lists and the output rows are only touched after it.  Each role triggers the
    // dependents when its last item is nearly done.  (The pool's expand is launched without the PDL attribute by default:
    // letting it in early was measured 8-10 us per step slower whether triggered here or at the top; KB_PDL_MASK.)

    if (threadIdx.x == 0) {
        for (int s = 0; s < FZ_NSTAGE; ++s) {
            ptx::mbar_init(b_full(s), 1);
            ptx::mbar_init(b_empty(s), 1);
        }
        ptx::mbar_init(p_full, 1);
        ptx::mbar_init(t_full, 1);
        ptx::mbar_init(act_ready, EW);
        ptx::mbar_init(region_clean, EW);
        ptx::fence_barrier_init();
    }
    if (warp == 2) {
        ptx::tmem_alloc(ptx::smem_u32(tmem_slot), 512);
        ptx::tmem_relinquish();
    }
    for (int i = threadIdx.x; i < P.n_bias; i += blockDim.x) sbias[i] = P.bias[i];
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // wait for the producing kernel's data: its completion (PDL), or -- split select -- one of its two progress counters.
    // Every lane polls with acquire loads (the data is then visible to each of them); bounded, so a lost update cannot hang the GPU.
    auto wait_input = [&](int which) {
        if (!P.sync) {
            pdl_wait();
            return;
        }
        const unsigned* ctr = P.sync + which;
        for (int tries = 0; tries < (1 << 22); ++tries) {
            unsigned v;
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
            if ((int)(v - P.sync_target) >= 0) break;
            __nanosleep(40);
        }
        __syncwarp();
    };

    if (warp == 0) {
        // ===== producer (converged warp, one elected lane issues) =====
        int stage = 0, sphase = 0;
        for (int ii = 0; ii < my_items; ++ii) {
            const int item = (int)blockIdx.x + ii * (int)gridDim.x;
            ptx::mbar_wait(region_clean, ii & 1);
            if (ii == 0) {
                wait_input(0);
                if (P.sync) asm volatile("fence.proxy.async;" ::: "memory");  // planes were written by generic stores of a kernel that is still running
            }
            if (ptx::elect_one()) {
                ptx::mbar_arrive_expect_tx(p_full, SLAB_BYTES);
                ptx::bulk_g2s(region_s + SLAB_BYTES, P.planes + (size_t)item * IN_SLABS * SLAB_U4, SLAB_BYTES, p_full);
            }
            __syncwarp();
            const uint4* w = P.w;
            for (int l = 0; l < P.n_layers; ++l) {
                const FusedLayer& L = P.layer[l];
                const int nblocks = L.slabs_in * L.ntaps * (L.n / L.n_sub);
                const uint32_t bytes = (uint32_t)(L.n_sub * LINE_BYTES);
                for (int b = 0; b < nblocks; ++b) {
                    ptx::mbar_wait(b_empty(stage), sphase ^ 1);
                    if (ptx::elect_one()) {
                        ptx::mbar_arrive_expect_tx(b_full(stage), bytes);
                        ptx::bulk_g2s(ring_s + stage * FZ_STAGE, w, bytes, b_full(stage));
                        if (P.ts && blockIdx.x == 0 && ii == 0 && l >= P.n_layers - 3) P.ts[96 + (l - (P.n_layers - 3)) * 9 + b] = clock64();  // (profiling hook)
                    }
                    __syncwarp();
                    w += bytes / 16;
                    if (++stage == FZ_NSTAGE) {
                        stage = 0;
                        sphase ^= 1;
                    }
                }
            }
        }
        pdl_launch_dependents();
    } else if (warp == 1) {
        // ===== MMA issuer (converged warp, one elected lane issues) =====
        int stage = 0, sphase = 0;
        uint32_t act_phase = 0;
        const int wgroup = P.wait_group;
        if (!P.sync) pdl_wait();  // (this role touches nothing of the producing kernel: its inputs arrive through p_full)
        const bool gather = P.legal_act != nullptr && P.gather != 0;
        for (int ii = 0; ii < my_items; ++ii) {
            for (int l = 0; l < P.n_layers; ++l) {
                const FusedLayer& L = P.layer[l];
                if (gather && L.kind == 1) {  // the epilogue warps read policyconv2's rows straight from the ring: keep its position
                    const int adv = stage + L.slabs_in * L.ntaps * (L.n / L.n_sub);
                    sphase ^= (adv / FZ_NSTAGE) & 1;
                    stage = adv % FZ_NSTAGE;
                    continue;
                }
                // input of the layer: the planes (layer 0), else the previous layer written to smem and TMEM drained
                const uint32_t in_bar = l == 0 ? p_full : act_ready, in_par = l == 0 ? (uint32_t)(ii & 1) : act_phase;
                if (l != 0) act_phase ^= 1;
                const bool mstamp = P.ts != nullptr && blockIdx.x == 0 && ii == 0;
                const uint32_t a_src = region_s + L.src_off;
                const int nsub = L.n / L.n_sub, ksteps = L.ksteps, ntaps = L.ntaps, n = L.n;
                const uint32_t idesc = L.idesc;
                const bool ws = P.mma_ws != 0 && (L.n_sub == 64 || L.n_sub == 128);  // (the .ws form takes N = 64, 128, 256 only)
                const uint32_t a_hi = ptx::sw128_hi(TALL_PITCH * LINE_BYTES), b_hi = ptx::sw128_hi(1024);
                const int nblocks = nsub * L.slabs_in * ntaps;
                // One election per layer: the elected lane waits for weight blocks, issues and commits on its own.  The
                // tensor pipe's issue queue is short and anything but uniform-register arithmetic between two MMAs shows
                // as idle tensor cycles (tools/mma_ws_probe.cu, cycles per 128x64x16 MMA: 43.8 issued straight; 51.0 with
                // an election + __syncwarp per 16; 55.4 with a wait per 16 -- an mbarrier try_wait that succeeds at once,
                // even a plain shared-memory load): so the lane also looks at the ring only every `wgroup` blocks and
                // then waits for that many stages at once.  Weight blocks arrive in the order (sub-block, slab, tap).
                const bool straight = ntaps == 9 && nsub == 1 && L.slabs_in == 1 && L.n_sub == 64 && (ksteps == 4 || ksteps == 2) && P.wait_group != 1;
                const bool head_a = ntaps == 1 && nsub == 2 && L.slabs_in == 1 && L.n_sub == 64 && ksteps == 4 && P.wait_group != 1;  // policyconv
                const bool head_b = ntaps == 1 && nsub == 1 && L.slabs_in == 2 && L.n_sub == 80 && ksteps == 4 && P.wait_group != 1;  // policyconv2
                if (head_a || head_b) {
                    if (ptx::elect_one()) {
                        const long long m_t0 = mstamp ? clock64() : 0;
                        const uint32_t a_lo0 = ptx::sw128_lo(a_src), ring_lo = ptx::sw128_lo(ring_s);
                        if (head_b) tower_issue_1x1<1, 2, 80, false>(a_lo0, a_hi, b_hi, tmem_base, idesc, ring_lo, b_full(0), b_empty(0), stage, sphase, t_full, in_bar, in_par);
                        else if (ws) tower_issue_1x1<2, 1, 64, true>(a_lo0, a_hi, b_hi, tmem_base, idesc, ring_lo, b_full(0), b_empty(0), stage, sphase, t_full, in_bar, in_par);
                        else tower_issue_1x1<2, 1, 64, false>(a_lo0, a_hi, b_hi, tmem_base, idesc, ring_lo, b_full(0), b_empty(0), stage, sphase, t_full, in_bar, in_par);
                        if (mstamp && l < 16) P.ts[65 + 2 * l] = clock64() - m_t0;
                    }
                } else if (straight) {
                    if (ptx::elect_one()) {
                        const long long m_t0 = mstamp ? clock64() : 0;
                        const uint32_t a_lo0 = ptx::sw128_lo(a_src), ring_lo = ptx::sw128_lo(ring_s);
                        if (ksteps == 4) {
                            if (ws) tower_issue_3x3_n64<4, true>(a_lo0, a_hi, b_hi, tmem_base, idesc, ring_lo, b_full(0), b_empty(0), stage, sphase, t_full, in_bar, in_par);
                            else tower_issue_3x3_n64<4, false>(a_lo0, a_hi, b_hi, tmem_base, idesc, ring_lo, b_full(0), b_empty(0), stage, sphase, t_full, in_bar, in_par);
                        } else {
                            if (ws) tower_issue_3x3_n64<2, true>(a_lo0, a_hi, b_hi, tmem_base, idesc, ring_lo, b_full(0), b_empty(0), stage, sphase, t_full, in_bar, in_par);
                            else tower_issue_3x3_n64<2, false>(a_lo0, a_hi, b_hi, tmem_base, idesc, ring_lo, b_full(0), b_empty(0), stage, sphase, t_full, in_bar, in_par);
                        }
                        if (mstamp && l < 16) P.ts[65 + 2 * l] = clock64() - m_t0;
                    }
                } else if (ptx::elect_one()) {
                    ptx::mbar_wait(in_bar, in_par);
                    ptx::tc_fence_after();
                    const long long m_t0 = mstamp ? clock64() : 0;
                    long long m_wait = 0;
                    int st = stage, ph = sphase, blocks_left = nblocks, in_group = 0;
                    for (int sub = 0; sub < nsub; ++sub)
                        for (int ks = 0; ks < L.slabs_in; ++ks) {
                            const uint32_t a_lo0 = ptx::sw128_lo(a_src + ks * SLAB_BYTES);
                            const uint32_t d_base = tmem_base + sub * L.n_sub;
                            int dy = ntaps == 9 ? -1 : 0, dx = ntaps == 9 ? -1 : 0;
                            for (int tap = 0; tap < ntaps; ++tap) {
                                const uint32_t a_tap = a_lo0 + (uint32_t)((dy * TALL_PITCH + dx + 1) * (LINE_BYTES / 16));
                                if (in_group == 0) {
                                    const long long m_c0 = mstamp ? clock64() : 0;
                                    in_group = blocks_left < wgroup ? blocks_left : wgroup;
                                    for (int g = 0; g < in_group; ++g) {
                                        const int sg = st + g;
                                        ptx::mbar_wait(b_full(sg < FZ_NSTAGE ? sg : sg - FZ_NSTAGE), sg < FZ_NSTAGE ? ph : ph ^ 1);
                                    }
                                    if (mstamp) m_wait += clock64() - m_c0;
                                }
                                --in_group;
                                --blocks_left;
                                const uint32_t b_lo0 = ptx::sw128_lo(ring_s + st * FZ_STAGE);
                                uint32_t first = (ks | tap) == 0 ? 0u : 1u;
                                constexpr uint32_t MT_STEP = 16 * TALL_PITCH * LINE_BYTES / 16;
                                if (ws) {
                                    // the weight block of a K step stays in a collector buffer for its four M tiles (buffers
                                    // alternate so the next fill does not wait for the last use)
                                    for (int kk = 0; kk < ksteps; ++kk) {
                                        const uint64_t bdesc = ptx::desc_pack(b_lo0 + kk * 2, b_hi);
                                        const uint32_t a0 = a_tap + kk * 2;
                                        if (kk & 1) {
                                            ptx::mma_bf16_ws<1, 0>(d_base, ptx::desc_pack(a0, a_hi), bdesc, idesc, first);
                                            ptx::mma_bf16_ws<1, 1>(d_base + n, ptx::desc_pack(a0 + MT_STEP, a_hi), bdesc, idesc, first);
                                            ptx::mma_bf16_ws<1, 1>(d_base + 2 * n, ptx::desc_pack(a0 + 2 * MT_STEP, a_hi), bdesc, idesc, first);
                                            ptx::mma_bf16_ws<1, 2>(d_base + 3 * n, ptx::desc_pack(a0 + 3 * MT_STEP, a_hi), bdesc, idesc, first);
                                        } else {
                                            ptx::mma_bf16_ws<0, 0>(d_base, ptx::desc_pack(a0, a_hi), bdesc, idesc, first);
                                            ptx::mma_bf16_ws<0, 1>(d_base + n, ptx::desc_pack(a0 + MT_STEP, a_hi), bdesc, idesc, first);
                                            ptx::mma_bf16_ws<0, 1>(d_base + 2 * n, ptx::desc_pack(a0 + 2 * MT_STEP, a_hi), bdesc, idesc, first);
                                            ptx::mma_bf16_ws<0, 2>(d_base + 3 * n, ptx::desc_pack(a0 + 3 * MT_STEP, a_hi), bdesc, idesc, first);
                                        }
                                        first = 1u;
                                    }
                                } else {
                                    for (int kk = 0; kk < ksteps; ++kk) {
                                        const uint64_t bdesc = ptx::desc_pack(b_lo0 + kk * 2, b_hi);
#pragma unroll
                                        for (int mt = 0; mt < 4; ++mt) {
                                            const uint64_t adesc = ptx::desc_pack(a_tap + kk * 2 + mt * MT_STEP, a_hi);
                                            ptx::mma_bf16(d_base + mt * n, adesc, bdesc, idesc, first);
                                        }
                                        first = 1u;
                                    }
                                }
                                ptx::mma_commit(b_empty(st));
                                if (++st == FZ_NSTAGE) {
                                    st = 0;
                                    ph ^= 1;
                                }
                                if (++dx > 1) {
                                    dx = -1;
                                    ++dy;
                                }
                            }
                        }
                    ptx::mma_commit(t_full);
                    if (mstamp && l < 16) {  // (profiling hook: weight waits and issue span of the issuing lane per layer)
                        P.ts[64 + 2 * l] = m_wait;
                        P.ts[65 + 2 * l] = clock64() - m_t0;
                    }
                }
                __syncwarp();
                {   // every lane keeps the ring position
                    const int adv = stage + nblocks;
                    sphase ^= (adv / FZ_NSTAGE) & 1;
                    stage = adv % FZ_NSTAGE;
                }
            }
            if (ii + 1 == my_items) pdl_launch_dependents();
            // the last layer's epilogue also signals act_ready (TMEM drained) -- consume it
            ptx::mbar_wait(act_ready, act_phase);
            act_phase ^= 1;
        }
    } else if (warp == 2 || warp == 3) {
        // ===== value head, second half: Linear(64 -> 256) + tanh (nn.cpp:87-88) on the value conv's outputs, off the
        // epilogue warps' critical path (they go on with the policy head meanwhile).  Thread j owns outputs 4j..4j+3. =====
        const int j = threadIdx.x - 64;  // 0..63
        wait_input(0);  // the value rows of the previous step are read by the producing kernel's expand phase, which precedes its planes
        for (int ii = 0; ii < my_items; ++ii) {
            const int item = (int)blockIdx.x + ii * (int)gridDim.x;
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(P.fcb) + j);
            float o[NB][4];
#pragma unroll
            for (int s = 0; s < NB; ++s) { o[s][0] = b4.x; o[s][1] = b4.y; o[s][2] = b4.z; o[s][3] = b4.w; }
            const float4* wt = reinterpret_cast<const float4*>(P.fct) + j;  // fct[p][256]
            constexpr int FCB = EW == 8 ? 16 : 8;  // weight rows in flight per round (register budget: 170 / 102 per thread)
            float4 wq[FCB];
#pragma unroll
            for (int p = 0; p < FCB; ++p) wq[p] = __ldg(wt + p * 64);  // first rows in flight before the wait
            named_bar_sync<FC_BAR>(2);
#pragma unroll
            for (int pq = 0; pq < 64 / FCB; ++pq) {
                float4 wn[FCB];
                if (pq + 1 < 64 / FCB) {
#pragma unroll
                    for (int p = 0; p < FCB; ++p) wn[p] = __ldg(wt + ((pq + 1) * FCB + p) * 64);
                }
#pragma unroll
                for (int p = 0; p < FCB; ++p) {
#pragma unroll
                    for (int s = 0; s < NB; ++s) {
                        const float v = vbuf[s * 64 + pq * FCB + p];
                        o[s][0] = fmaf(v, wq[p].x, o[s][0]);
                        o[s][1] = fmaf(v, wq[p].y, o[s][1]);
                        o[s][2] = fmaf(v, wq[p].z, o[s][2]);
                        o[s][3] = fmaf(v, wq[p].w, o[s][3]);
                    }
                }
                if (pq + 1 < 64 / FCB) {
#pragma unroll
                    for (int p = 0; p < FCB; ++p) wq[p] = wn[p];
                }
            }
            if (ii + 1 < my_items) named_bar_arrive<FC_BAR>(3);  // vbuf may be overwritten
            else pdl_launch_dependents();
            bool bad = false;
#pragma unroll
            for (int s = 0; s < NB; ++s) {
                const int board = item * NB + s;
                if (board < P.boards) {
                    const float4 t = make_float4(tanhf(o[s][0]), tanhf(o[s][1]), tanhf(o[s][2]), tanhf(o[s][3]));
                    bad |= (t.x != t.x) | (t.y != t.y) | (t.z != t.z) | (t.w != t.w);
                    reinterpret_cast<float4*>(P.value256 + (size_t)board * 256)[j] = t;
                }
            }
            if (bad) atomicExch(P.nan_flag, 1);
        }
    } else if (warp >= 4) {
        // ===== epilogue warps =====
        const int e = warp - 4, q = e & 3, half = e >> 2;  // half: which of the PARTS tile sets of its lane quarter this warp owns
        const int et = threadIdx.x - 128;  // 0..255
        uint32_t t_phase = 0;
        const bool stamp = P.ts && blockIdx.x == 0 && et == 0;
        int nts = 0;
#define KB_STAMP() do { if (stamp && nts < 60) P.ts[nts++] = clock64(); } while (0)
        KB_STAMP();
        for (int ii = 0; ii < my_items; ++ii) {
            const int item = (int)blockIdx.x + ii * (int)gridDim.x;
            {   // The 3x3 layers read the pad pixels (the row above each board, the column on each side) as zero.  Y gets its
                // pads with the input slab (the TMA copy below brings the whole slab, pads included, and no epilogue ever
                // writes a pad); X's pads were overwritten by the previous item's logits, or never initialised: clear those
                // 192 pixel lines (24 KB) -- not the whole 160 KB region, which cost 6 k cycles at the head of every item.
                uint4* x4 = reinterpret_cast<uint4*>(region);  // X slab = region offset 0
                const uint4 z = make_uint4(0, 0, 0, 0);
                for (int i = et; i < 192 * 8; i += FZ_EPI_THREADS) {
                    const int pp = i >> 3, j = i & 7;
                    int row, col;
                    if (pp < 80) {  // the 8 pad rows 0, 9, ..., 63
                        row = (pp / TALL_PITCH) * 9;
                        col = pp - (pp / TALL_PITCH) * TALL_PITCH;
                    } else {        // columns 0 and 9 of the 56 board rows
                        const int qq = pp - 80, br = qq >> 1;
                        row = 1 + (br >> 3) * 9 + (br & 7);
                        col = (qq & 1) ? TALL_PITCH - 1 : 0;
                    }
                    x4[(row * TALL_PITCH + col) * 8 + j] = z;
                }
                fence_async_smem();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(region_clean);
            }
            if (ii == 0 && !P.sync) pdl_wait();
            KB_STAMP();
            const bool gather = P.legal_act != nullptr && P.gather != 0;
            int g_n = 0;                       // gather head: move count and four actions per lane of board `e`'s list
            uint2 g_act4 = make_uint2(0, 0);
            for (int l = 0; l < P.n_layers - (gather ? 1 : 0); ++l) {
                const FusedLayer& L = P.layer[l];
                if (gather && l == P.tower_layers && !P.sync && e < NB) {
                    // the gather head below needs the boards' move lists in shared memory: warp b requests board b's row
                    // now (count + 128 actions, one coalesced read) and parks it after this layer's epilogue, which
                    // hides the L2 latency.  (Rows hold 128 entries: reading past the count is harmless.)
                    const int pboard = item * NB + e;
                    if (pboard < P.boards) {
                        g_n = *reinterpret_cast<const int*>(P.legal_n + (size_t)pboard * P.legal_stride);
                        g_act4 = reinterpret_cast<const uint2*>(P.legal_act + (size_t)pboard * P.legal_stride)[lane];
                    }
                }
                ptx::mbar_wait(t_full, t_phase);
                t_phase ^= 1;
                ptx::tc_fence_after();
                KB_STAMP();
                const int ncg = L.n / 16;
                // The two warps of a lane quarter split the TILES (even / odd), not the columns: the same instruction count
                // per warp, but half as many dependent TMEM-load -> convert -> store chains for the 64-column tower layers.
#pragma unroll 1
                for (int mt = half; mt < 4; mt += PARTS) {
                    const int r = 32 * q + lane;
                    const int R = 16 * mt + (r >> 3), x = r & 7;
                    const int slot = (R - 1) / 9, y = (R - 1) - slot * 9;
                    const bool valid = R >= 1 && y < 8 && slot < NB;
                    const int px = R * TALL_PITCH + 1 + x;
                    const uint32_t taddr = tmem_base + ((uint32_t)(32 * q) << 16) + mt * L.n;
#pragma unroll 1
                  constexpr int RG = EW == 8 ? 4 : 2;  // column groups (16 fp32 columns each) per round: register budget
                  for (int cg0 = 0; cg0 < ncg; cg0 += RG) {
                    const int cg1 = cg0 + RG < ncg ? cg0 + RG : ncg;
                    // up to RG column groups per round: issue every TMEM load, wait once
                    uint32_t v[RG][16];
#pragma unroll
                    for (int g = 0; g < RG; ++g)
                        if (cg0 + g < cg1) ptx::tmem_ld16(taddr + (cg0 + g) * 16, v[g]);
                    ptx::tmem_ld_wait();
                    if (!valid) continue;
#pragma unroll
                    for (int g = 0; g < RG; ++g) {
                        const int cg = cg0 + g;
                        if (cg >= cg1) break;
                        float f[16];
                        const float4* b4 = reinterpret_cast<const float4*>(sbias + L.bias_off + cg * 16);
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float4 bb = b4[j];
                            f[4 * j + 0] = __uint_as_float(v[g][4 * j + 0]) + bb.x;
                            f[4 * j + 1] = __uint_as_float(v[g][4 * j + 1]) + bb.y;
                            f[4 * j + 2] = __uint_as_float(v[g][4 * j + 2]) + bb.z;
                            f[4 * j + 3] = __uint_as_float(v[g][4 * j + 3]) + bb.w;
                        }
                        if (L.relu) {
#pragma unroll
                            for (int j = 0; j < 16; ++j) f[j] = fmaxf(f[j], 0.0f);
                        }
                        if (L.kind == 1) {
                            float* lg = reinterpret_cast<float*>(region) + (size_t)slot * KB_PSIZE + (y * 8 + x) * 73;
#pragma unroll
                            for (int j = 0; j < 16; ++j)
                                if (cg * 16 + j < 73) lg[cg * 16 + j] = f[j];
                        } else {
#pragma unroll
                            for (int h = 0; h < 2; ++h) {
                                const int ch = cg * 16 + 8 * h;
                                uint4* dst = reinterpret_cast<uint4*>(region + L.dst_off) + (size_t)(ch >> 6) * SLAB_U4 + chunk_u4(px, (ch >> 3) & 7);
                                if (L.skip) {  // x = skip + relu(...), the skip is the destination itself (nn.cpp:31)
                                    const uint4 s4 = *dst;
                                    const __nv_bfloat162* sb = reinterpret_cast<const __nv_bfloat162*>(&s4);
#pragma unroll
                                    for (int k = 0; k < 4; ++k) {
                                        const float2 sv = __bfloat1622float2(sb[k]);
                                        f[h * 8 + 2 * k] += sv.x;
                                        f[h * 8 + 2 * k + 1] += sv.y;
                                    }
                                }
                                uint32_t w4[4];
#pragma unroll
                                for (int k = 0; k < 4; ++k) {
                                    const __nv_bfloat162 b = __floats2bfloat162_rn(f[h * 8 + 2 * k], f[h * 8 + 2 * k + 1]);
                                    w4[k] = *reinterpret_cast<const uint32_t*>(&b);
                                }
                                *dst = make_uint4(w4[0], w4[1], w4[2], w4[3]);
                            }
                        }
                    }
                  }
                }
                fence_async_smem();       // generic-proxy smem writes -> visible to tcgen05.mma
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(act_ready);
                KB_STAMP();
                if (l == P.tower_layers - 1) {
                    // ---- value conv 1x1 + ReLU on X (nn.cpp:83-86); the 64 -> 256 Linear + tanh runs on warps 2-3 ----
                    // Every thread takes the pixels it has just written itself (and is the only one to overwrite with the
                    // policy head's output later): no barrier on either side.  The folded weights sit behind the biases
                    // in shared memory.  fp32 fma chain in channel order from the folded bias.
                    if (ii > 0) named_bar_sync<FC_BAR>(3);  // vbuf consumed by the previous item's Linear
                    const uint4* X = reinterpret_cast<const uint4*>(region + P.layer[l].dst_off);
                    const float4* wv4 = reinterpret_cast<const float4*>(sbias + P.wv_off);
#pragma unroll 1
                    for (int mt = half; mt < 4; mt += PARTS) {
                        const int r = 32 * q + lane;
                        const int R = 16 * mt + (r >> 3), x = r & 7;
                        const int slot = (R - 1) / 9, y = (R - 1) - slot * 9;
                        if (!(R >= 1 && y < 8 && slot < NB)) continue;
                        const int px = R * TALL_PITCH + 1 + x;
                        float acc = P.bv;
#pragma unroll
                        for (int c = 0; c < 8; ++c) {
                            const uint4 a4 = X[chunk_u4(px, c)];
                            const float4 wa = wv4[2 * c], wb = wv4[2 * c + 1];
                            const __nv_bfloat162* ab = reinterpret_cast<const __nv_bfloat162*>(&a4);
                            const float wvv[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const float2 av = __bfloat1622float2(ab[k]);
                                acc = fmaf(av.x, wvv[2 * k], acc);
                                acc = fmaf(av.y, wvv[2 * k + 1], acc);
                            }
                        }
                        vbuf[slot * 64 + y * 8 + x] = fmaxf(acc, 0.0f);
                    }
                    named_bar_arrive<FC_BAR>(2);  // vbuf ready for warps 2-3
                }
            }
            auto pad_row = [&](int b) { return region + (size_t)(9 * b) * TALL_PITCH * LINE_BYTES; };  // 1280 bytes: [128] logits, count, [128] actions
            if (gather) {
                if (e < NB) {  // park board e's move list (see the gather head below)
                    const int board = item * NB + e;
                    if (P.sync && board < P.boards) {  // split select: the lists were not there earlier
                        wait_input(1);
                        g_n = *reinterpret_cast<const int*>(P.legal_n + (size_t)board * P.legal_stride);
                        g_act4 = reinterpret_cast<const uint2*>(P.legal_act + (size_t)board * P.legal_stride)[lane];
                    }
                    if (board >= P.boards) g_n = 0;
                    if (lane == 0) *reinterpret_cast<int*>(pad_row(e) + 512) = g_n;
                    reinterpret_cast<uint2*>(pad_row(e) + 528)[lane] = g_act4;
                }
                const int s_abs = ii * P.blocks_per_item + P.blocks_before_logits;  // policyconv2's two blocks in the ring
                ptx::mbar_wait(b_full(s_abs % FZ_NSTAGE), (uint32_t)((s_abs / FZ_NSTAGE) & 1));
                ptx::mbar_wait(b_full((s_abs + 1) % FZ_NSTAGE), (uint32_t)(((s_abs + 1) / FZ_NSTAGE) & 1));
            }
            // ---- softmax over the 4672 logits of each board (nn.cpp:80) ----
            epi_bar();  // (gather head: H complete, move lists parked)
            if (ii + 1 == my_items) pdl_launch_dependents();
            KB_STAMP();
            if (gather) {
                // ---- policyconv2 + softmax numerators for the legal moves only (nn.cpp:76-80) ----
                // A board has ~30 legal moves, so 30 of its 4 672 logits are ever read: instead of the 1x1 conv over all
                // pixels (32 MMAs, then 73 fp32 stores per pixel, then a gather) every legal move takes its pixel's 128
                // features from H and policyconv2's row from the weight ring and does the 128-term dot product itself.
                // Work unit = eight moves of one board, four lanes per move (32 features each, two fma chains, xor-shuffle
                // sum); the units of the item's seven boards are dealt round-robin to ALL epilogue warps -- one warp runs
                // this at ~4.5 cycles per instruction (dependent integer / convert / fma chains), so what matters is how
                // many warps share it.  Move lists and raw logits wait in pad lines of the X slab (the pad row above each
                // board: never written by the policy head, zero-filled again before the next item).
                const FusedLayer& L = P.layer[P.n_layers - 1];
                const int s_abs = ii * P.blocks_per_item + P.blocks_before_logits;  // (policyconv2 = two K slabs = two blocks)
                const int st0 = s_abs % FZ_NSTAGE, st1 = (s_abs + 1) % FZ_NSTAGE;
                {
                    int ub[NB + 1];  // unit prefix over the boards
                    ub[0] = 0;
#pragma unroll
                    for (int b = 0; b < NB; ++b) ub[b + 1] = ub[b] + ((*reinterpret_cast<const int*>(pad_row(b) + 512) + 7) >> 3);
                    const int mslot = lane >> 2, kq = lane & 3;
                    const uint4* hslab = reinterpret_cast<const uint4*>(region + (kq >> 1) * SLAB_BYTES);
                    const uint4* wblk = reinterpret_cast<const uint4*>(smem + FZ_HDR + FZ_REGION + ((kq >> 1) ? st1 : st0) * FZ_STAGE);
                    const int c0 = (kq & 1) * 4;  // first 16-byte chunk of this lane's 32 features inside the 128-byte line
#pragma unroll 1
                    for (int u = e; u < ub[NB]; u += EW) {
                        int b = 0;
#pragma unroll
                        for (int k = 1; k < NB; ++k) b += u >= ub[k] ? 1 : 0;
                        int ubase = 0;
#pragma unroll
                        for (int k = 1; k < NB; ++k) ubase = b == k ? ub[k] : ubase;
                        const uint8_t* row = pad_row(b);
                        const int n = *reinterpret_cast<const int*>(row + 512);
                        const int i = (u - ubase) * 8 + mslot;
                        float a0 = 0.0f, a1 = 0.0f;
                        int c = 0;
                        if (i < n) {
                            const int act = reinterpret_cast<const uint16_t*>(row + 528)[i];
                            const int sq = act / 73;
                            c = act - sq * 73;
                            const int px = tall_pixel(b, sq);
                            const uint4* hrow = hslab + px * 8;
                            const uint4* wrow = wblk + c * 8;
                            uint4 h4[4], w4[4];
#pragma unroll
                            for (int ch = 0; ch < 4; ++ch) {
                                h4[ch] = hrow[(c0 + ch) ^ (px & 7)];
                                w4[ch] = wrow[(c0 + ch) ^ (c & 7)];
                            }
#pragma unroll
                            for (int ch = 0; ch < 4; ++ch) {
                                const __nv_bfloat162* hb = reinterpret_cast<const __nv_bfloat162*>(&h4[ch]);
                                const __nv_bfloat162* wb = reinterpret_cast<const __nv_bfloat162*>(&w4[ch]);
#pragma unroll
                                for (int t = 0; t < 4; ++t) {
                                    const float2 hv = __bfloat1622float2(hb[t]), wv2 = __bfloat1622float2(wb[t]);
                                    a0 = fmaf(hv.x, wv2.x, a0);
                                    a1 = fmaf(hv.y, wv2.y, a1);
                                }
                            }
                        }
                        float lg = a0 + a1;
                        lg += __shfl_xor_sync(0xffffffffu, lg, 1);
                        lg += __shfl_xor_sync(0xffffffffu, lg, 2);
                        if (i < n && kq == 0) reinterpret_cast<float*>(const_cast<uint8_t*>(row))[i] = lg + sbias[L.bias_off + c];
                    }
                }
                KB_STAMP();
                epi_bar();  // every logit written; every read of the weight ring done
                KB_STAMP();
                if (et == 0) {
                    ptx::mbar_arrive(b_empty(st0));
                    ptx::mbar_arrive(b_empty(st1));
                }
                if (e < NB) {  // warp b: maximum over board b's logits, numerators out
                    const int board = item * NB + e;
                    const uint8_t* row = pad_row(e);
                    const int n = *reinterpret_cast<const int*>(row + 512);
                    const float* lgs = reinterpret_cast<const float*>(row);
                    float l[4];
                    float m = -INFINITY;
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        const int i = lane + 32 * r;
                        l[r] = i < n ? lgs[i] : -INFINITY;
                        m = fmaxf(m, l[r]);
                    }
                    for (int off = 16; off; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
                    bool bad = false;
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        const int i = lane + 32 * r;
                        if (i < n) {
                            const float o = __expf(l[r] - m);
                            bad |= (o != o);
                            P.prior[(size_t)board * 128 + i] = o;
                        }
                    }
                    if (bad) atomicExch(P.nan_flag, 1);
                }
            } else if (P.legal_act) {
                if (P.sync) wait_input(1);  // the move lists are the last thing the producing kernel writes
                // ---- softmax numerators over the legal moves only: warp e gathers board e's logits ----
                const float* lg = reinterpret_cast<const float*>(region);
                const int board = item * NB + e;
                if (e < NB && board < P.boards) {
                    const int n = *reinterpret_cast<const int*>(P.legal_n + (size_t)board * P.legal_stride);
                    const uint16_t* acts = reinterpret_cast<const uint16_t*>(P.legal_act + (size_t)board * P.legal_stride);
                    float l[4];
                    float m = -INFINITY;
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        const int i = lane + 32 * r;
                        l[r] = i < n ? lg[e * KB_PSIZE + acts[i]] : -INFINITY;
                        m = fmaxf(m, l[r]);
                    }
                    for (int off = 16; off; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
                    bool bad = false;
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        const int i = lane + 32 * r;
                        if (i < n) {
                            const float o = __expf(l[r] - m);
                            bad |= (o != o);
                            P.prior[(size_t)board * 128 + i] = o;
                        }
                    }
                    if (bad) atomicExch(P.nan_flag, 1);
                }
            } else {
                // every thread owns 19 logits of each of the 7 boards; two block-wide reductions in total
                float* red = sred;  // [EW warps][7] maxima, then [EW][7] sums
                const float* lg = reinterpret_cast<const float*>(region);
                constexpr int PER = (KB_PSIZE + FZ_EPI_THREADS - 1) / FZ_EPI_THREADS;  // 19
                float m[NB], sum[NB];
#pragma unroll
                for (int b = 0; b < NB; ++b) {
                    float mm = -INFINITY;
#pragma unroll
                    for (int i = 0; i < PER; ++i) {
                        const int idx = et + i * FZ_EPI_THREADS;
                        if (idx < KB_PSIZE) mm = fmaxf(mm, lg[b * KB_PSIZE + idx]);
                    }
                    for (int off = 16; off; off >>= 1) mm = fmaxf(mm, __shfl_xor_sync(0xffffffffu, mm, off));
                    if (lane == 0) red[e * NB + b] = mm;
                }
                epi_bar();
#pragma unroll
                for (int b = 0; b < NB; ++b) {
                    float mm = red[b];
#pragma unroll
                    for (int w = 1; w < EW; ++w) mm = fmaxf(mm, red[w * NB + b]);
                    m[b] = mm;
                }
                epi_bar();  // maxima consumed before the sums reuse the scratch
#pragma unroll
                for (int b = 0; b < NB; ++b) {
                    float ss = 0.0f;
#pragma unroll
                    for (int i = 0; i < PER; ++i) {
                        const int idx = et + i * FZ_EPI_THREADS;
                        if (idx < KB_PSIZE) ss += __expf(lg[b * KB_PSIZE + idx] - m[b]);
                    }
                    for (int off = 16; off; off >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, off);
                    if (lane == 0) red[e * NB + b] = ss;
                }
                epi_bar();
                bool bad = false;
#pragma unroll
                for (int b = 0; b < NB; ++b) {
                    float ss = red[b];
#pragma unroll
                    for (int w = 1; w < EW; ++w) ss += red[w * NB + b];
                    sum[b] = ss;
                    const int board = item * NB + b;
                    if (board < P.boards) {
                        const float inv = 1.0f / ss;
                        float* out = P.policy + (size_t)board * KB_PSIZE;
#pragma unroll
                        for (int i = 0; i < PER; ++i) {
                            const int idx = et + i * FZ_EPI_THREADS;
                            if (idx < KB_PSIZE) {
                                const float o = __expf(lg[b * KB_PSIZE + idx] - m[b]) * inv;
                                bad |= (o != o);
                                out[idx] = o;
                            }
                        }
                    }
                }
                (void)sum;
                if (bad) atomicExch(P.nan_flag, 1);
            }
            epi_bar();
            KB_STAMP();
        }
#undef KB_STAMP
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 2) ptx::tmem_dealloc(tmem_base, 512);
    if (P.ts && threadIdx.x == 0 && blockIdx.x < 160) {
        long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        P.ts[129 + 2 * blockIdx.x] = t;
    }
}

#include "tower_pipe.inl"

// valueconv 1x1 (F -> 1) + BN + ReLU, Linear(64 -> 256), tanh (nn.cpp:83-88).  One block per
// board.  wv / bv: BN-folded conv weights; fct: valuefc.weight transposed to [64][256].
__global__ void __launch_bounds__(256) k_value_head(const uint4* x, int slabs, int boards, const float* wv, float bv, const float* fct,
                                                     const float* fcb, float* value256, int* nan_flag) {
    __shared__ float part[4][64];
    __shared__ float v[64];
    const int b = blockIdx.x;
    if (b >= boards) return;
    const int item = b / NB, slot = b - item * NB;
    const int t = threadIdx.x, pix = t & 63, quarter = t >> 6;
    const int px = tall_pixel(slot, pix);
    float acc = 0.0f;
    for (int c = quarter; c < slabs * 8; c += 4) {  // c = 8-channel chunk index over all slabs
        const uint4 a4 = x[((size_t)item * slabs + (c >> 3)) * SLAB_U4 + chunk_u4(px, c & 7)];
        const __nv_bfloat162* ab = reinterpret_cast<const __nv_bfloat162*>(&a4);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float2 av = __bfloat1622float2(ab[k]);
            acc = fmaf(av.x, wv[c * 8 + 2 * k], acc);
            acc = fmaf(av.y, wv[c * 8 + 2 * k + 1], acc);
        }
    }
    part[quarter][pix] = acc;
    __syncthreads();
    if (t < 64) v[t] = fmaxf(part[0][t] + part[1][t] + part[2][t] + part[3][t] + bv, 0.0f);
    __syncthreads();
    float o = fcb[t];
#pragma unroll 8
    for (int p = 0; p < 64; ++p) o = fmaf(v[p], fct[p * 256 + t], o);
    o = tanhf(o);
    if (o != o) atomicExch(nan_flag, 1);
    value256[(size_t)b * 256 + t] = o;
}

// exp(log_softmax(logits)) over the 4672 actions of each board, in place (nn.cpp:80)
__global__ void __launch_bounds__(256) k_softmax(float* logits, int boards, int* nan_flag) {
    __shared__ float red[8];
    __shared__ float bcast;
    const int b = blockIdx.x;
    if (b >= boards) return;
    float* row = logits + (size_t)b * KB_PSIZE;
    const int t = threadIdx.x;
    constexpr int PER = (KB_PSIZE + 255) / 256;
    float v[PER];
    float m = -INFINITY;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
        const int idx = t + 256 * i;
        v[i] = idx < KB_PSIZE ? row[idx] : -INFINITY;
        m = fmaxf(m, v[i]);
    }
    for (int off = 16; off; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
    if ((t & 31) == 0) red[t >> 5] = m;
    __syncthreads();
    if (t == 0) {
        float mm = red[0];
        for (int i = 1; i < 8; ++i) mm = fmaxf(mm, red[i]);
        bcast = mm;
    }
    __syncthreads();
    m = bcast;
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
        v[i] = (t + 256 * i) < KB_PSIZE ? expf(v[i] - m) : 0.0f;
        s += v[i];
    }
    for (int off = 16; off; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    __syncthreads();
    if ((t & 31) == 0) red[t >> 5] = s;
    __syncthreads();
    if (t == 0) {
        float ss = 0.0f;
        for (int i = 0; i < 8; ++i) ss += red[i];
        bcast = ss;
    }
    __syncthreads();
    const float inv = 1.0f / bcast;
    bool bad = false;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
        const int idx = t + 256 * i;
        if (idx < KB_PSIZE) {
            const float o = v[i] * inv;
            bad |= (o != o);
            row[idx] = o;
        }
    }
    if (bad) atomicExch(nan_flag, 1);
}

// Legal-gather for the per-layer path: one warp per board turns fp32 logits [board][4672] into
// softmax numerators over that board's legal actions (see net_forward_legal_async).
__global__ void __launch_bounds__(256) k_legal_prior(const float* logits, int boards, const uint8_t* act_base, const uint8_t* nact_base,
                                                      unsigned long long stride, float* prior, int* nan_flag) {
    const int board = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (board >= boards) return;
    const int n = *reinterpret_cast<const int*>(nact_base + (size_t)board * stride);
    const uint16_t* acts = reinterpret_cast<const uint16_t*>(act_base + (size_t)board * stride);
    const float* lg = logits + (size_t)board * KB_PSIZE;
    float l[4];
    float m = -INFINITY;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int i = lane + 32 * r;
        l[r] = i < n ? lg[acts[i]] : -INFINITY;
        m = fmaxf(m, l[r]);
    }
    for (int off = 16; off; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
    bool bad = false;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int i = lane + 32 * r;
        if (i < n) {
            const float o = __expf(l[r] - m);
            bad |= (o != o);
            prior[(size_t)board * 128 + i] = o;
        }
    }
    if (bad) atomicExch(nan_flag, 1);
}

}  // namespace kb

// ==========================================================================================
// host side
// ==========================================================================================
using namespace kb;

namespace {

struct Layer {
    int slabs_in, ksteps, n_tile, n_total, n_valid, ntaps, relu;
    uint4* w = nullptr;
    float* bias = nullptr;
    std::vector<uint16_t> hw;   // host copy of the packed weights (fused-kernel concatenation)
    std::vector<float> hbias;
};

uint16_t f2bf(float f) {  // round-to-nearest-even, like __float2bfloat16_rn
    uint32_t u;
    memcpy(&u, &f, 4);
    if ((u & 0x7FFFFFFFu) > 0x7F800000u) return (uint16_t)((u >> 16) | 0x40);
    u += 0x7FFFu + ((u >> 16) & 1u);
    return (uint16_t)(u >> 16);
}

}  // namespace

// Shared / exclusive lock that prefers writers: inference threads hold the weights shared almost all the time (a
// kb_pool_step call is one long shared section), so a reader-preferring lock (glibc's default rwlock behind
// std::shared_mutex) could starve NN::read / NN::train's weight swap indefinitely.  A waiting writer stops new readers.
class WeightLock {
    std::mutex m;
    std::condition_variable cv;
    int readers = 0, writers_waiting = 0;
    bool writing = false;

   public:
    void lock_shared() {
        std::unique_lock<std::mutex> g(m);
        cv.wait(g, [&] { return !writing && writers_waiting == 0; });
        ++readers;
    }
    void unlock_shared() {
        std::unique_lock<std::mutex> g(m);
        if (--readers == 0) cv.notify_all();
    }
    void lock() {
        std::unique_lock<std::mutex> g(m);
        ++writers_waiting;
        cv.wait(g, [&] { return !writing && readers == 0; });
        --writers_waiting;
        writing = true;
    }
    void unlock() {
        std::unique_lock<std::mutex> g(m);
        writing = false;
        cv.notify_all();
    }
};

// One host-pointer inference call's private state: workspace + staging (device and pinned host)
struct InferCtx {
    kb::NetWs ws;
    float *obs_dev = nullptr, *pol_dev = nullptr, *val_dev = nullptr;
    float *obs_pin = nullptr, *pol_pin = nullptr, *val_pin = nullptr;
    int stage_cap = 0;
    cudaEvent_t ev[2] = {nullptr, nullptr};
};

struct kb_net {
    int filters, residuals;
    int device = 0;
    bool loaded = false;
    std::vector<Layer> layers;  // conv1, (res conv1, res conv2)*, policyconv, policyconv2
    float *wv = nullptr, *fct = nullptr, *fcb = nullptr;
    float bv = 0.0f;
    int launches = 0;
    // fused single-kernel path (filters == 64)
    bool fused = false;
    int tower_pipe = 0;  // fused path: 1 = k_tower64p (per-tile hand-over between layers, tower_pipe.inl), 0 = k_tower64
    kb::FusedParams fp;
    uint4* fused_w = nullptr;
    float* fused_bias = nullptr;
    long long* ts_dev = nullptr;
    // weights: shared by every forward, exclusive for load_blob (nn.cpp:164-168, 206)
    WeightLock mu;
    // host-pointer calls (kb_net_infer / kb_net_forward_full): one private context per call in flight
    std::mutex ctx_mu;
    std::vector<InferCtx*> ctx_free;
    std::vector<InferCtx*> ctx_kept;  // contexts host threads keep for kb_net_forward_dev (freed with the net)
    unsigned long long uid = 0;       // distinguishes this net from an earlier one at the same address (thread-local caches)
    InferCtx* dbg_ctx = nullptr;  // context of the last kb_net_forward_full (kb_net_debug_activation reads it)
};

namespace kb {

void net_lock_shared(kb_net* net) { net->mu.lock_shared(); }
void net_unlock_shared(kb_net* net) { net->mu.unlock_shared(); }
int net_device(kb_net* net) { return net->device; }

void ws_free(NetWs& ws) {
    cudaFree(ws.P); cudaFree(ws.X); cudaFree(ws.Y); cudaFree(ws.H); cudaFree(ws.logits); cudaFree(ws.nan_flag);
    ws = NetWs{};
}

int ws_reserve(NetWs& ws, kb_net* net, int batch, cudaStream_t st) {
    const int fs = net->fused ? 0 : (net->filters + 63) / 64;  // the fused tower keeps its activations in shared memory
    if (batch <= ws.cap_boards && fs <= ws.slabs) return KB_OK;
    if (batch < ws.cap_boards) batch = ws.cap_boards;
    cudaStreamSynchronize(st);  // the owner's stream is the only one that can still be using the old buffers
    cudaFree(ws.P); cudaFree(ws.X); cudaFree(ws.Y); cudaFree(ws.H);
    ws.P = ws.X = ws.Y = ws.H = nullptr;
    ws.cap_boards = 0;
    KB_CUDA(cudaMalloc(&ws.P, act_bytes(batch, IN_SLABS)));
    // pad pixels (and unused channel chunks) must read as zero forever; epilogues and encoders
    // only ever write the chunks of board pixels they own
    KB_CUDA(cudaMemsetAsync(ws.P, 0, act_bytes(batch, IN_SLABS), st));
    if (fs) {
        KB_CUDA(cudaMalloc(&ws.X, act_bytes(batch, fs)));
        KB_CUDA(cudaMalloc(&ws.Y, act_bytes(batch, fs)));
        KB_CUDA(cudaMalloc(&ws.H, act_bytes(batch, 2)));
        KB_CUDA(cudaMemsetAsync(ws.X, 0, act_bytes(batch, fs), st));
        KB_CUDA(cudaMemsetAsync(ws.Y, 0, act_bytes(batch, fs), st));
        KB_CUDA(cudaMemsetAsync(ws.H, 0, act_bytes(batch, 2), st));
    }
    if (!ws.nan_flag) {
        KB_CUDA(cudaMalloc(&ws.nan_flag, sizeof(int)));
        KB_CUDA(cudaMemsetAsync(ws.nan_flag, 0, sizeof(int), st));
    }
    ws.cap_boards = batch;
    ws.slabs = fs;
    return KB_OK;
}
int net_launches_per_forward(kb_net* net) { return net->fused ? 1 : (int)net->layers.size() + 2; }

static long long* g_conv_ts = nullptr;  // profiling hook (kb_net_debug_timestamps)
static int g_conv_ts_slot = 0;          // 16 counters per launch, 8 launches kept

template <int N_TILE>
static int launch_conv(const ConvParams& p, cudaStream_t st) {
    using C = ConvCfg<N_TILE>;
    static bool configured[16] = {};  // function attributes are per device
    if (!configured[current_device() & 15]) {
        KB_CUDA(cudaFuncSetAttribute(k_conv<N_TILE>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
        configured[current_device() & 15] = true;
    }
    const int grid = p.items < sm_count() ? p.items : sm_count();
    k_conv<N_TILE><<<grid, 256, C::SMEM, st>>>(p);
    KB_CUDA(cudaGetLastError());
    return KB_OK;
}

template <int N_TILE>
static int launch_conv2(const ConvParams& p, cudaStream_t st) {
    using C = Conv2Cfg<N_TILE>;
    static bool configured[16] = {};
    if (!configured[current_device() & 15]) {
        KB_CUDA(cudaFuncSetAttribute(k_conv2<N_TILE>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
        configured[current_device() & 15] = true;
    }
    const int pairs = (p.items + 1) / 2, max_pairs = sm_count() / 2;
    const int grid = 2 * (pairs < max_pairs ? pairs : max_pairs);
    k_conv2<N_TILE><<<grid, 384, C::SMEM, st>>>(p);  // cluster dims (2,1,1) are a kernel attribute
    KB_CUDA(cudaGetLastError());
    return KB_OK;
}

static bool use_pair_kernel() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("KB_NO_PAIR_CONV");
        v = (e && e[0] == '1') ? 0 : 1;
    }
    return v == 1;
}

static int run_conv(const Layer& L, const uint4* in, uint4* out, const uint4* skip, float* out_f32, int boards, cudaStream_t st) {
    ConvParams p;
    p.in = in;
    p.out = out;
    p.skip = skip;
    p.out_f32 = out_f32;
    p.w = L.w;
    p.bias = L.bias;
    p.items = items_for(boards);
    p.boards = boards;
    p.slabs_in = L.slabs_in;
    p.slabs_out = (L.n_total + 63) / 64;
    p.ksteps = L.ksteps;
    p.n_total = L.n_total;
    p.n_valid = L.n_valid;
    p.ntaps = L.ntaps;
    p.relu = L.relu;
    {
        static int skew = -1;
        if (skew < 0) {
            const char* e = getenv("KB_CONV_SKEW");
            skew = e ? atoi(e) : 0;
        }
        p.skew = skew;
    }
    p.ts = g_conv_ts ? g_conv_ts + 16 * (g_conv_ts_slot++ % 8) : nullptr;
    if (L.n_tile == 128 && out && use_pair_kernel()) return launch_conv2<128>(p, st);
    if (L.n_tile == 64) return launch_conv<64>(p, st);
    if (L.n_tile == 128) return launch_conv<128>(p, st);
    if (L.n_tile == 80) return launch_conv<80>(p, st);
    set_error("no conv kernel for n_tile=%d", L.n_tile);
    return KB_ERR_UNSUPPORTED;
}

static int ws_stage_logits(NetWs& ws, int batch, float** out, cudaStream_t st) {
    if (batch > ws.logits_cap) {
        cudaStreamSynchronize(st);
        cudaFree(ws.logits);
        ws.logits = nullptr;
        ws.logits_cap = 0;
        if (cudaMalloc(&ws.logits, sizeof(float) * KB_PSIZE * (size_t)batch) != cudaSuccess) {
            set_error("out of device memory for the logits scratch");
            return 1;
        }
        ws.logits_cap = batch;
    }
    *out = ws.logits;
    return 0;
}

struct LegalRef {
    const uint8_t* act = nullptr;
    const uint8_t* nact = nullptr;
    size_t stride = 0;
    float* prior = nullptr;
    const unsigned* sync = nullptr;
    unsigned sync_target = 0;
};

// item0: first item of the activation workspace (X / Y / H) this forward may use, so that forwards of disjoint
// groups of boards can run concurrently on different streams.  The workspace must already hold batch + item0 * NB
// boards (ws_reserve by its owner, before any group is launched).
static int net_forward_impl(kb_net* net, NetWs& ws, const void* planes, int batch, float* policy_dev, float* value256_dev, const LegalRef& lr,
                            cudaStream_t st, int item0) {
    if (!net->loaded) {
        set_error("network weights not loaded");
        return KB_ERR_STATE;
    }
    if (batch + item0 * NB > ws.cap_boards || (!net->fused && ws.slabs < (net->filters + 63) / 64)) {
        set_error("activation workspace holds %d boards, forward needs %d", ws.cap_boards, batch + item0 * NB);
        return KB_ERR_STATE;
    }
    const uint4* in = reinterpret_cast<const uint4*>(planes);
    if (net->fused) {
        static bool configured[16] = {};  // function attributes are per device
        static const int epi_warps = [] {
            const char* e = getenv("KB_TOWER_EPI_WARPS");
            return e && atoi(e) == 8 ? 8 : 16;
        }();
        const int pipe = net->tower_pipe;
        if (!configured[net->device]) {
            KB_CUDA(cudaFuncSetAttribute(k_tower64p, cudaFuncAttributeMaxDynamicSharedMemorySize, FZ_SMEM));
            KB_CUDA(cudaFuncSetAttribute(k_tower64<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, FZ_SMEM));
            KB_CUDA(cudaFuncSetAttribute(k_tower64<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, FZ_SMEM));
            configured[net->device] = true;
        }
        FusedParams fp = net->fp;
        fp.planes = in;
        fp.policy = policy_dev;
        fp.value256 = value256_dev;
        fp.nan_flag = ws.nan_flag;
        fp.ts = net->ts_dev;
        fp.legal_act = lr.act;
        fp.legal_n = lr.nact;
        fp.legal_stride = lr.stride;
        fp.prior = lr.prior;
        fp.sync = lr.sync;
        fp.sync_target = lr.sync_target;
        fp.items = items_for(batch);
        fp.boards = batch;
        const int grid = fp.items < sm_count() ? fp.items : sm_count();
        if (pipe && !lr.sync) KB_CUDA(launch_pdl(1, k_tower64p, dim3(grid), dim3(FP_THREADS), FZ_SMEM, st, fp));
        else if (epi_warps == 16) KB_CUDA(launch_pdl(1, k_tower64<16>, dim3(grid), dim3(128 + 32 * 16), FZ_SMEM, st, fp));
        else KB_CUDA(launch_pdl(1, k_tower64<8>, dim3(grid), dim3(128 + 32 * 8), FZ_SMEM, st, fp));
        return KB_OK;
    }
    int r;
    float* logits = policy_dev;
    if (lr.sync) {
        set_error("split select needs the fused tower");
        return KB_ERR_UNSUPPORTED;
    }
    if (lr.act) {  // the dense logits are scratch in legal mode
        if (item0 != 0) {
            set_error("legal-gather forward of the per-layer path does not support workspace groups");
            return KB_ERR_UNSUPPORTED;
        }
        if (ws_stage_logits(ws, batch, &logits, st)) return KB_ERR_CUDA;
    }
    size_t li = 0;
    const int fs = (net->filters + 63) / 64;
    uint4* X = ws.X + (size_t)item0 * fs * SLAB_U4;
    uint4* Y = ws.Y + (size_t)item0 * fs * SLAB_U4;
    uint4* H = ws.H + (size_t)item0 * 2 * SLAB_U4;
    if ((r = run_conv(net->layers[li++], in, X, nullptr, nullptr, batch, st))) return r;
    for (int i = 0; i < net->residuals; ++i) {
        if ((r = run_conv(net->layers[li++], X, Y, nullptr, nullptr, batch, st))) return r;
        if ((r = run_conv(net->layers[li++], Y, X, X, nullptr, batch, st))) return r;  // x = skip + relu(...)
    }
    if ((r = run_conv(net->layers[li++], X, H, nullptr, nullptr, batch, st))) return r;
    if ((r = run_conv(net->layers[li++], H, nullptr, nullptr, logits, batch, st))) return r;
    if (lr.act)
        k_legal_prior<<<(batch + 7) / 8, 256, 0, st>>>(logits, batch, lr.act, lr.nact, lr.stride, lr.prior, ws.nan_flag);
    else
        k_softmax<<<batch, 256, 0, st>>>(logits, batch, ws.nan_flag);
    KB_CUDA(cudaGetLastError());
    k_value_head<<<batch, 256, 0, st>>>(X, fs, batch, net->wv, net->bv, net->fct, net->fcb, value256_dev, ws.nan_flag);
    KB_CUDA(cudaGetLastError());
    return KB_OK;
}

int net_forward_async(kb_net* net, NetWs& ws, const void* planes, int batch, float* policy_dev, float* value256_dev, cudaStream_t st, int item0) {
    return net_forward_impl(net, ws, planes, batch, policy_dev, value256_dev, LegalRef{}, st, item0);
}

bool net_is_fused(kb_net* net) { return net->fused; }

int net_forward_legal_async(kb_net* net, NetWs& ws, const void* planes, int batch, const void* act_base, const void* nact_base, size_t stride,
                            float* prior_dev, float* value256_dev, cudaStream_t st, int item0, const unsigned* sync, unsigned sync_target) {
    LegalRef lr;
    lr.sync = sync;
    lr.sync_target = sync_target;
    lr.act = reinterpret_cast<const uint8_t*>(act_base);
    lr.nact = reinterpret_cast<const uint8_t*>(nact_base);
    lr.stride = stride;
    lr.prior = prior_dev;
    return net_forward_impl(net, ws, planes, batch, nullptr, value256_dev, lr, st, item0);
}

}  // namespace kb

namespace {

struct BlobCursor {
    const float* p;
    size_t left;
    const float* take(size_t n) {
        if (n > left) return nullptr;
        const float* r = p;
        p += n;
        left -= n;
        return r;
    }
};

// BN (eval, eps 1e-5, nn.cpp:116) folded into the preceding conv: w' = w*s, b' = (b-mean)*s + beta
void fold_bn(const float* g, const float* beta, const float* mean, const float* var, int c, std::vector<float>& scale, std::vector<float>& shift) {
    scale.resize(c);
    shift.resize(c);
    for (int i = 0; i < c; ++i) {
        const float s = g[i] / sqrtf(var[i] + 1e-5f);
        scale[i] = s;
        shift[i] = beta[i] - mean[i] * s;
    }
}

// Packs conv weights [O][Cin][k][k] (scaled per output channel) into the kernel's B layout:
// blocks ordered (pass, slab, tap); a block is [n in tile][64 input channels] bf16, one 128-byte
// line per output channel, 128B-swizzled (chunk j of line n at slot j ^ (n & 7)) -- the canonical
// K-major SWIZZLE_128B operand, copied verbatim into 1024-byte aligned shared memory.
int pack_conv(Layer& L, const float* w, const float* b, int O, int Cin, int k, const std::vector<float>* scale, const std::vector<float>* shift) {
    const int kslices = L.slabs_in, npass = L.n_total / L.n_tile, taps = k * k;
    const size_t block = (size_t)L.n_tile * 64;
    std::vector<uint16_t> packed((size_t)npass * kslices * taps * block, 0);
    for (int pass = 0; pass < npass; ++pass)
        for (int ks = 0; ks < kslices; ++ks)
            for (int tap = 0; tap < taps; ++tap) {
                uint16_t* dst = packed.data() + ((size_t)(pass * kslices + ks) * taps + tap) * block;
                for (int n = 0; n < L.n_tile; ++n)
                    for (int c = 0; c < 64; ++c) {
                        const int o = pass * L.n_tile + n, ci = ks * 64 + c;
                        float v = 0.0f;
                        if (o < O && ci < Cin) {
                            v = w[((size_t)o * Cin + ci) * taps + tap];
                            if (scale) v *= (*scale)[o];
                        }
                        dst[(size_t)n * 64 + (((c >> 3) ^ (n & 7)) << 3) + (c & 7)] = f2bf(v);
                    }
            }
    std::vector<float> bias(L.n_total, 0.0f);
    for (int o = 0; o < O; ++o) bias[o] = scale ? b[o] * (*scale)[o] + (*shift)[o] : b[o];
    KB_CUDA(cudaMalloc(&L.w, packed.size() * 2));
    KB_CUDA(cudaMemcpy(L.w, packed.data(), packed.size() * 2, cudaMemcpyHostToDevice));
    KB_CUDA(cudaMalloc(&L.bias, bias.size() * 4));
    KB_CUDA(cudaMemcpy(L.bias, bias.data(), bias.size() * 4, cudaMemcpyHostToDevice));
    L.hw.swap(packed);
    L.hbias.swap(bias);
    return KB_OK;
}

}  // namespace

extern "C" {

int kb_net_create(kb_net** out, int filters, int residuals) {
    KB_REQUIRE_INIT();
    KB_ARG(out, "out");
    // (the same set kb_trainer_create takes: output-channel passes are 64 or 128 wide, so 192 has no kernel on either side)
    KB_ARG(filters == 64 || filters == 128 || filters == 256, "filters must be 64, 128 or 256");
    KB_ARG(residuals >= 0 && residuals <= 64, "residuals in [0, 64]");
    kb_net* n = new (std::nothrow) kb_net();
    if (!n) return KB_ERR_ARG;
    n->filters = filters;
    n->residuals = residuals;
    n->device = current_device();
    static std::atomic<unsigned long long> next_uid{1};
    n->uid = next_uid.fetch_add(1);
    *out = n;
    return KB_OK;
}

static void ctx_free(InferCtx* c) {
    ws_free(c->ws);
    cudaFree(c->obs_dev); cudaFree(c->pol_dev); cudaFree(c->val_dev);
    cudaFreeHost(c->obs_pin); cudaFreeHost(c->pol_pin); cudaFreeHost(c->val_pin);
    if (c->ev[0]) cudaEventDestroy(c->ev[0]);
    if (c->ev[1]) cudaEventDestroy(c->ev[1]);
    delete c;
}

int kb_net_destroy(kb_net* n) {
    if (!n) return KB_OK;
    if (bind_device(n->device) == KB_OK) cudaDeviceSynchronize();
    for (auto& L : n->layers) {
        cudaFree(L.w);
        cudaFree(L.bias);
    }
    cudaFree(n->wv); cudaFree(n->fct); cudaFree(n->fcb);
    for (InferCtx* c : n->ctx_free) ctx_free(c);
    for (InferCtx* c : n->ctx_kept) ctx_free(c);
    cudaFree(n->fused_w); cudaFree(n->fused_bias); cudaFree(n->ts_dev);
    delete n;
    return KB_OK;
}

size_t kb_net_blob_floats(int F, int R) {
    size_t conv1 = (size_t)F * 30 * 9 + F + 4 * F;
    size_t res = 2 * ((size_t)F * F * 9 + F + 4 * F);
    size_t pol = (size_t)128 * F + 128 + 4 * 128 + (size_t)73 * 128 + 73;
    size_t val = (size_t)F + 1 + 4 + 256 * 64 + 256;
    return conv1 + R * res + pol + val;
}

// Blob order = oracle/nn_oracle.py:param_order (reference module names, nn.cpp:20-23, 45-56).
int kb_net_load_blob(kb_net* net, const float* blob, size_t n_floats) {
    KB_REQUIRE_INIT();
    KB_ARG(net && blob, "net/blob");
    const int F = net->filters, R = net->residuals;
    if (n_floats != kb_net_blob_floats(F, R)) {
        set_error("weight blob has %zu floats, expected %zu for filters=%d residuals=%d", n_floats, kb_net_blob_floats(F, R), F, R);
        return KB_ERR_ARG;
    }
    {
        int rb = bind_device(net->device);
        if (rb) return rb;
    }
    // NN::read / NN::train replace the weights under the exclusive lock (nn.cpp:206, 226): every call that reads them
    // holds the lock shared until its kernels have completed, so nothing is in flight once the lock is ours; the device
    // synchronisation covers callers of the asynchronous kb_net_forward_dev.
    std::unique_lock<WeightLock> wlock(net->mu);
    cudaDeviceSynchronize();
    for (auto& L : net->layers) {
        cudaFree(L.w);
        cudaFree(L.bias);
    }
    net->layers.clear();
    net->loaded = false;
    BlobCursor c{blob, n_floats};
    std::vector<float> sc, sh;
    const int ntile = F == 64 ? 64 : 128;
    auto conv_bn = [&](int O, int Cin, int k, int slabs_in, int ksteps, int n_tile, int n_total, int relu, bool has_bn) -> int {
        const float* w = c.take((size_t)O * Cin * k * k);
        const float* b = c.take(O);
        Layer L;
        L.slabs_in = slabs_in;
        L.ksteps = ksteps;
        L.n_tile = n_tile;
        L.n_total = n_total;
        L.n_valid = O;
        L.ntaps = k * k;
        L.relu = relu;
        int r;
        if (has_bn) {
            const float *g = c.take(O), *beta = c.take(O), *mean = c.take(O), *var = c.take(O);
            fold_bn(g, beta, mean, var, O, sc, sh);
            r = pack_conv(L, w, b, O, Cin, k, &sc, &sh);
        } else
            r = pack_conv(L, w, b, O, Cin, k, nullptr, nullptr);
        if (r) return r;
        net->layers.push_back(L);
        return KB_OK;
    };
    int r;
    const int fs = F / 64;
    if ((r = conv_bn(F, 30, 3, 1, 2, ntile, F, 1, true))) return r;                   // conv1 + batchnorm1 + relu (32 of 64 input channels)
    for (int i = 0; i < R; ++i) {
        if ((r = conv_bn(F, F, 3, fs, 4, ntile, F, 1, true))) return r;               // residual conv1 + bn1 + relu
        if ((r = conv_bn(F, F, 3, fs, 4, ntile, F, 1, true))) return r;               // residual conv2 + bn2 + relu (+ skip)
    }
    if ((r = conv_bn(128, F, 1, fs, 4, ntile, 128, 1, true))) return r;               // policyconv + pbatchnorm + relu
    if ((r = conv_bn(73, 128, 1, 2, 4, 80, 80, 0, false))) return r;                  // policyconv2 (logits)
    std::vector<float> wv_host;  // (the fused tower also keeps the value conv's weights in shared memory)
    {   // value head
        const float* w = c.take(F);
        const float* b = c.take(1);
        const float *g = c.take(1), *beta = c.take(1), *mean = c.take(1), *var = c.take(1);
        const float* fw = c.take(256 * 64);
        const float* fb = c.take(256);
        const float s = g[0] / sqrtf(var[0] + 1e-5f);
        std::vector<float> wv(F), fct(64 * 256);
        for (int i = 0; i < F; ++i) wv[i] = w[i] * s;
        net->bv = (b[0] - mean[0]) * s + beta[0];
        for (int j = 0; j < 256; ++j)
            for (int p = 0; p < 64; ++p) fct[p * 256 + j] = fw[j * 64 + p];
        cudaFree(net->wv); cudaFree(net->fct); cudaFree(net->fcb);
        KB_CUDA(cudaMalloc(&net->wv, F * 4));
        KB_CUDA(cudaMalloc(&net->fct, 64 * 256 * 4));
        KB_CUDA(cudaMalloc(&net->fcb, 256 * 4));
        KB_CUDA(cudaMemcpy(net->wv, wv.data(), F * 4, cudaMemcpyHostToDevice));
        KB_CUDA(cudaMemcpy(net->fct, fct.data(), 64 * 256 * 4, cudaMemcpyHostToDevice));
        KB_CUDA(cudaMemcpy(net->fcb, fb, 256 * 4, cudaMemcpyHostToDevice));
        wv_host = wv;
    }
    // fused single-kernel path: 64 filters, biases fit the kernel's shared-memory table
    net->fused = false;
    cudaFree(net->fused_w);
    cudaFree(net->fused_bias);
    net->fused_w = nullptr;
    net->fused_bias = nullptr;
    {   // KB_TOWER_PIPE is read when the weights are loaded, so one process can hold both variants (tests compare them)
        const char* e = getenv("KB_TOWER_PIPE");
        net->tower_pipe = e ? atoi(e) : KB_TOWER_PIPE_DEFAULT;
    }
    const char* nofuse = getenv("KB_NO_FUSED_TOWER");
    if (F == 64 && R <= 6 && !(nofuse && nofuse[0] == '1')) {
        std::vector<uint16_t> allw;
        std::vector<float> allb;
        FusedParams& fp = net->fp;
        memset(&fp, 0, sizeof(fp));
        {   // (read per load like KB_TOWER_PIPE, so a test can compare both issue forms)
            const char* w = getenv("KB_TOWER_WS");
            fp.mma_ws = w ? atoi(w) : 1;
            const char* ge = getenv("KB_TOWER_GATHER");
            fp.gather = ge ? atoi(ge) : 0;
            const char* g = getenv("KB_TOWER_WAIT_GROUP");
            fp.wait_group = g ? atoi(g) : 3;
            if (fp.wait_group < 1) fp.wait_group = 1;
            if (fp.wait_group > FZ_NSTAGE) fp.wait_group = FZ_NSTAGE;
        }
        const int XOFF = 0, YOFF = SLAB_BYTES;
        fp.n_layers = (int)net->layers.size();
        fp.tower_layers = 1 + 2 * R;
        for (int l = 0; l < fp.n_layers; ++l) {
            const Layer& L = net->layers[l];
            FusedLayer& f = fp.layer[l];
            f.slabs_in = L.slabs_in;
            f.ksteps = L.ksteps;
            f.n = L.n_total;
            f.n_sub = L.n_tile;
            f.ntaps = L.ntaps;
            f.relu = L.relu;
            f.skip = 0;
            f.kind = 0;
            f.bias_off = (int)allb.size();
            f.idesc = ptx::idesc_bf16(128, L.n_tile);
            if (l == 0) {                       // conv1: P (parked in Y) -> X
                f.src_off = YOFF;
                f.dst_off = XOFF;
            } else if (l < fp.tower_layers) {
                const bool first = (l - 1) % 2 == 0;  // residual conv1: X -> Y, conv2: Y -> X (+skip)
                f.src_off = first ? XOFF : YOFF;
                f.dst_off = first ? YOFF : XOFF;
                f.skip = first ? 0 : 1;
            } else if (l == fp.tower_layers) {  // policyconv: X -> H (overlays X|Y)
                f.src_off = XOFF;
                f.dst_off = 0;
            } else {                            // policyconv2: H -> fp32 logits
                f.src_off = 0;
                f.dst_off = 0;
                f.kind = 1;
            }
            allw.insert(allw.end(), L.hw.begin(), L.hw.end());
            allb.insert(allb.end(), L.hbias.begin(), L.hbias.end());
        }
        fp.blocks_per_item = fp.blocks_before_logits = 0;
        for (int l = 0; l < fp.n_layers; ++l) {
            const FusedLayer& f = fp.layer[l];
            if (l == fp.n_layers - 1) fp.blocks_before_logits = fp.blocks_per_item;
            fp.blocks_per_item += f.slabs_in * f.ntaps * (f.n / f.n_sub);
        }
        if (fp.blocks_per_item - fp.blocks_before_logits != 2) fp.gather = 0;  // (the gather head reads exactly two K slabs)
        fp.wv_off = (int)allb.size();
        allb.insert(allb.end(), wv_host.begin(), wv_host.end());
        fp.n_bias = (int)allb.size();
        if (fp.n_bias <= 1340 && fp.n_layers <= 16 && fp.wv_off % 4 == 0) {
            KB_CUDA(cudaMalloc(&net->fused_w, allw.size() * 2));
            KB_CUDA(cudaMemcpy(net->fused_w, allw.data(), allw.size() * 2, cudaMemcpyHostToDevice));
            KB_CUDA(cudaMalloc(&net->fused_bias, allb.size() * 4));
            KB_CUDA(cudaMemcpy(net->fused_bias, allb.data(), allb.size() * 4, cudaMemcpyHostToDevice));
            fp.w = net->fused_w;
            fp.bias = net->fused_bias;
            fp.wv = net->wv;
            fp.fct = net->fct;
            fp.fcb = net->fcb;
            fp.bv = net->bv;
            net->fused = true;
        }
    }
    // the uploads above went through the legacy stream, which does not order with the (non-blocking) main stream
    KB_CUDA(cudaDeviceSynchronize());
    net->loaded = true;
    return KB_OK;
}

size_t kb_net_planes_bytes(int batch) { return act_bytes(batch, IN_SLABS); }

// Asynchronous, device-resident: the caller keeps planes / outputs in HBM.  The activation workspace is a per-thread
// context of the net (returned to the free list right away: the next forward of the same thread on the same stream is
// ordered behind this one).  The caller must not load new weights while forwards it has not waited for are in flight.
static InferCtx* ctx_take(kb_net* net) {
    std::lock_guard<std::mutex> g(net->ctx_mu);
    if (!net->ctx_free.empty()) {
        InferCtx* c = net->ctx_free.back();
        net->ctx_free.pop_back();
        return c;
    }
    return new (std::nothrow) InferCtx();
}
static void ctx_give(kb_net* net, InferCtx* c) {
    std::lock_guard<std::mutex> g(net->ctx_mu);
    net->ctx_free.push_back(c);
}
static thread_local InferCtx* tl_dev_ctx = nullptr;  // kb_net_forward_dev: one workspace per host thread, kept
static thread_local unsigned long long tl_dev_ctx_uid = 0;

int kb_net_forward_dev(kb_net* net, const void* planes_dev, int batch, float* policy_dev, float* value256_dev) {
    KB_REQUIRE_INIT();
    KB_ARG(net && planes_dev && policy_dev && value256_dev && batch > 0, "net/planes/policy/value/batch");
    int r = bind_device(net->device);
    if (r) return r;
    NetReadGuard lock(net);
    if (tl_dev_ctx_uid != net->uid || !tl_dev_ctx) {
        tl_dev_ctx = ctx_take(net);  // (kept by this thread for the life of the net: device-resident benches call this in loops)
        if (!tl_dev_ctx) return KB_ERR_ARG;
        tl_dev_ctx_uid = net->uid;
        std::lock_guard<std::mutex> g(net->ctx_mu);
        net->ctx_kept.push_back(tl_dev_ctx);
    }
    if ((r = ws_reserve(tl_dev_ctx->ws, net, batch, main_stream()))) return r;
    return net_forward_async(net, tl_dev_ctx->ws, planes_dev, batch, policy_dev, value256_dev, main_stream());
}

static int ctx_stage(InferCtx* c, int batch, cudaStream_t st) {
    if (batch <= c->stage_cap) return KB_OK;
    cudaStreamSynchronize(st);
    cudaFree(c->obs_dev); cudaFree(c->pol_dev); cudaFree(c->val_dev);
    cudaFreeHost(c->obs_pin); cudaFreeHost(c->pol_pin); cudaFreeHost(c->val_pin);
    c->obs_dev = c->pol_dev = c->val_dev = c->obs_pin = c->pol_pin = c->val_pin = nullptr;
    c->stage_cap = 0;
    KB_CUDA(cudaMalloc(&c->obs_dev, sizeof(float) * KB_OBSIZE * (size_t)batch));
    KB_CUDA(cudaMalloc(&c->pol_dev, sizeof(float) * KB_PSIZE * (size_t)batch));
    KB_CUDA(cudaMalloc(&c->val_dev, sizeof(float) * KB_VALUE_WIDTH * (size_t)batch));
    KB_CUDA(cudaMallocHost(&c->obs_pin, sizeof(float) * KB_OBSIZE * (size_t)batch));
    KB_CUDA(cudaMallocHost(&c->pol_pin, sizeof(float) * KB_PSIZE * (size_t)batch));
    KB_CUDA(cudaMallocHost(&c->val_pin, sizeof(float) * KB_VALUE_WIDTH * (size_t)batch));
    if (!c->ev[0]) {
        KB_CUDA(cudaEventCreateWithFlags(&c->ev[0], cudaEventDisableTiming));
        KB_CUDA(cudaEventCreateWithFlags(&c->ev[1], cudaEventDisableTiming));
    }
    c->stage_cap = batch;
    return KB_OK;
}

// true when the CUDA runtime can DMA straight from / into `p` (cudaMallocHost / cudaHostRegister memory)
static bool host_is_pinned(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost;
}

// Host buffers in, host buffers out: observations [batch][1920] -> policy [batch][4672] (optional), value rows
// ([batch][256] when value256, else the first `batch` floats of that tensor = NN::infer's value[i] = vh.flat[i], Q1).
// Pinned caller buffers are used directly; pageable ones are staged through this call's private pinned buffers in
// pieces, so the CPU copy of one piece overlaps the DMA of the previous one.  Everything -- workspace, staging, stream
// (the calling thread's) -- is private to the call: any number of host threads may be inside at once (nn.cpp:164-168).
static int net_infer_host(kb_net* net, const float* obs, int batch, float* policy, float* value, bool value256, bool keep_dbg) {
    KB_REQUIRE_INIT();
    KB_ARG(net && obs && batch > 0, "net/obs/batch");
    int r = bind_device(net->device);
    if (r) return r;
    NetReadGuard lock(net);
    InferCtx* c = ctx_take(net);
    if (!c) return KB_ERR_ARG;
    struct Give {
        kb_net* n;
        InferCtx* c;
        ~Give() { ctx_give(n, c); }
    } give{net, c};
    cudaStream_t st = main_stream();
    if ((r = ctx_stage(c, batch, st)) || (r = ws_reserve(c->ws, net, batch, st))) return r;
    const size_t obs_b = sizeof(float) * KB_OBSIZE * (size_t)batch, pol_b = sizeof(float) * KB_PSIZE * (size_t)batch;
    const size_t val_b = sizeof(float) * (value256 ? KB_VALUE_WIDTH * (size_t)batch : (size_t)batch);
    constexpr size_t PIECE = 1 << 20;
    if (host_is_pinned(obs)) {
        KB_CUDA(cudaMemcpyAsync(c->obs_dev, obs, obs_b, cudaMemcpyHostToDevice, st));
    } else {
        for (size_t off = 0; off < obs_b; off += PIECE) {
            const size_t nb = obs_b - off < PIECE ? obs_b - off : PIECE;
            memcpy((char*)c->obs_pin + off, (const char*)obs + off, nb);
            KB_CUDA(cudaMemcpyAsync((char*)c->obs_dev + off, (char*)c->obs_pin + off, nb, cudaMemcpyHostToDevice, st));
        }
    }
    if ((r = obs_to_tall_launch(c->obs_dev, batch, c->ws.P, st))) return r;
    if ((r = net_forward_async(net, c->ws, c->ws.P, batch, c->pol_dev, c->val_dev, st))) return r;
    int flag = 0;
    const bool pol_direct = policy && host_is_pinned(policy), val_direct = value && host_is_pinned(value);
    if (value) KB_CUDA(cudaMemcpyAsync(val_direct ? value : c->val_pin, c->val_dev, val_b, cudaMemcpyDeviceToHost, st));
    if (policy && pol_direct) KB_CUDA(cudaMemcpyAsync(policy, c->pol_dev, pol_b, cudaMemcpyDeviceToHost, st));
    if (policy && !pol_direct) {
        // D2H in pieces on the stream; the CPU copies piece i out of the pinned buffer while piece i + 1 is still arriving
        size_t done = 0;
        for (size_t off = 0; off < pol_b; off += PIECE) {
            const size_t nb = pol_b - off < PIECE ? pol_b - off : PIECE;
            KB_CUDA(cudaMemcpyAsync((char*)c->pol_pin + off, (char*)c->pol_dev + off, nb, cudaMemcpyDeviceToHost, st));
            KB_CUDA(cudaEventRecord(c->ev[(off / PIECE) & 1], st));
            if (off >= PIECE) {
                KB_CUDA(cudaEventSynchronize(c->ev[((off / PIECE) - 1) & 1]));
                memcpy((char*)policy + done, (char*)c->pol_pin + done, PIECE);
                done += PIECE;
            }
        }
        KB_CUDA(cudaStreamSynchronize(st));
        memcpy((char*)policy + done, (char*)c->pol_pin + done, pol_b - done);
    }
    KB_CUDA(cudaMemcpyAsync(&flag, c->ws.nan_flag, sizeof(int), cudaMemcpyDeviceToHost, st));
    KB_CUDA(cudaStreamSynchronize(st));
    if (value && !val_direct) memcpy(value, c->val_pin, val_b);
    if (keep_dbg) net->dbg_ctx = c;
    if (flag) {
        KB_CUDA(cudaMemsetAsync(c->ws.nan_flag, 0, sizeof(int), st));
        KB_CUDA(cudaStreamSynchronize(st));
        set_error("inference output contains NaN");
        return KB_ERR_NAN;
    }
    return KB_OK;
}

int kb_net_forward_full(kb_net* net, const float* obs, int batch, float* policy, float* value256) {
    return net_infer_host(net, obs, batch, policy, value256, true, true);
}

// NN::infer (nn.cpp:155-187).  The reference copies the first `batch` floats of its [batch,256]
// value tensor, i.e. value[i] = vh[i / 256][i % 256]; reproduced bit for bit.
int kb_net_infer(kb_net* net, const float* obs, int batch, float* policy, float* value) {
    KB_ARG(value && policy, "policy/value");
    return net_infer_host(net, obs, batch, policy, value, false, false);
}

// Test hook: download one board of an internal activation tensor as fp32 [channels][64].
// which: 0 input planes (32 ch), 1 tower output X, 2 residual scratch Y, 3 policy hidden H (128 ch).
int kb_net_debug_activation(kb_net* net, int which, int board, float* out, int* channels) {
    KB_REQUIRE_INIT();
    KB_ARG(net && out && channels && board >= 0, "net/out/board");
    KB_ARG(net->dbg_ctx && board < net->dbg_ctx->ws.cap_boards, "no kb_net_forward_full before this call / board out of range");
    KB_ARG(which == 0 || net->dbg_ctx->ws.slabs > 0, "the fused tower keeps X / Y / H in shared memory (only which = 0 exists)");
    const NetWs& ws = net->dbg_ctx->ws;
    const uint4* buf = which == 0 ? ws.P : which == 1 ? ws.X : which == 2 ? ws.Y : ws.H;
    const int slabs = which == 0 ? IN_SLABS : which == 3 ? 2 : (net->filters + 63) / 64;
    const int chunks = slabs * 8;
    KB_CUDA(cudaStreamSynchronize(main_stream()));
    const int item = board / NB, slot = board % NB;
    std::vector<uint16_t> slab((size_t)SLAB_BYTES / 2);
    for (int sl = 0; sl < slabs; ++sl) {
        KB_CUDA(cudaMemcpy(slab.data(), buf + ((size_t)item * slabs + sl) * SLAB_U4, SLAB_BYTES, cudaMemcpyDeviceToHost));
        for (int q = 0; q < 64; ++q)
            for (int c = 0; c < 64; ++c) {
                const int px = tall_pixel(slot, q);
                uint32_t u = (uint32_t)slab[(size_t)chunk_u4(px, c >> 3) * 8 + (c & 7)] << 16;
                float f;
                memcpy(&f, &u, 4);
                out[(size_t)(sl * 64 + c) * 64 + q] = f;
            }
    }
    *channels = chunks * 8;
    return KB_OK;
}

// Profiling hook: clock64 stamps of CTA 0's first epilogue thread in k_tower64 (start, region
// cleared, then per layer [accumulators ready, layer written], softmax start, end).
int kb_net_debug_timestamps(kb_net* net, int enable, long long* out, int cap, int* count) {
    KB_REQUIRE_INIT();
    KB_ARG(net, "net");
    if (enable && !net->ts_dev) {
        KB_CUDA(cudaMalloc(&net->ts_dev, (128 + 2 * 160) * sizeof(long long)));
        KB_CUDA(cudaMemsetAsync(net->ts_dev, 0, (128 + 2 * 160) * sizeof(long long), main_stream()));
    }
    kb::g_conv_ts = enable && !net->fused ? net->ts_dev : nullptr;  // per-layer kernels: MMA-thread wait counters
    if (!out) kb::g_conv_ts_slot = 0;
    if (out && net->ts_dev) {
        long long h[128];
        KB_CUDA(cudaStreamSynchronize(main_stream()));
        KB_CUDA(cudaMemcpy(h, net->ts_dev, sizeof(h), cudaMemcpyDeviceToHost));
        int n = 0;
        if (cap >= 128) {  // raw: stamps in [0, 60), the MMA warp's per-layer {weight wait, issue span} pairs from 64
            for (; n < 128; ++n) out[n] = h[n];
        } else {
            while (n < (kb::g_conv_ts ? 128 : 60) && n < cap && (h[n] || kb::g_conv_ts)) { out[n] = h[n]; ++n; }
        }
        if (count) *count = n;
    }
    if (!enable && net->ts_dev) {
        cudaFree(net->ts_dev);
        net->ts_dev = nullptr;
    }
    return KB_OK;
}

// profiling hook: globaltimer (ns) at entry and exit of every CTA of the last k_tower64 launch, [cta][2]
int kb_net_debug_cta_spans(kb_net* net, long long* out, int cap_ctas, int* count) {
    KB_REQUIRE_INIT();
    KB_ARG(net && out && count && cap_ctas > 0, "net/out/count");
    *count = 0;
    if (!net->ts_dev) return KB_OK;
    KB_CUDA(cudaStreamSynchronize(main_stream()));
    const int n = cap_ctas < 160 ? cap_ctas : 160;
    KB_CUDA(cudaMemcpy(out, net->ts_dev + 128, sizeof(long long) * 2 * (size_t)n, cudaMemcpyDeviceToHost));
    *count = n;
    return KB_OK;
}

int kb_net_flops(kb_net* net, double* tower, double* heads) {
    KB_ARG(net, "net");
    const double F = net->filters, R = net->residuals;
    if (tower) *tower = 2.0 * 64 * (9 * 30) * F + R * 2.0 * (2.0 * 64 * 9 * F * F);
    if (heads) *heads = 2.0 * 64 * F * 128 + 2.0 * 64 * 128 * 73 + 2.0 * 64 * F + 2.0 * 64 * 256;
    return KB_OK;
}

}  // extern "C"

#include "train.inl"
