// Training step of the policy/value network on sm_100a: NN::train's mini-batch (kami/nn/nn.cpp:224-377)
// = NNModule::forward in training mode (BatchNorm with batch statistics, nn.cpp:59-91, 228),
// NNModule::loss (nn.cpp:93-105), backward, plain SGD (nn.cpp:239-241, 353-354).
//
// Included at the end of net.cu (same translation unit: it launches k_conv2 / k_conv).
//
//   forward   conv (tcgen05 implicit GEMM, raw bf16 weights + bias, no ReLU) -> pre   [bf16, tall layout]
//             batch statistics, y = relu(gamma * xhat + beta) (+ skip)       -> post  [bf16, tall layout]
//   backward  BatchNorm/ReLU backward (two passes: per-channel sums, then dpre)
//             dgrad  = the same conv kernel on flipped / transposed weights (tcgen05)
//             wgrad  = k_wgrad: dW[co][ci][tap] = sum_pixels dpre[p][co] * x[p + tap][ci] as an
//                      MN-major tcgen05 GEMM whose K dimension is the pixel axis of the tall layout:
//                      a pixel line (64 channels, 128 B, swizzled) is one K row of the canonical
//                      MN-major SWIZZLE_128B operand, and a tap is again a start-address offset
//   heads     fp32 CUDA-core kernels (value head, softmax / loss / dlogits): < 1 % of the FLOPs
//
// Master weights, gradients and BatchNorm running statistics are fp32 in the blob order of
// kb_net_load_blob; activations and activation gradients are bf16.
namespace kb {

constexpr float BN_EPS_T = 1e-5f;
constexpr float BN_MOMENTUM_T = 0.1f;

__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        f[2 * k] = __uint_as_float(w[k] << 16);
        f[2 * k + 1] = __uint_as_float(w[k] & 0xffff0000u);
    }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
    return make_uint4(ptx::pack_bf16x2(f[0], f[1]), ptx::pack_bf16x2(f[2], f[3]), ptx::pack_bf16x2(f[4], f[5]), ptx::pack_bf16x2(f[6], f[7]));
}
// uint4 index of channel chunk j (8 channels) of square sq of board b in slab s of a tensor with `slabs` slabs
__device__ __forceinline__ size_t act_idx(int b, int sq, int slabs, int s, int j) {
    const int item = b / NB, slot = b - item * NB;
    return ((size_t)item * slabs + s) * SLAB_U4 + chunk_u4(tall_pixel(slot, sq), j);
}

// Thread layout of the per-channel kernels: 256 threads = 32 pixel lanes x 8 chunks of one slab
// (blockIdx.x = slab, blockIdx.y strides over the batch's real pixels).  Sums the 8 channels a
// thread owns over its pixels, then over the block, then atomically into out[0..1][C].
__device__ __forceinline__ void block_channel_sums(float (&a)[8], float (&q)[8], float* out0, float* out1, int ch_base) {
    __shared__ float red[8][8][16];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, j = threadIdx.x & 7;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        a[k] += __shfl_xor_sync(0xffffffffu, a[k], 8);
        a[k] += __shfl_xor_sync(0xffffffffu, a[k], 16);
        q[k] += __shfl_xor_sync(0xffffffffu, q[k], 8);
        q[k] += __shfl_xor_sync(0xffffffffu, q[k], 16);
    }
    if (lane < 8) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            red[warp][j][k] = a[k];
            red[warp][j][8 + k] = q[k];
        }
    }
    __syncthreads();
    if (threadIdx.x < 128) {
        const int jj = threadIdx.x >> 4, e = threadIdx.x & 15;
        float s = 0.0f;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += red[w][jj][e];
        const int ch = ch_base + jj * 8 + (e & 7);
        atomicAdd((e < 8 ? out0 : out1) + ch, s);
    }
}

constexpr int BN_U = 4;  // pixel lines in flight per thread and trip (8 measured the same and spills)
// ---- BatchNorm forward (training mode) ---------------------------------------------------------
__global__ void __launch_bounds__(256, 2) k_bn_stats(const uint4* pre, int boards, int slabs, float* sum, float* sumsq) {
    const int s = blockIdx.x, j = threadIdx.x & 7, pl = threadIdx.x >> 3;
    float a[8] = {0, 0, 0, 0, 0, 0, 0, 0}, q[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const int npix = boards * 64, stride = gridDim.y * 32;
    for (int g0 = blockIdx.y * 32 + pl; g0 < npix; g0 += BN_U * stride) {  // BN_U independent loads in flight per thread
        uint4 v[BN_U];
#pragma unroll
        for (int u = 0; u < BN_U; ++u) {
            const int g = g0 + u * stride;
            v[u] = g < npix ? pre[act_idx(g >> 6, g & 63, slabs, s, j)] : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int u = 0; u < BN_U; ++u) {
            float f[8];
            unpack8(v[u], f);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                a[k] += f[k];
                q[k] = fmaf(f[k], f[k], q[k]);
            }
        }
    }
    block_channel_sums(a, q, sum, sumsq, s * 64);
}
// post = relu(gamma * xhat + beta) (+ skip): nn.cpp:28-33, 63-65, 72-74
// sum / sumsq: this layer's own accumulators (zeroed once per step).  Every thread derives mean / rstd of its 8
// channels; blocks with blockIdx.y == 0 publish them for the backward pass and update the running statistics
// (momentum 0.1, unbiased variance: LibTorch BatchNorm2d defaults).
__global__ void __launch_bounds__(256, 2) k_bn_apply(const uint4* pre, uint4* post, const uint4* skip, int boards, int slabs, const float* sum,
                                                  const float* sumsq, float n, float* mean, float* rstd, float* run_mean, float* run_var,
                                                  const float* gamma, const float* beta) {
    const int s = blockIdx.x, j = threadIdx.x & 7, pl = threadIdx.x >> 3;
    const int c0 = s * 64 + j * 8;
    float sc[8], sh[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const float m = sum[c0 + k] / n;
        const float var = fmaxf(sumsq[c0 + k] / n - m * m, 0.0f);
        const float rs = rsqrtf(var + BN_EPS_T);
        sc[k] = gamma[c0 + k] * rs;
        sh[k] = beta[c0 + k] - m * sc[k];
        if (blockIdx.y == 0 && pl == 0) {
            mean[c0 + k] = m;
            rstd[c0 + k] = rs;
            run_mean[c0 + k] = (1.0f - BN_MOMENTUM_T) * run_mean[c0 + k] + BN_MOMENTUM_T * m;
            run_var[c0 + k] = (1.0f - BN_MOMENTUM_T) * run_var[c0 + k] + BN_MOMENTUM_T * var * (n / fmaxf(n - 1.0f, 1.0f));
        }
    }
    const int npix = boards * 64, stride = gridDim.y * 32;
    for (int g0 = blockIdx.y * 32 + pl; g0 < npix; g0 += BN_U * stride) {
        size_t idx[BN_U];
        uint4 v[BN_U], sk[BN_U];
#pragma unroll
        for (int u = 0; u < BN_U; ++u) {
            const int g = g0 + u * stride;
            idx[u] = g < npix ? act_idx(g >> 6, g & 63, slabs, s, j) : 0;
            v[u] = pre[idx[u]];
            sk[u] = skip ? skip[idx[u]] : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int u = 0; u < BN_U; ++u) {
            if (g0 + u * stride >= npix) break;
            float f[8], t[8];
            unpack8(v[u], f);
            unpack8(sk[u], t);
#pragma unroll
            for (int k = 0; k < 8; ++k) f[k] = fmaxf(fmaf(f[k], sc[k], sh[k]), 0.0f) + t[k];
            post[idx[u]] = pack8(f);
        }
    }
}

// ---- BatchNorm + ReLU backward -----------------------------------------------------------------
// dy: gradient w.r.t. the ReLU output.  s1 = sum(dy * mask), s2 = sum(dy * mask * xhat), mask = (gamma*xhat+beta > 0)
__global__ void __launch_bounds__(256, 2) k_bn_bwd_reduce(const uint4* dy, const uint4* pre, int boards, int slabs, const float* mean, const float* rstd,
                                                       const float* gamma, const float* beta, float* s1, float* s2) {
    const int s = blockIdx.x, j = threadIdx.x & 7, pl = threadIdx.x >> 3;
    const int c0 = s * 64 + j * 8;
    float mu[8], rs[8], ga[8], be[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        mu[k] = mean[c0 + k];
        rs[k] = rstd[c0 + k];
        ga[k] = gamma[c0 + k];
        be[k] = beta[c0 + k];
    }
    float a[8] = {0, 0, 0, 0, 0, 0, 0, 0}, q[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const int npix = boards * 64, stride = gridDim.y * 32;
    for (int g0 = blockIdx.y * 32 + pl; g0 < npix; g0 += BN_U * stride) {
        uint4 vx[BN_U], vd[BN_U];
#pragma unroll
        for (int u = 0; u < BN_U; ++u) {
            const int g = g0 + u * stride;
            const size_t i = g < npix ? act_idx(g >> 6, g & 63, slabs, s, j) : 0;
            vx[u] = pre[i];
            vd[u] = g < npix ? dy[i] : make_uint4(0, 0, 0, 0);  // a zero gradient adds nothing to either sum
        }
#pragma unroll
        for (int u = 0; u < BN_U; ++u) {
            float x[8], d[8];
            unpack8(vx[u], x);
            unpack8(vd[u], d);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const float xh = (x[k] - mu[k]) * rs[k];
                const float dm = fmaf(ga[k], xh, be[k]) > 0.0f ? d[k] : 0.0f;
                a[k] += dm;
                q[k] = fmaf(dm, xh, q[k]);
            }
        }
    }
    block_channel_sums(a, q, s1, s2, s * 64);
}
// dpre = gamma * rstd * (dy*mask - s1/n - xhat * s2/n)
__global__ void __launch_bounds__(256, 2) k_bn_bwd_apply(const uint4* dy, const uint4* pre, uint4* dpre, int boards, int slabs, const float* mean,
                                                      const float* rstd, const float* gamma, const float* beta, const float* s1, const float* s2, float n,
                                                      float* dgamma, float* dbeta) {
    const int s = blockIdx.x, j = threadIdx.x & 7, pl = threadIdx.x >> 3;
    const int c0 = s * 64 + j * 8;
    float mu[8], rs[8], ga[8], be[8], m1[8], m2[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        mu[k] = mean[c0 + k];
        rs[k] = rstd[c0 + k];
        ga[k] = gamma[c0 + k];
        be[k] = beta[c0 + k];
        m1[k] = s1[c0 + k] / n;
        m2[k] = s2[c0 + k] / n;
        if (blockIdx.y == 0 && pl == 0) {  // dgamma = sum(dy * mask * xhat), dbeta = sum(dy * mask)
            dgamma[c0 + k] = s2[c0 + k];
            dbeta[c0 + k] = s1[c0 + k];
        }
    }
    const int npix = boards * 64, stride = gridDim.y * 32;
    for (int g0 = blockIdx.y * 32 + pl; g0 < npix; g0 += BN_U * stride) {
        size_t idx[BN_U];
        uint4 vx[BN_U], vd[BN_U];
#pragma unroll
        for (int u = 0; u < BN_U; ++u) {
            const int g = g0 + u * stride;
            idx[u] = g < npix ? act_idx(g >> 6, g & 63, slabs, s, j) : 0;
            vx[u] = pre[idx[u]];
            vd[u] = dy[idx[u]];
        }
#pragma unroll
        for (int u = 0; u < BN_U; ++u) {
            if (g0 + u * stride >= npix) break;
            float x[8], d[8], o[8];
            unpack8(vx[u], x);
            unpack8(vd[u], d);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const float xh = (x[k] - mu[k]) * rs[k];
                const float dm = fmaf(ga[k], xh, be[k]) > 0.0f ? d[k] : 0.0f;
                o[k] = ga[k] * rs[k] * (dm - m1[k] - xh * m2[k]);
            }
            dpre[idx[u]] = pack8(o);
        }
    }
}
// ---- weight packing (fp32 master -> bf16 operand blocks of k_conv / k_conv2) -------------------------
// Logical conv: out channel o, in channel ci, tap t.  transpose_flip = 0: w[o][ci][t] (forward);
// 1: w[ci][o][taps-1-t] (the dgrad conv: in/out swapped, kernel rotated by 180 degrees).
struct PackJob {
    const float* w;
    uint16_t* dst;
    int O, I, taps, transpose_flip, n_tile, kslices, npass;
    int block0, nblocks;  // this job's share of the fused launch
};
__device__ __forceinline__ void pack_conv_elems(const PackJob& J, size_t first, size_t step) {
    const int O = J.O, I = J.I, taps = J.taps, transpose_flip = J.transpose_flip, n_tile = J.n_tile, kslices = J.kslices;
    const float* __restrict__ w = J.w;
    uint16_t* __restrict__ dst = J.dst;
    const size_t total = (size_t)J.npass * kslices * taps * n_tile * 64;
    for (size_t e = first; e < total; e += step) {
        const int within = (int)(e % ((size_t)n_tile * 64));
        const size_t blk = e / ((size_t)n_tile * 64);
        const int tap = (int)(blk % taps);
        const int ks = (int)((blk / taps) % kslices);
        const int pass = (int)(blk / ((size_t)taps * kslices));
        const int n = within >> 6, slot = (within >> 3) & 7, k = within & 7;
        const int c = ((slot ^ (n & 7)) << 3) | k;  // the element stored at this position (128B swizzle)
        const int o = pass * n_tile + n, ci = ks * 64 + c;
        const int Ol = transpose_flip ? I : O, Il = transpose_flip ? O : I;
        float v = 0.0f;
        if (o < Ol && ci < Il) v = transpose_flip ? w[((size_t)ci * I + o) * taps + (taps - 1 - tap)] : w[((size_t)o * I + ci) * taps + tap];
        const __nv_bfloat16 h = __float2bfloat16_rn(v);
        dst[e] = *reinterpret_cast<const uint16_t*>(&h);
    }
}
// every conv of the network (forward blocks and, where needed, the transposed / rotated dgrad blocks) in ONE launch:
// the 85 separate 6.6 us launches of the 20x256 net were launch-bound (0.56 ms per training step)
__global__ void __launch_bounds__(256) k_pack_all(const PackJob* jobs, int njobs) {
    int j = 0;
    while (j + 1 < njobs && (int)blockIdx.x >= jobs[j + 1].block0) ++j;
    const PackJob J = jobs[j];
    pack_conv_elems(J, (size_t)((int)blockIdx.x - J.block0) * blockDim.x + threadIdx.x, (size_t)J.nblocks * blockDim.x);
}

// ---- wgrad: dW[co][ci][tap] += sum over pixels dy[p][co] * x[p + shift(tap)][ci] (tcgen05, MN-major operands) ------
struct WgradParams {
    const uint8_t* dy;  // [items][co_slabs] slabs, gradient w.r.t. the conv output (pads and idle boards are zero)
    const uint8_t* x;   // [items][ci_slabs] slabs, the conv input
    float* dw;          // fp32 [O][I][ntaps], accumulated with red.global.add
    float* part;        // k_wgrad_row: per-subset partial sums [subset][tap][ci][co] (reduced by k_wgrad_reduce), or nullptr
    int items, co_slabs, ci_slabs, O, I, ntaps, subsets;
};
constexpr int WG_KC = 64;                      // pixels (K rows) per stage
constexpr int WG_A_SLAB = WG_KC * LINE_BYTES;  // 8192
constexpr int WG_B_SLAB = WG_A_SLAB + 1024;    // room to start at line (p0 & 7): the swizzle phase follows the pixel index
constexpr int WG_STAGE = 4 * WG_A_SLAB + 4 * WG_B_SLAB;
constexpr int WG_NSTAGE = 3;
constexpr int WG_SMEM = 1024 + WG_NSTAGE * WG_STAGE;
static_assert(WG_SMEM <= 232448 && WG_STAGE % 1024 == 0, "wgrad shared memory");

// MN-major SWIZZLE_128B descriptor: 64 MN elements (128 B) per line, consecutive K rows are consecutive
// lines, 8-row K groups sbo bytes apart, 64-element MN atoms lbo bytes apart
__device__ __forceinline__ uint64_t wg_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ULL << 46) |
           (2ULL << 61);
}

__global__ void __launch_bounds__(256, 1) k_wgrad(const WgradParams P) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int warp = ptx::uniform_warp_id(), lane = threadIdx.x & 31;
    const uint32_t bar0 = ptx::smem_u32(smem);
    auto full = [&](int s) { return bar0 + 8u * s; };
    auto empty = [&](int s) { return bar0 + 8u * (4 + s); };
    const uint32_t t_full = bar0 + 8u * 8;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 256);
    const uint32_t stage0 = bar0 + 1024;
    const int tap = (int)blockIdx.x % P.ntaps, subset = (int)blockIdx.x / P.ntaps;
    const int shift = P.ntaps == 9 ? (tap / 3 - 1) * TALL_PITCH + (tap % 3 - 1) : 0;
    const int my_items = P.items > subset ? (P.items - 1 - subset) / P.subsets + 1 : 0;
    constexpr int CHUNKS = SLAB_PIX / WG_KC;  // 10
    const int total = my_items * CHUNKS;
    const int N = 64 * P.ci_slabs, halves = (P.co_slabs + 1) / 2;

    if (threadIdx.x == 0) {
        for (int s = 0; s < WG_NSTAGE; ++s) {
            ptx::mbar_init(full(s), 1);
            ptx::mbar_init(empty(s), 1);
        }
        ptx::mbar_init(t_full, 1);
        ptx::fence_barrier_init();
    }
    if (warp == 2) {
        ptx::tmem_alloc(ptx::smem_u32(tmem_slot), 512);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== producer: per stage, 64 pixel lines of every dy slab and of every x slab (shifted by the tap) =====
        int stage = 0, phase = 0;
        for (int c = 0; c < total; ++c) {
            const int ii = c / CHUNKS, k0 = (c - ii * CHUNKS) * WG_KC;
            const int item = subset + ii * P.subsets;
            const int p0 = k0 + shift, ph = p0 & 7;  // two's complement: also right for p0 < 0
            ptx::mbar_wait(empty(stage), phase ^ 1);
            if (ptx::elect_one()) {
                const uint32_t sa = stage0 + stage * WG_STAGE, sb = sa + 4 * WG_A_SLAB;
                ptx::mbar_arrive_expect_tx(full(stage), (uint32_t)(P.co_slabs + P.ci_slabs) * WG_A_SLAB);
                for (int s = 0; s < P.co_slabs; ++s)
                    ptx::bulk_g2s(sa + s * WG_A_SLAB, P.dy + ((size_t)item * P.co_slabs + s) * SLAB_BYTES + (size_t)k0 * LINE_BYTES, WG_A_SLAB, full(stage));
                for (int s = 0; s < P.ci_slabs; ++s)
                    ptx::bulk_g2s(sb + s * WG_B_SLAB + ph * LINE_BYTES,
                                  P.x + ((ptrdiff_t)((size_t)item * P.ci_slabs + s) * SLAB_BYTES + (ptrdiff_t)p0 * LINE_BYTES), WG_A_SLAB, full(stage));
            }
            __syncwarp();
            if (++stage == WG_NSTAGE) {
                stage = 0;
                phase ^= 1;
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: D[128 co x N ci] (two co halves) accumulate over every pixel of every item =====
        // instruction descriptor: fp32 accumulate, bf16 x bf16, A and B both MN-major (bits 15, 16)
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
        int stage = 0, phase = 0;
        for (int c = 0; c < total; ++c) {
            const int ii = c / CHUNKS, k0 = (c - ii * CHUNKS) * WG_KC;
            const int ph = (k0 + shift) & 7;
            ptx::mbar_wait(full(stage), phase);
            ptx::tc_fence_after();
            if (ptx::elect_one()) {
                const uint32_t sa = stage0 + stage * WG_STAGE, sb = sa + 4 * WG_A_SLAB + ph * LINE_BYTES;
#pragma unroll
                for (int kk = 0; kk < WG_KC / 16; ++kk) {
                    const uint64_t bdesc = wg_desc(sb + kk * 16 * LINE_BYTES, WG_B_SLAB, 1024);
                    for (int h = 0; h < halves; ++h) {
                        const uint64_t adesc = wg_desc(sa + h * 2 * WG_A_SLAB + kk * 16 * LINE_BYTES, WG_A_SLAB, 1024);
                        ptx::mma_bf16(tmem_base + h * 256, adesc, bdesc, idesc, (c | kk) == 0 ? 0u : 1u);
                    }
                }
                ptx::mma_commit(empty(stage));
                if (c == total - 1) ptx::mma_commit(t_full);
            }
            __syncwarp();
            if (++stage == WG_NSTAGE) {
                stage = 0;
                phase ^= 1;
            }
        }
    } else if (warp >= 4 && total > 0) {
        // ===== epilogue: TMEM -> red.global.add into dw[co][ci][tap] =====
        const int q = warp & 3;
        ptx::mbar_wait(t_full, 0);
        ptx::tc_fence_after();
        for (int h = 0; h < halves; ++h) {
            const int co = h * 128 + 32 * q + lane;
            for (int cg = 0; cg < N / 16; ++cg) {
                uint32_t v[16];
                ptx::tmem_ld16(tmem_base + ((uint32_t)(32 * q) << 16) + h * 256 + cg * 16, v);
                ptx::tmem_ld_wait();
                if (co < P.O) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const int ci = cg * 16 + i;
                        if (ci < P.I) atomicAdd(P.dw + ((size_t)co * P.I + ci) * P.ntaps + tap, __uint_as_float(v[i]));
                    }
                }
            }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 2) ptx::tmem_dealloc(tmem_base, 512);
}

// ---- wgrad of the 3x3 convolutions, one kernel ROW per CTA ----------------------------------------------------
// k_wgrad re-streams the activations once per tap: 68 KB of operands for 8 MMAs of 128 cycles = 66 B/clk per SM
// against an L2 -> SM cap of ~42 (ncu: 867 MB per launch, tensor pipe 45 % active).  Here a CTA owns the three
// taps of one kernel row for a (128 co) x (128 ci) block: D = 3 x [128 x 128] fp32 (384 TMEM columns).  The three
// taps read the SAME dy chunk and x windows one line apart, so a stage is 16 KB of dy + a 66-line x window
// (16.5 KB) for 12 MMAs of 64 cycles: 43 B/clk -- the L2 stream and the tensor pipe now take the same time.
constexpr int WR_B_SLAB = 10240;                      // >= (7 + 66) lines, multiple of 1024
constexpr int WR_STAGE = 2 * WG_A_SLAB + 2 * WR_B_SLAB;  // 36 KB
constexpr int WR_NSTAGE = 5;
constexpr int WR_SMEM = 1024 + WR_NSTAGE * WR_STAGE;
static_assert(WR_SMEM <= 232448 && WR_STAGE % 1024 == 0, "wgrad-row shared memory");

__global__ void __launch_bounds__(256, 1) k_wgrad_row(const WgradParams P) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int warp = ptx::uniform_warp_id(), lane = threadIdx.x & 31;
    const uint32_t bar0 = ptx::smem_u32(smem);
    auto full = [&](int s) { return bar0 + 8u * s; };
    auto empty = [&](int s) { return bar0 + 8u * (8 + s); };
    const uint32_t t_full = bar0 + 8u * 16;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 256);
    const uint32_t stage0 = bar0 + 1024;
    const int co_halves = (P.co_slabs + 1) / 2, ci_halves = (P.ci_slabs + 1) / 2;
    const int kinds = 3 * co_halves * ci_halves;
    const int kind = (int)blockIdx.x % kinds, subset = (int)blockIdx.x / kinds;
    const int row = kind % 3, coh = (kind / 3) % co_halves, cih = kind / (3 * co_halves);
    const int dyk = row - 1;
    const int nco = P.co_slabs - 2 * coh < 2 ? P.co_slabs - 2 * coh : 2, nci = P.ci_slabs - 2 * cih < 2 ? P.ci_slabs - 2 * cih : 2;
    const int N = 64 * nci;
    const int my_items = P.items > subset ? (P.items - 1 - subset) / P.subsets + 1 : 0;
    constexpr int CHUNKS = SLAB_PIX / WG_KC;  // 10
    constexpr int B_BYTES = (WG_KC + 2) * LINE_BYTES;  // 66-line window: the rows of the taps dx = -1, 0, +1
    const int total = my_items * CHUNKS;

    if (threadIdx.x == 0) {
        for (int s = 0; s < WR_NSTAGE; ++s) {
            ptx::mbar_init(full(s), 1);
            ptx::mbar_init(empty(s), 1);
        }
        ptx::mbar_init(t_full, 1);
        ptx::fence_barrier_init();
    }
    if (warp == 2) {
        ptx::tmem_alloc(ptx::smem_u32(tmem_slot), 512);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== producer =====
        int stage = 0, phase = 0;
        for (int c = 0; c < total; ++c) {
            const int ii = c / CHUNKS, k0 = (c - ii * CHUNKS) * WG_KC;
            const int item = subset + ii * P.subsets;
            const int p0 = k0 + dyk * TALL_PITCH - 1, ph = p0 & 7;  // first line of the x window; two's complement handles p0 < 0
            ptx::mbar_wait(empty(stage), phase ^ 1);
            if (ptx::elect_one()) {
                const uint32_t sa = stage0 + stage * WR_STAGE, sb = sa + 2 * WG_A_SLAB;
                ptx::mbar_arrive_expect_tx(full(stage), (uint32_t)nco * WG_A_SLAB + (uint32_t)nci * B_BYTES);
                for (int s = 0; s < nco; ++s)
                    ptx::bulk_g2s(sa + s * WG_A_SLAB, P.dy + ((size_t)item * P.co_slabs + 2 * coh + s) * SLAB_BYTES + (size_t)k0 * LINE_BYTES, WG_A_SLAB,
                                  full(stage));
                for (int s = 0; s < nci; ++s)
                    ptx::bulk_g2s(sb + s * WR_B_SLAB + ph * LINE_BYTES,
                                  P.x + ((ptrdiff_t)((size_t)item * P.ci_slabs + 2 * cih + s) * SLAB_BYTES + (ptrdiff_t)p0 * LINE_BYTES), B_BYTES, full(stage));
            }
            __syncwarp();
            if (++stage == WR_NSTAGE) {
                stage = 0;
                phase ^= 1;
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: D[tap][128 co x N ci], A and B MN-major =====
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
        int stage = 0, phase = 0;
        for (int c = 0; c < total; ++c) {
            const int ii = c / CHUNKS, k0 = (c - ii * CHUNKS) * WG_KC;
            const int ph = (k0 + dyk * TALL_PITCH - 1) & 7;
            ptx::mbar_wait(full(stage), phase);
            ptx::tc_fence_after();
            if (ptx::elect_one()) {
                const uint32_t sa = stage0 + stage * WR_STAGE, sb = sa + 2 * WG_A_SLAB + ph * LINE_BYTES;
#pragma unroll
                for (int kk = 0; kk < WG_KC / 16; ++kk) {
                    const uint64_t adesc = wg_desc(sa + kk * 16 * LINE_BYTES, WG_A_SLAB, 1024);
#pragma unroll
                    for (int tp = 0; tp < 3; ++tp)  // tap dx = tp - 1 starts tp lines into the window
                        ptx::mma_bf16(tmem_base + tp * 128, adesc, wg_desc(sb + (tp + kk * 16) * LINE_BYTES, WR_B_SLAB, 1024), idesc,
                                      (c | kk) == 0 ? 0u : 1u);
                }
                ptx::mma_commit(empty(stage));
                if (c == total - 1) ptx::mma_commit(t_full);
            }
            __syncwarp();
            if (++stage == WR_NSTAGE) {
                stage = 0;
                phase ^= 1;
            }
        }
    } else if (warp >= 4 && total > 0) {
        // ===== epilogue: TMEM -> red.global.add into dw[co][ci][tap] =====
        const int q = warp & 3;
        ptx::mbar_wait(t_full, 0);
        ptx::tc_fence_after();
        const int co = coh * 128 + 32 * q + lane;
        const int COP = 64 * P.co_slabs, CIP = 64 * P.ci_slabs;  // padded extents of the partial buffer
        for (int tp = 0; tp < 3; ++tp) {
            const int tap = row * 3 + tp;
            for (int cg = 0; cg < N / 16; ++cg) {
                uint32_t v[16];
                ptx::tmem_ld16(tmem_base + ((uint32_t)(32 * q) << 16) + tp * 128 + cg * 16, v);
                ptx::tmem_ld_wait();
                if (32 * q + lane < 64 * nco) {
                    if (P.part) {  // plain stores, consecutive lanes = consecutive co: one 128-byte line per warp store
                        float* dst = P.part + (((size_t)subset * 9 + tap) * CIP + (cih * 128 + cg * 16)) * COP + co;
#pragma unroll
                        for (int i = 0; i < 16; ++i) dst[(size_t)i * COP] = __uint_as_float(v[i]);
                    } else if (co < P.O) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const int ci = cih * 128 + cg * 16 + i;
                            if (ci < P.I) atomicAdd(P.dw + ((size_t)co * P.I + ci) * 9 + tap, __uint_as_float(v[i]));
                        }
                    }
                }
            }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 2) ptx::tmem_dealloc(tmem_base, 512);
}

// dw[co][ci][tap] = sum over subsets of part[subset][tap][ci][co]  (reads coalesced along co)
__global__ void __launch_bounds__(256) k_wgrad_reduce(const float* part, int subsets, int COP, int CIP, int O, int I, float* dw) {
    const size_t per = (size_t)9 * CIP * COP;
    for (size_t e = (size_t)blockIdx.x * 256 + threadIdx.x; e < per; e += (size_t)gridDim.x * 256) {
        const int co = (int)(e % COP);
        const int ci = (int)((e / COP) % CIP);
        const int tap = (int)(e / ((size_t)COP * CIP));
        if (co >= O || ci >= I) continue;
        float s = 0.0f;
        for (int k = 0; k < subsets; ++k) s += part[(size_t)k * per + e];
        dw[((size_t)co * I + ci) * 9 + tap] = s;
    }
}

// ---- policy head: softmax, loss and dlogits (nn.cpp:80, 98-102) -----------------------------------------
// logits [boards][64][73] fp32 (index 73*sq + t == the action code).  Writes dlogits as bf16 into a
// 2-slab tall tensor (channels 0..72; 73..127 stay zero), accumulates the policy loss and dbias.
__global__ void __launch_bounds__(256) k_policy_loss(const float* logits, const float* obs_p, int boards, uint16_t* dlogits, float* loss, float* dbias) {
    __shared__ float red[8];
    __shared__ float bc;
    __shared__ float sdb[73];
    const int b = blockIdx.x, t = threadIdx.x;
    if (b >= boards) return;
    const float* row = logits + (size_t)b * KB_PSIZE;
    const float* tgt = obs_p + (size_t)b * KB_PSIZE;
    constexpr int PER = (KB_PSIZE + 255) / 256;
    if (t < 73) sdb[t] = 0.0f;
    float v[PER];
    float m = -INFINITY;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
        const int idx = t + 256 * i;
        v[i] = idx < KB_PSIZE ? row[idx] : -INFINITY;
        m = fmaxf(m, v[i]);
    }
    auto block_max = [&](float x) {
        for (int off = 16; off; off >>= 1) x = fmaxf(x, __shfl_xor_sync(0xffffffffu, x, off));
        __syncthreads();
        if ((t & 31) == 0) red[t >> 5] = x;
        __syncthreads();
        if (t == 0) {
            float r = red[0];
            for (int i = 1; i < 8; ++i) r = fmaxf(r, red[i]);
            bc = r;
        }
        __syncthreads();
        return bc;
    };
    auto block_sum = [&](float x) {
        for (int off = 16; off; off >>= 1) x += __shfl_xor_sync(0xffffffffu, x, off);
        __syncthreads();
        if ((t & 31) == 0) red[t >> 5] = x;
        __syncthreads();
        if (t == 0) {
            float r = 0.0f;
            for (int i = 0; i < 8; ++i) r += red[i];
            bc = r;
        }
        __syncthreads();
        return bc;
    };
    m = block_max(m);
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
        v[i] = (t + 256 * i) < KB_PSIZE ? expf(v[i] - m) : 0.0f;
        s += v[i];
    }
    const float inv = 1.0f / block_sum(s);
    float g[PER];
    float lp = 0.0f, S = 0.0f;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
        const int idx = t + 256 * i;
        g[i] = 0.0f;
        if (idx < KB_PSIZE) {
            v[i] *= inv;  // p
            const float o = tgt[idx];
            if (o != 0.0f) {
                lp -= o * logf(v[i] + 0.001f);
                g[i] = o / (v[i] + 0.001f);
                S = fmaf(g[i], v[i], S);
            }
        }
    }
    S = block_sum(S);
    lp = block_sum(lp);
    const int item = b / NB, slot = b - item * NB;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
        const int idx = t + 256 * i;
        if (idx < KB_PSIZE) {
            const float d = v[i] * (S - g[i]);  // dL/dz_k = p_k * (sum_a g_a p_a - g_k)
            const int sq = idx / 73, k = idx - sq * 73;
            const int px = tall_pixel(slot, sq);
            const size_t u4 = ((size_t)item * 2 + (k >> 6)) * SLAB_U4 + chunk_u4(px, (k >> 3) & 7);
            const __nv_bfloat16 h = __float2bfloat16_rn(d);
            dlogits[u4 * 8 + (k & 7)] = *reinterpret_cast<const uint16_t*>(&h);
            atomicAdd(&sdb[k], d);
        }
    }
    __syncthreads();
    if (t < 73) atomicAdd(dbias + t, sdb[t]);
    if (t == 0) atomicAdd(loss, lp);
}

// ---- value head (nn.cpp:83-88), fp32 ------------------------------------------------------------------
// vpre[b][sq] = valueconv(x) ; accumulates sum / sumsq for the single-channel BatchNorm
__global__ void __launch_bounds__(256) k_value_conv(const uint4* x, int slabs, int boards, const float* wv, const float* bv, float* vpre, float* sums) {
    __shared__ float r0[8], r1[8];
    const int g = blockIdx.x * 256 + threadIdx.x;
    float v = 0.0f;
    const bool ok = g < boards * 64;
    if (ok) {
        v = bv[0];
        for (int c = 0; c < slabs * 8; ++c) {
            float f[8];
            unpack8(x[act_idx(g >> 6, g & 63, slabs, c >> 3, c & 7)], f);
#pragma unroll
            for (int k = 0; k < 8; ++k) v = fmaf(f[k], wv[c * 8 + k], v);
        }
        vpre[g] = v;
    }
    float a = ok ? v : 0.0f, q = ok ? v * v : 0.0f;
    for (int off = 16; off; off >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, off);
        q += __shfl_xor_sync(0xffffffffu, q, off);
    }
    if ((threadIdx.x & 31) == 0) {
        r0[threadIdx.x >> 5] = a;
        r1[threadIdx.x >> 5] = q;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < 8; ++i) {
            a += r0[i];
            q += r1[i];
        }
        atomicAdd(sums, a);
        atomicAdd(sums + 1, q);
    }
}
// per board: vact = relu(bn(vpre)); v = tanh(fc(vact)); value loss and d(loss)/d(fc pre-activation)
__global__ void __launch_bounds__(256) k_value_fc(const float* vpre, int boards, const float* stats /*mean,rstd*/, const float* gamma, const float* beta,
                                                  const float* fcw /*[256][64]*/, const float* fcb, const float* obs_v, float* vact, float* dfc, float* loss) {
    __shared__ float a[64];
    __shared__ float red[8];
    const int b = blockIdx.x, t = threadIdx.x;
    if (t < 64) {
        const float xh = (vpre[b * 64 + t] - stats[0]) * stats[1];
        const float y = fmaxf(fmaf(gamma[0], xh, beta[0]), 0.0f);
        a[t] = y;
        vact[b * 64 + t] = y;
    }
    __syncthreads();
    float o = fcb[t];
#pragma unroll 8
    for (int p = 0; p < 64; ++p) o = fmaf(a[p], fcw[t * 64 + p], o);
    const float v = tanhf(o);
    const float e = v - obs_v[b];
    const float scale = 1.0f / (256.0f * (float)boards);  // mse_loss over the broadcast [B,256] (nn.cpp:96)
    dfc[b * 256 + t] = 2.0f * e * scale * (1.0f - v * v);
    float l = e * e * scale;
    for (int off = 16; off; off >>= 1) l += __shfl_xor_sync(0xffffffffu, l, off);
    if ((t & 31) == 0) red[t >> 5] = l;
    __syncthreads();
    if (t == 0) {
        for (int i = 1; i < 8; ++i) l += red[i];
        atomicAdd(loss, l);
    }
}
// dW_fc[t][p] = sum_b dfc[b][t] * vact[b][p]; db_fc[t] = sum_b dfc[b][t]
__global__ void __launch_bounds__(256) k_value_fc_wgrad(const float* dfc, const float* vact, int boards, float* dw, float* db) {
    const int e = blockIdx.x * 256 + threadIdx.x;  // 256*64 weights + 256 biases
    if (e < 256 * 64) {
        const int t = e >> 6, p = e & 63;
        float s = 0.0f;
        for (int b = 0; b < boards; ++b) s = fmaf(dfc[b * 256 + t], vact[b * 64 + p], s);
        dw[e] = s;
    } else if (e < 256 * 64 + 256) {
        const int t = e - 256 * 64;
        float s = 0.0f;
        for (int b = 0; b < boards; ++b) s += dfc[b * 256 + t];
        db[t] = s;
    }
}
// dvn[b][p] = (sum_t dfc[b][t] * fcw[t][p]) * (vact > 0); accumulates the BatchNorm(1) backward sums
__global__ void __launch_bounds__(64) k_value_fc_dgrad(const float* dfc, const float* fcw, const float* vact, const float* vpre, const float* stats, float* dvn,
                                                        float* s12) {
    const int b = blockIdx.x, p = threadIdx.x;
    float s = 0.0f;
    for (int t = 0; t < 256; ++t) s = fmaf(dfc[b * 256 + t], fcw[t * 64 + p], s);
    const float d = vact[b * 64 + p] > 0.0f ? s : 0.0f;
    dvn[b * 64 + p] = d;
    const float xh = (vpre[b * 64 + p] - stats[0]) * stats[1];
    float a = d, q = d * xh;
    for (int off = 16; off; off >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, off);
        q += __shfl_xor_sync(0xffffffffu, q, off);
    }
    if ((p & 31) == 0) {
        atomicAdd(s12, a);
        atomicAdd(s12 + 1, q);
    }
}
__global__ void k_value_bn_finalize(float* sums, float n, float* stats, float* run_mean, float* run_var) {
    const float m = sums[0] / n, var = fmaxf(sums[1] / n - m * m, 0.0f);
    stats[0] = m;
    stats[1] = rsqrtf(var + BN_EPS_T);
    run_mean[0] = (1.0f - BN_MOMENTUM_T) * run_mean[0] + BN_MOMENTUM_T * m;
    run_var[0] = (1.0f - BN_MOMENTUM_T) * run_var[0] + BN_MOMENTUM_T * var * (n / fmaxf(n - 1.0f, 1.0f));
    sums[0] = 0.0f;
    sums[1] = 0.0f;
}
// dvpre = gamma*rstd*(dvn - s1/n - xhat*s2/n); dX[b,p,c] += dvpre * wv[c]; d wv[c] += sum dvpre * x[c] (same thread layout as the BN kernels)
__global__ void __launch_bounds__(256) k_value_conv_bwd(const float* dvn, const float* vpre, const float* stats, const float* gamma, const float* s12, float n,
                                                        const uint4* x, uint4* dx, int boards, int slabs, const float* wv, float* dwv, float* dbv_dummy) {
    const int s = blockIdx.x, j = threadIdx.x & 7, pl = threadIdx.x >> 3;
    const int c0 = s * 64 + j * 8;
    float w[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) w[k] = wv[c0 + k];
    const float m1 = s12[0] / n, m2 = s12[1] / n, gr = gamma[0] * stats[1];
    float a[8] = {0, 0, 0, 0, 0, 0, 0, 0}, q[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const int npix = boards * 64;
    for (int g = blockIdx.y * 32 + pl; g < npix; g += gridDim.y * 32) {
        const float xh = (vpre[g] - stats[0]) * stats[1];
        const float dv = gr * (dvn[g] - m1 - xh * m2);
        const size_t i = act_idx(g >> 6, g & 63, slabs, s, j);
        float f[8], d[8];
        unpack8(x[i], f);
        unpack8(dx[i], d);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            a[k] = fmaf(dv, f[k], a[k]);
            d[k] = fmaf(dv, w[k], d[k]);
        }
        dx[i] = pack8(d);
    }
    block_channel_sums(a, q, dwv, dbv_dummy, s * 64);
}
// vbatchnorm dgamma / dbeta from the accumulated sums; clears them
__global__ void k_value_bn_bwd_finalize(float* s12, float* dgamma, float* dbeta) {
    dgamma[0] = s12[1];
    dbeta[0] = s12[0];
    s12[0] = 0.0f;
    s12[1] = 0.0f;
}

__global__ void k_sgd(float* params, const float* grads, size_t n, float lr_scale) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) params[i] -= lr_scale * grads[i];
}

}  // namespace kb

// =============================================================================================================
// host side of the trainer
// =============================================================================================================
namespace {

struct TConv {            // one convolution of the network and everything the step needs for it
    int O, I, k;          // reference weight shape [O][I][k][k]
    size_t w_off, b_off;  // offsets into the flat parameter / gradient vectors
    size_t g_off = 0, be_off = 0, rm_off = 0, rv_off = 0;  // BatchNorm (0 when the conv has none)
    bool bn = false;
    Layer fwd, dgrad;     // packed bf16 operands (trainer-owned)
    bool has_dgrad = false;
    uint4 *pre = nullptr, *post = nullptr;  // saved activations (bf16 tall layout)
    float *mean = nullptr, *rstd = nullptr; // batch statistics [O]
    float *accum = nullptr;                 // this layer's [4][O] sums: forward sum / sumsq, backward s1 / s2 (zeroed once per step)
    int out_slabs = 0, in_slabs = 0;
};

}  // namespace

struct kb_trainer {
    int device = 0;
    cudaStream_t last_stream = nullptr;
    int F = 0, R = 0, cap = 0, batch = 0;
    size_t n_floats = 0;
    float *params = nullptr, *grads = nullptr;
    std::vector<TConv> conv;  // conv1, (res conv1, res conv2)*, policyconv, policyconv2
    size_t vw_off = 0, vb_off = 0, vg_off = 0, vbe_off = 0, vrm_off = 0, vrv_off = 0, fcw_off = 0, fcb_off = 0;
    // buffers (bf16 tall layout, each with guard bytes in front and behind: k_wgrad reads up to 11 lines outside a slab)
    std::vector<void*> allocs;
    uint4 *P = nullptr, *dlogits = nullptr, *dH = nullptr, *dXa = nullptr, *dXb = nullptr, *dpre = nullptr, *dpre2 = nullptr;
    float *logits = nullptr, *obs_dev = nullptr, *pi_dev = nullptr, *z_dev = nullptr;
    float *vpre = nullptr, *vact = nullptr, *dfc = nullptr, *dvn = nullptr, *vstats = nullptr;
    float *acc = nullptr;     // [2][256] scratch + [4] value-head scalars + [1] loss
    float *bn_arena = nullptr; // every BatchNorm layer's [4][O] accumulators, one memset per step
    size_t bn_arena_floats = 0;
    float *zeros = nullptr;   // 256 zero biases for the dgrad convs
    float *wpart = nullptr;   // k_wgrad_row partial sums: [subsets <= 49][9][<= 256][<= 256]
    size_t wpart_floats = 0;
    float last_loss = 0.0f;
    void* pack_jobs = nullptr;  // device table of k_pack_all
    int n_pack_jobs = 0, pack_blocks = 0;
};

namespace {

constexpr size_t T_GUARD = 4096;

int t_alloc(kb_trainer* t, void** out, size_t bytes, bool guard) {
    uint8_t* p = nullptr;
    const size_t total = bytes + (guard ? 2 * T_GUARD : 0);
    KB_CUDA(cudaMalloc(&p, total));
    KB_CUDA(cudaMemsetAsync(p, 0, total, main_stream()));
    t->allocs.push_back(p);
    *out = p + (guard ? T_GUARD : 0);
    return KB_OK;
}
int t_alloc_act(kb_trainer* t, uint4** out, int slabs) {
    void* p;
    int r = t_alloc(t, &p, act_bytes(t->cap, slabs), true);
    *out = reinterpret_cast<uint4*>(p);
    return r;
}

int t_layer(kb_trainer* t, Layer& L, int slabs_in, int ksteps, int n_total, int n_valid, int ntaps, const float* bias) {
    L.slabs_in = slabs_in;
    L.ksteps = ksteps;
    L.n_total = n_total;
    L.n_tile = n_total == 80 ? 80 : (n_total % 128 == 0 ? 128 : 64);
    L.n_valid = n_valid;
    L.ntaps = ntaps;
    L.relu = 0;
    L.bias = const_cast<float*>(bias);
    const size_t bytes = (size_t)(n_total / L.n_tile) * slabs_in * ntaps * L.n_tile * 128;
    void* p;
    int r = t_alloc(t, &p, bytes, false);
    L.w = reinterpret_cast<uint4*>(p);
    return r;
}

int t_repack(kb_trainer* t, cudaStream_t st) {
    if (!t->pack_jobs) {  // the job table is fixed for the life of the trainer
        std::vector<PackJob> jobs;
        int block0 = 0;
        auto add = [&](const TConv& c, const Layer& L, bool flip) {
            PackJob J;
            J.w = t->params + c.w_off;
            J.dst = reinterpret_cast<uint16_t*>(L.w);
            J.O = c.O;
            J.I = c.I;
            J.taps = c.k * c.k;
            J.transpose_flip = flip ? 1 : 0;
            J.n_tile = L.n_tile;
            J.kslices = L.slabs_in;
            J.npass = L.n_total / L.n_tile;
            const size_t total = (size_t)J.npass * J.kslices * J.taps * J.n_tile * 64;
            J.nblocks = (int)((total + 2047) / 2048 < 288 ? (total + 2047) / 2048 : 288);  // >= 8 elements per thread
            if (J.nblocks < 1) J.nblocks = 1;
            J.block0 = block0;
            block0 += J.nblocks;
            jobs.push_back(J);
        };
        for (auto& c : t->conv) {
            add(c, c.fwd, false);
            if (c.has_dgrad) add(c, c.dgrad, true);
        }
        void* d = nullptr;
        int r = t_alloc(t, &d, sizeof(PackJob) * jobs.size(), false);
        if (r) return r;
        KB_CUDA(cudaMemcpyAsync(d, jobs.data(), sizeof(PackJob) * jobs.size(), cudaMemcpyHostToDevice, st));
        KB_CUDA(cudaStreamSynchronize(st));  // jobs is a local
        t->pack_jobs = d;
        t->n_pack_jobs = (int)jobs.size();
        t->pack_blocks = block0;
    }
    k_pack_all<<<t->pack_blocks, 256, 0, st>>>(reinterpret_cast<const PackJob*>(t->pack_jobs), t->n_pack_jobs);
    KB_CUDA(cudaGetLastError());
    return KB_OK;
}

// Two resident blocks per SM in total, each thread streaming BN_U pixel lines per trip: measured on the 20x256 step,
// 592 x slabs one-trip blocks 13.3 ms, 74 x 4 blocks 12.3 ms (prologue parameter loads and block turnover dominate short
// blocks); fewer than two blocks per SM loses again.
dim3 bn_grid(int slabs, int boards) {
    int y = (boards * 64 + 32 * BN_U - 1) / (32 * BN_U);
    static int cap_blocks = 0;
    if (!cap_blocks) {
        const char* e = getenv("KB_BN_BLOCKS");
        cap_blocks = e ? atoi(e) : 2 * sm_count();
        if (cap_blocks < 1) cap_blocks = 2 * sm_count();
    }
    const int cap = cap_blocks / (slabs > 0 ? slabs : 1) > 0 ? cap_blocks / (slabs > 0 ? slabs : 1) : 1;
    if (y > cap) y = cap;
    return dim3(slabs, y > 0 ? y : 1);
}

int t_wgrad(kb_trainer* t, const TConv& c, const uint4* dy, const uint4* x, int x_slabs, cudaStream_t st) {
    static bool configured[16] = {};  // function attributes are per device
    if (!configured[current_device() & 15]) {
        KB_CUDA(cudaFuncSetAttribute(k_wgrad, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM));
        configured[current_device() & 15] = true;
    }
    WgradParams p;
    p.dy = reinterpret_cast<const uint8_t*>(dy);
    p.x = reinterpret_cast<const uint8_t*>(x);
    p.dw = t->grads + c.w_off;
    p.items = items_for(t->batch);
    p.co_slabs = c.out_slabs;
    p.ci_slabs = x_slabs;
    p.O = c.O;
    p.I = c.I;
    p.ntaps = c.k * c.k;
    p.part = nullptr;
    static int use_row = -1;
    if (use_row < 0) {
        const char* e = getenv("KB_WGRAD_PER_TAP");
        use_row = (e && e[0] == '1') ? 0 : 1;
    }
    if (p.ntaps == 9 && use_row) {  // one kernel row (three taps) x 128 co x 128 ci per CTA
        static bool configured_row[16] = {};
        if (!configured_row[current_device() & 15]) {
            KB_CUDA(cudaFuncSetAttribute(k_wgrad_row, cudaFuncAttributeMaxDynamicSharedMemorySize, WR_SMEM));
            configured_row[current_device() & 15] = true;
        }
        const int kinds = 3 * ((p.co_slabs + 1) / 2) * ((p.ci_slabs + 1) / 2);
        int subsets = sm_count() / kinds;
        if (subsets > p.items) subsets = p.items;
        if (subsets < 1) subsets = 1;
        p.subsets = subsets;
        const int COP = 64 * p.co_slabs, CIP = 64 * p.ci_slabs;
        const size_t need = (size_t)subsets * 9 * CIP * COP;
        p.part = need <= t->wpart_floats ? t->wpart : nullptr;  // too many subsets for the scratch: fall back to atomics
        k_wgrad_row<<<kinds * subsets, 256, WR_SMEM, st>>>(p);
        KB_CUDA(cudaGetLastError());
        if (p.part) {
            k_wgrad_reduce<<<592, 256, 0, st>>>(p.part, subsets, COP, CIP, p.O, p.I, p.dw);
            KB_CUDA(cudaGetLastError());
        }
        return KB_OK;
    }
    int subsets = sm_count() / p.ntaps;
    if (subsets > p.items) subsets = p.items;
    if (subsets < 1) subsets = 1;
    p.subsets = subsets;
    k_wgrad<<<p.ntaps * subsets, 256, WG_SMEM, st>>>(p);
    KB_CUDA(cudaGetLastError());
    return KB_OK;
}

// BatchNorm + ReLU backward of conv c: dy (w.r.t. the ReLU output) -> dpre_out; gamma/beta gradients
int t_bn_bwd(kb_trainer* t, const TConv& c, const uint4* dy, uint4* dpre_out, cudaStream_t st) {
    const float n = (float)t->batch * 64.0f;
    float *s1 = c.accum + 2 * c.O, *s2 = c.accum + 3 * c.O;
    const float *g = t->params + c.g_off, *be = t->params + c.be_off;
    k_bn_bwd_reduce<<<bn_grid(c.out_slabs, t->batch), 256, 0, st>>>(dy, c.pre, t->batch, c.out_slabs, c.mean, c.rstd, g, be, s1, s2);
    k_bn_bwd_apply<<<bn_grid(c.out_slabs, t->batch), 256, 0, st>>>(dy, c.pre, dpre_out, t->batch, c.out_slabs, c.mean, c.rstd, g, be, s1, s2, n,
                                                                   t->grads + c.g_off, t->grads + c.be_off);
    KB_CUDA(cudaGetLastError());
    return KB_OK;
}

// conv forward in training mode: in -> pre (conv + bias), batch statistics, post = relu(bn(pre)) (+ skip)
int t_conv_bn_fwd(kb_trainer* t, TConv& c, const uint4* in, const uint4* skip, cudaStream_t st) {
    int r = run_conv(c.fwd, in, c.pre, nullptr, nullptr, t->batch, st);
    if (r) return r;
    const float n = (float)t->batch * 64.0f;
    float *s1 = c.accum, *s2 = c.accum + c.O;
    k_bn_stats<<<bn_grid(c.out_slabs, t->batch), 256, 0, st>>>(c.pre, t->batch, c.out_slabs, s1, s2);
    k_bn_apply<<<bn_grid(c.out_slabs, t->batch), 256, 0, st>>>(c.pre, c.post, skip, t->batch, c.out_slabs, s1, s2, n, c.mean, c.rstd,
                                                               t->params + c.rm_off, t->params + c.rv_off, t->params + c.g_off, t->params + c.be_off);
    KB_CUDA(cudaGetLastError());
    return KB_OK;
}

int t_forward_backward(kb_trainer* t, const float* obs_dev, const float* pi_dev, const float* z_dev, int batch, cudaStream_t st) {
    if (batch != t->batch) {  // idle boards of the last item and pad pixels must read as zero
        for (uint4* b : {t->P, t->dlogits, t->dH, t->dXa, t->dXb, t->dpre, t->dpre2})
            KB_CUDA(cudaMemsetAsync(b, 0, act_bytes(t->cap, b == t->P ? 1 : (b == t->dlogits || b == t->dH ? 2 : t->F / 64)), st));
        for (auto& c : t->conv)
            if (c.pre) {
                KB_CUDA(cudaMemsetAsync(c.pre, 0, act_bytes(t->cap, c.out_slabs), st));
                KB_CUDA(cudaMemsetAsync(c.post, 0, act_bytes(t->cap, c.out_slabs), st));
            }
        t->batch = batch;
    }
    const int R = t->R, fs = t->F / 64;
    const float n = (float)batch * 64.0f;
    int r;
    KB_CUDA(cudaMemsetAsync(t->grads, 0, sizeof(float) * t->n_floats, st));
    float* loss = t->acc + 516;
    float* vsum = t->acc + 512;
    KB_CUDA(cudaMemsetAsync(t->acc, 0, sizeof(float) * 520, st));
    KB_CUDA(cudaMemsetAsync(t->bn_arena, 0, sizeof(float) * t->bn_arena_floats, st));
    // ---------------- forward ----------------
    if ((r = obs_to_tall_launch(obs_dev, batch, t->P, st))) return r;
    if ((r = t_conv_bn_fwd(t, t->conv[0], t->P, nullptr, st))) return r;
    const uint4* x = t->conv[0].post;
    for (int i = 0; i < R; ++i) {
        TConv &c1 = t->conv[1 + 2 * i], &c2 = t->conv[2 + 2 * i];
        if ((r = t_conv_bn_fwd(t, c1, x, nullptr, st))) return r;
        if ((r = t_conv_bn_fwd(t, c2, c1.post, x, st))) return r;  // x = skip + relu(bn2(conv2(.)))
        x = c2.post;
    }
    TConv &pc = t->conv[1 + 2 * R], &pc2 = t->conv[2 + 2 * R];
    if ((r = t_conv_bn_fwd(t, pc, x, nullptr, st))) return r;
    if ((r = run_conv(pc2.fwd, pc.post, nullptr, nullptr, t->logits, batch, st))) return r;
    k_policy_loss<<<batch, 256, 0, st>>>(t->logits, pi_dev, batch, reinterpret_cast<uint16_t*>(t->dlogits), loss, t->grads + pc2.b_off);
    // value head
    const float* P_ = t->params;
    k_value_conv<<<(batch * 64 + 255) / 256, 256, 0, st>>>(x, fs, batch, P_ + t->vw_off, P_ + t->vb_off, t->vpre, vsum);
    k_value_bn_finalize<<<1, 1, 0, st>>>(vsum, n, t->vstats, t->params + t->vrm_off, t->params + t->vrv_off);
    k_value_fc<<<batch, 256, 0, st>>>(t->vpre, batch, t->vstats, P_ + t->vg_off, P_ + t->vbe_off, P_ + t->fcw_off, P_ + t->fcb_off, z_dev, t->vact, t->dfc, loss);
    KB_CUDA(cudaGetLastError());
    // ---------------- backward: heads ----------------
    k_value_fc_wgrad<<<(256 * 64 + 256 + 255) / 256, 256, 0, st>>>(t->dfc, t->vact, batch, t->grads + t->fcw_off, t->grads + t->fcb_off);
    k_value_fc_dgrad<<<batch, 64, 0, st>>>(t->dfc, P_ + t->fcw_off, t->vact, t->vpre, t->vstats, t->dvn, vsum + 2);
    // policy: dlogits -> dH (dgrad of policyconv2), dW2; BN backward; dX (dgrad of policyconv), dW1
    if ((r = t_wgrad(t, pc2, t->dlogits, pc.post, 2, st))) return r;
    if ((r = run_conv(pc2.dgrad, t->dlogits, t->dH, nullptr, nullptr, batch, st))) return r;
    if ((r = t_bn_bwd(t, pc, t->dH, t->dH, st))) return r;  // in place: dH becomes d(policyconv pre-activation)
    if ((r = t_wgrad(t, pc, t->dH, x, fs, st))) return r;
    uint4 *dcur = t->dXa, *dnext = t->dXb;
    if ((r = run_conv(pc.dgrad, t->dH, dcur, nullptr, nullptr, batch, st))) return r;
    // value: dX += dvpre * wv, d wv, vbatchnorm gradients
    k_value_conv_bwd<<<bn_grid(fs, batch), 256, 0, st>>>(t->dvn, t->vpre, t->vstats, P_ + t->vg_off, vsum + 2, n, x, dcur, batch, fs, P_ + t->vw_off,
                                                         t->grads + t->vw_off, t->acc + 256);
    k_value_bn_bwd_finalize<<<1, 1, 0, st>>>(vsum + 2, t->grads + t->vg_off, t->grads + t->vbe_off);
    KB_CUDA(cudaMemsetAsync(t->acc + 256, 0, sizeof(float) * 256, st));  // scratch second output of k_value_conv_bwd
    KB_CUDA(cudaGetLastError());
    // ---------------- backward: tower ----------------
    for (int i = R - 1; i >= 0; --i) {
        TConv &c1 = t->conv[1 + 2 * i], &c2 = t->conv[2 + 2 * i];
        const uint4* xin = i == 0 ? t->conv[0].post : t->conv[2 * i].post;  // block input (skip)
        if ((r = t_bn_bwd(t, c2, dcur, t->dpre, st))) return r;
        if ((r = t_wgrad(t, c2, t->dpre, c1.post, fs, st))) return r;
        if ((r = run_conv(c2.dgrad, t->dpre, t->dpre2, nullptr, nullptr, batch, st))) return r;  // d(relu(bn1(.)))
        if ((r = t_bn_bwd(t, c1, t->dpre2, t->dpre, st))) return r;
        if ((r = t_wgrad(t, c1, t->dpre, xin, fs, st))) return r;
        if ((r = run_conv(c1.dgrad, t->dpre, dnext, dcur, nullptr, batch, st))) return r;       // + skip path
        uint4* tmp = dcur;
        dcur = dnext;
        dnext = tmp;
    }
    if ((r = t_bn_bwd(t, t->conv[0], dcur, t->dpre, st))) return r;
    if ((r = t_wgrad(t, t->conv[0], t->dpre, t->P, 1, st))) return r;
    KB_CUDA(cudaMemcpyAsync(&t->last_loss, loss, sizeof(float), cudaMemcpyDeviceToHost, st));
    return KB_OK;
}

}  // namespace

extern "C" {

int kb_trainer_create(kb_trainer** out, int filters, int residuals, int max_batch) {
    KB_REQUIRE_INIT();
    KB_ARG(out && (filters == 64 || filters == 128 || filters == 256) && residuals >= 0 && residuals <= 64 && max_batch > 0,
           "out / filters in {64, 128, 256} / residuals / max_batch");
    kb_trainer* t = new (std::nothrow) kb_trainer();
    if (!t) return KB_ERR_ARG;
    t->device = current_device();
    t->last_stream = main_stream();
    t->F = filters;
    t->R = residuals;
    t->cap = max_batch;
    t->n_floats = kb_net_blob_floats(filters, residuals);
    const int F = filters, fs = F / 64;
    int r;
    void* p;
#define T_TRY(x) do { if ((r = (x))) { kb_trainer_destroy(t); return r; } } while (0)
    T_TRY(t_alloc(t, &p, sizeof(float) * t->n_floats, false));
    t->params = (float*)p;
    T_TRY(t_alloc(t, &p, sizeof(float) * t->n_floats, false));
    t->grads = (float*)p;
    T_TRY(t_alloc(t, &p, sizeof(float) * 256, false));
    t->zeros = (float*)p;
    T_TRY(t_alloc(t, &p, sizeof(float) * 520, false));
    t->acc = (float*)p;
    {
        const int halves = (filters / 64 + 1) / 2, kinds = 3 * halves * halves;
        int subsets = sm_count() / kinds;
        if (subsets < 1) subsets = 1;
        t->wpart_floats = (size_t)subsets * 9 * filters * filters;
        T_TRY(t_alloc(t, &p, sizeof(float) * t->wpart_floats, false));
        t->wpart = (float*)p;
    }
    size_t off = 0;
    auto add_conv = [&](int O, int I, int k, bool bn, int in_slabs, int ksteps, bool dgrad, bool save) -> int {
        TConv c;
        c.O = O;
        c.I = I;
        c.k = k;
        c.w_off = off;
        off += (size_t)O * I * k * k;
        c.b_off = off;
        off += O;
        c.bn = bn;
        if (bn) {
            c.g_off = off;
            c.be_off = off + O;
            c.rm_off = off + 2 * O;
            c.rv_off = off + 3 * O;
            off += 4 * (size_t)O;
        }
        c.in_slabs = in_slabs;
        c.out_slabs = (O + 63) / 64;
        const int n_total = O == 73 ? 80 : O;
        int rr = t_layer(t, c.fwd, in_slabs, ksteps, n_total, O, k * k, t->params + c.b_off);
        if (rr) return rr;
        c.has_dgrad = dgrad;
        if (dgrad) {  // logical conv: c.out_slabs*64 input channels -> I output channels
            if ((rr = t_layer(t, c.dgrad, c.out_slabs, 4, I, I, k * k, t->zeros))) return rr;
        }
        if (save) {
            if ((rr = t_alloc_act(t, &c.pre, c.out_slabs)) || (rr = t_alloc_act(t, &c.post, c.out_slabs))) return rr;
            void* q;
            if ((rr = t_alloc(t, &q, sizeof(float) * O, false))) return rr;
            c.mean = (float*)q;
            if ((rr = t_alloc(t, &q, sizeof(float) * O, false))) return rr;
            c.rstd = (float*)q;
        }
        t->conv.push_back(c);
        return KB_OK;
    };
    T_TRY(add_conv(F, 30, 3, true, 1, 2, false, true));
    for (int i = 0; i < residuals; ++i) {
        T_TRY(add_conv(F, F, 3, true, fs, 4, true, true));
        T_TRY(add_conv(F, F, 3, true, fs, 4, true, true));
    }
    T_TRY(add_conv(128, F, 1, true, fs, 4, true, true));
    T_TRY(add_conv(73, 128, 1, false, 2, 4, true, false));
    for (auto& c : t->conv)
        if (c.bn) t->bn_arena_floats += 4 * (size_t)c.O;
    T_TRY(t_alloc(t, &p, sizeof(float) * t->bn_arena_floats, false));
    t->bn_arena = (float*)p;
    {
        size_t o = 0;
        for (auto& c : t->conv)
            if (c.bn) {
                c.accum = t->bn_arena + o;
                o += 4 * (size_t)c.O;
            }
    }
    t->vw_off = off;
    off += F;
    t->vb_off = off;
    off += 1;
    t->vg_off = off;
    t->vbe_off = off + 1;
    t->vrm_off = off + 2;
    t->vrv_off = off + 3;
    off += 4;
    t->fcw_off = off;
    off += 256 * 64;
    t->fcb_off = off;
    off += 256;
    if (off != t->n_floats) {
        set_error("trainer parameter layout mismatch (%zu vs %zu)", off, t->n_floats);
        kb_trainer_destroy(t);
        return KB_ERR_STATE;
    }
    T_TRY(t_alloc_act(t, &t->P, 1));
    T_TRY(t_alloc_act(t, &t->dlogits, 2));
    T_TRY(t_alloc_act(t, &t->dH, 2));
    T_TRY(t_alloc_act(t, &t->dXa, fs));
    T_TRY(t_alloc_act(t, &t->dXb, fs));
    T_TRY(t_alloc_act(t, &t->dpre, fs));
    T_TRY(t_alloc_act(t, &t->dpre2, fs));
    const size_t B = (size_t)max_batch;
    T_TRY(t_alloc(t, &p, sizeof(float) * KB_PSIZE * B, false));
    t->logits = (float*)p;
    T_TRY(t_alloc(t, &p, sizeof(float) * KB_OBSIZE * B, false));
    t->obs_dev = (float*)p;
    T_TRY(t_alloc(t, &p, sizeof(float) * KB_PSIZE * B, false));
    t->pi_dev = (float*)p;
    T_TRY(t_alloc(t, &p, sizeof(float) * B, false));
    t->z_dev = (float*)p;
    T_TRY(t_alloc(t, &p, sizeof(float) * 64 * B, false));
    t->vpre = (float*)p;
    T_TRY(t_alloc(t, &p, sizeof(float) * 64 * B, false));
    t->vact = (float*)p;
    T_TRY(t_alloc(t, &p, sizeof(float) * 256 * B, false));
    t->dfc = (float*)p;
    T_TRY(t_alloc(t, &p, sizeof(float) * 64 * B, false));
    t->dvn = (float*)p;
    T_TRY(t_alloc(t, &p, sizeof(float) * 4, false));
    t->vstats = (float*)p;
#undef T_TRY
    KB_CUDA(cudaDeviceSynchronize());
    *out = t;
    return KB_OK;
}

int kb_trainer_destroy(kb_trainer* t) {
    if (!t) return KB_OK;
    KB_BIND(t);
    cudaStreamSynchronize(main_stream());
    for (void* p : t->allocs) cudaFree(p);
    delete t;
    return KB_OK;
}

// fp32 master weights and BatchNorm running statistics, blob order of kb_net_load_blob (reference module names)
int kb_trainer_load_blob(kb_trainer* t, const float* blob, size_t n_floats) {
    KB_REQUIRE_INIT();
    KB_ARG(t && blob && n_floats == t->n_floats, "trainer / blob / size");
    KB_BIND(t);
    KB_CUDA(cudaMemcpyAsync(t->params, blob, sizeof(float) * n_floats, cudaMemcpyHostToDevice, main_stream()));
    int r = t_repack(t, main_stream());
    if (r) return r;
    KB_CUDA(cudaStreamSynchronize(main_stream()));
    return KB_OK;
}
int kb_trainer_export_blob(kb_trainer* t, float* blob, size_t n_floats) {
    KB_ARG(t && blob && n_floats == t->n_floats, "trainer / blob / size");
    KB_BIND(t);
    KB_CUDA(cudaStreamSynchronize(main_stream()));
    KB_CUDA(cudaMemcpy(blob, t->params, sizeof(float) * n_floats, cudaMemcpyDeviceToHost));
    return KB_OK;
}
int kb_trainer_export_grads(kb_trainer* t, float* out, size_t n_floats) {
    KB_ARG(t && out && n_floats == t->n_floats, "trainer / out / size");
    KB_BIND(t);
    KB_CUDA(cudaStreamSynchronize(main_stream()));
    KB_CUDA(cudaMemcpy(out, t->grads, sizeof(float) * n_floats, cudaMemcpyDeviceToHost));
    return KB_OK;
}
// device address of the flat fp32 gradient vector: the buffer a data-parallel host all-reduces (NCCL) between
// kb_trainer_forward_backward and kb_trainer_apply_sgd
int kb_trainer_grad_buffer(kb_trainer* t, void** dev_ptr, size_t* n_floats) {
    KB_ARG(t && dev_ptr && n_floats, "trainer / out");
    KB_BIND(t);
    *dev_ptr = t->grads;
    *n_floats = t->n_floats;
    return KB_OK;
}

// forward (training mode) + loss + backward of one mini-batch; host arrays like NN::train's (nn.cpp:224)
int kb_trainer_forward_backward(kb_trainer* t, const float* obs, const float* obs_p, const float* obs_v, int batch, float* loss) {
    KB_REQUIRE_INIT();
    KB_ARG(t && obs && obs_p && obs_v && batch >= 1 && batch <= t->cap, "trainer / arrays / 1 <= batch <= max_batch");
    KB_BIND(t);
    cudaStream_t st = main_stream();
    KB_CUDA(cudaMemcpyAsync(t->obs_dev, obs, sizeof(float) * KB_OBSIZE * (size_t)batch, cudaMemcpyHostToDevice, st));
    KB_CUDA(cudaMemcpyAsync(t->pi_dev, obs_p, sizeof(float) * KB_PSIZE * (size_t)batch, cudaMemcpyHostToDevice, st));
    KB_CUDA(cudaMemcpyAsync(t->z_dev, obs_v, sizeof(float) * (size_t)batch, cudaMemcpyHostToDevice, st));
    int r = t_forward_backward(t, t->obs_dev, t->pi_dev, t->z_dev, batch, st);
    if (r) return r;
    KB_CUDA(cudaStreamSynchronize(st));
    if (loss) *loss = t->last_loss;
    return KB_OK;
}
// the same with the batch already resident (replay samples expanded on the device)
int kb_trainer_forward_backward_dev(kb_trainer* t, const float* obs_dev, const float* obs_p_dev, const float* obs_v_dev, int batch, float* loss) {
    KB_REQUIRE_INIT();
    KB_ARG(t && obs_dev && obs_p_dev && obs_v_dev && batch >= 1 && batch <= t->cap, "trainer / arrays / 1 <= batch <= max_batch");
    KB_BIND(t);
    int r = t_forward_backward(t, obs_dev, obs_p_dev, obs_v_dev, batch, main_stream());
    if (r) return r;
    if (loss) {
        KB_CUDA(cudaStreamSynchronize(main_stream()));
        *loss = t->last_loss;
    }
    return KB_OK;
}
// Test hook: one board of a saved activation as fp32 [channels][64].  which: 0 pre-BatchNorm conv output, 1 layer output
// (after BatchNorm/ReLU/skip), for conv `layer` in [0, 2 + 2R) (conv1, residual convs, policyconv).
int kb_trainer_debug_activation(kb_trainer* t, int layer, int which, int board, float* out, int* channels) {
    KB_ARG(t && out && channels && layer >= 0 && layer < (int)t->conv.size() - 1 && board >= 0 && board < t->batch, "trainer / layer / board");
    KB_BIND(t);
    const TConv& c = t->conv[layer];
    const uint4* buf = which == 0 ? c.pre : c.post;
    KB_CUDA(cudaStreamSynchronize(main_stream()));
    const int item = board / NB, slot = board % NB, slabs = c.out_slabs;
    std::vector<uint16_t> slab((size_t)SLAB_BYTES / 2);
    for (int sl = 0; sl < slabs; ++sl) {
        KB_CUDA(cudaMemcpy(slab.data(), buf + ((size_t)item * slabs + sl) * SLAB_U4, SLAB_BYTES, cudaMemcpyDeviceToHost));
        for (int q = 0; q < 64; ++q)
            for (int ch = 0; ch < 64; ++ch) {
                const int px = tall_pixel(slot, q);
                uint32_t u = (uint32_t)slab[(size_t)chunk_u4(px, ch >> 3) * 8 + (ch & 7)] << 16;
                float f;
                memcpy(&f, &u, 4);
                out[(size_t)(sl * 64 + ch) * 64 + q] = f;
            }
    }
    *channels = slabs * 64;
    return KB_OK;
}

// Where the BatchNorm running statistics sit in the flat parameter vector: (offset, count) per layer -- running_mean and
// running_var are adjacent.  Data-parallel hosts average these ranges over the replicas (kb_dp_*), everything else in the
// vector is trainable and moves by the all-reduced gradient.
int kb_trainer_stat_ranges(kb_trainer* t, size_t* offsets, size_t* counts, int cap, int* n) {
    KB_ARG(t && offsets && counts && n, "trainer / out");
    int k = 0;
    for (const TConv& c : t->conv)
        if (c.bn) {
            if (k < cap) {
                offsets[k] = c.rm_off;
                counts[k] = 2 * (size_t)c.O;
            }
            ++k;
        }
    if (k < cap) {
        offsets[k] = t->vrm_off;
        counts[k] = 2;
    }
    ++k;
    *n = k;
    if (k > cap) {
        set_error("%d statistic ranges, caller's arrays hold %d", k, cap);
        return KB_ERR_CAPACITY;
    }
    return KB_OK;
}
int kb_trainer_param_buffer(kb_trainer* t, void** dev_ptr, size_t* n_floats) {
    KB_ARG(t && dev_ptr && n_floats, "trainer / out");
    KB_BIND(t);
    *dev_ptr = t->params;
    *n_floats = t->n_floats;
    return KB_OK;
}

// plain SGD (nn.cpp:239-241): w -= lr * grad_scale * grad, then the bf16 operand blocks are rebuilt
int kb_trainer_apply_sgd(kb_trainer* t, float lr, float grad_scale) {
    KB_REQUIRE_INIT();
    KB_ARG(t, "trainer");
    KB_BIND(t);
    cudaStream_t st = main_stream();
    k_sgd<<<1024, 256, 0, st>>>(t->params, t->grads, t->n_floats, lr * grad_scale);
    KB_CUDA(cudaGetLastError());
    return t_repack(t, st);
}

}  // extern "C"
