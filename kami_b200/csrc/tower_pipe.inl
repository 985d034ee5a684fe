// k_tower64p -- k_tower64 with the layer hand-over broken up per M tile.  Included in net.cu (namespace kb).
//
// k_tower64 alternates strictly: the MMA warp issues a whole layer (4 M tiles x 36 k-steps), then the epilogue warps
// drain all four accumulators, then the next layer may start -- 9 k cycles of MMAs with the epilogue warps idle, then
// 2.3-3.2 k cycles of epilogue with the tensor pipe idle, seven times per item.  Here:
//
//   * 3x3 layers are issued as two halves, tiles {0,1} then {2,3}; the weights of the layer are streamed through the ring
//     once per half.  The epilogue of half 0 runs under the MMAs of half 1.
//   * The epilogue goes tile by tile with ALL 16 epilogue warps on one tile (a warp = one TMEM lane quarter x one group
//     of 16 accumulator columns) and signals each tile on its own barrier.  A 3x3 tile of the next layer needs rows -1..+16
//     of its own range: half 0 of layer l+1 starts when tiles 0, 1 and 2 of layer l are written, i.e. while tile 3 is still
//     being drained; half 1 waits for tile 3.  Exposed per layer: the epilogue of ONE tile instead of four.
//   * The 1x1 head layers need no halo.  Their weight blocks (2 + 2) stay in the ring for the whole layer and the MMAs are
//     issued tile-major: policyconv tile t starts behind the last tower layer's tile t, the logits conv tile t behind
//     policyconv's tile t.  (The logits overlay the activations, so the logits EPILOGUE still waits for all logits MMAs.)
//
// Same shared-memory layout, weights, numerics and outputs as k_tower64 (tests run both and compare bit for bit).

constexpr int FP_EW = 16;                   // epilogue warps
constexpr int FP_THREADS = 128 + 32 * FP_EW;

__global__ void __launch_bounds__(FP_THREADS, 1) k_tower64p(const __grid_constant__ FusedParams P) {
    constexpr int EW = FP_EW, FZ_EPI_THREADS = 32 * EW, FC_BAR = FZ_EPI_THREADS + 64;
    extern __shared__ __align__(1024) uint8_t smem[];
    const int warp = ptx::uniform_warp_id(), lane = threadIdx.x & 31;
    const uint32_t s0 = ptx::smem_u32(smem);
    auto b_full = [&](int s) { return s0 + 8u * s; };
    auto b_empty = [&](int s) { return s0 + 8u * (8 + s); };
    const uint32_t p_full = s0 + 8u * 16, region_clean = s0 + 8u * 17;
    auto tf = [&](int h) { return s0 + 8u * (20 + h); };    // tower half h: accumulators complete
    auto tfh = [&](int t) { return s0 + 8u * (24 + t); };   // policyconv tile t: accumulators complete
    const uint32_t tf6 = s0 + 8u * 28;                      // logits conv: all accumulators complete
    auto ar = [&](int t) { return s0 + 8u * (32 + t); };    // tower layer: tile t written to shared memory, its accumulators drained
    auto ar5 = [&](int t) { return s0 + 8u * (36 + t); };   // policyconv: tile t of H written
    const uint32_t ar6 = s0 + 8u * 40;                      // logits drained from TMEM
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 384);
    float* vbuf = reinterpret_cast<float*>(smem + 512);     // [7][64] value-conv outputs
    float* sred = reinterpret_cast<float*>(smem + 2304);    // [EW][7] block-reduction scratch of the dense softmax
    float* sbias = reinterpret_cast<float*>(smem + 2816);   // all folded biases (n_bias <= 1340 floats)
    uint8_t* region = smem + FZ_HDR;
    const uint32_t region_s = s0 + FZ_HDR;
    const uint32_t ring_s = region_s + FZ_REGION;
    auto epi_bar = [] { named_bar_sync<FZ_EPI_THREADS>(1); };
    const int TL = P.tower_layers;

    const int my_items = P.items > (int)blockIdx.x ? (P.items - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    if (P.ts && threadIdx.x == 0 && blockIdx.x < 160) {
        long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        P.ts[128 + 2 * blockIdx.x] = t;
    }
    if (threadIdx.x == 0) {
        for (int s = 0; s < FZ_NSTAGE; ++s) {
            ptx::mbar_init(b_full(s), 1);
            ptx::mbar_init(b_empty(s), 1);
        }
        ptx::mbar_init(p_full, 1);
        ptx::mbar_init(region_clean, EW);
        ptx::mbar_init(tf(0), 1);
        ptx::mbar_init(tf(1), 1);
        ptx::mbar_init(tf6, 1);
        ptx::mbar_init(ar6, EW);
        for (int t = 0; t < 4; ++t) {
            ptx::mbar_init(tfh(t), 1);
            ptx::mbar_init(ar(t), EW);
            ptx::mbar_init(ar5(t), EW);
        }
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

    if (warp == 0) {
        // ===== producer: input slab, then every weight block in consumption order (3x3 layers twice: once per half) =====
        int stage = 0, sphase = 0;
        for (int ii = 0; ii < my_items; ++ii) {
            const int item = (int)blockIdx.x + ii * (int)gridDim.x;
            ptx::mbar_wait(region_clean, ii & 1);
            if (ii == 0) pdl_wait();
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
                const int reps = l < TL ? 2 : 1;
                for (int rep = 0; rep < reps; ++rep) {
                    const uint4* wl = w;
                    for (int b = 0; b < nblocks; ++b) {
                        ptx::mbar_wait(b_empty(stage), sphase ^ 1);
                        if (ptx::elect_one()) {
                            ptx::mbar_arrive_expect_tx(b_full(stage), bytes);
                            ptx::bulk_g2s(ring_s + stage * FZ_STAGE, wl, bytes, b_full(stage));
                        }
                        __syncwarp();
                        wl += bytes / 16;
                        if (++stage == FZ_NSTAGE) {
                            stage = 0;
                            sphase ^= 1;
                        }
                    }
                }
                w += (size_t)nblocks * (bytes / 16);
            }
        }
        pdl_launch_dependents();
    } else if (warp == 1) {
        // ===== MMA issuer (converged warp, one elected lane issues) =====
        int stage = 0, sphase = 0;
        uint32_t ar_ph = 0;  // bit t: parity of the next completion of ar(t) this warp waits for
        auto wait_ar = [&](int t) {
            ptx::mbar_wait(ar(t), (ar_ph >> t) & 1u);
            ar_ph ^= 1u << t;
        };
        const uint32_t a_hi = ptx::sw128_hi(TALL_PITCH * LINE_BYTES), b_hi = ptx::sw128_hi(1024);
        constexpr uint32_t TILE_STEP = 16 * TALL_PITCH * LINE_BYTES / 16;  // descriptor units between M tiles
        pdl_wait();
        for (int ii = 0; ii < my_items; ++ii) {
            // ---- the 3x3 tower: two halves per layer ----
            for (int l = 0; l < TL; ++l) {
                const FusedLayer& L = P.layer[l];
                const uint32_t a_src = region_s + L.src_off;
                const int ksteps = L.ksteps, ntaps = L.ntaps, n = L.n;
                const uint32_t idesc = L.idesc;
                for (int h = 0; h < 2; ++h) {
                    if (l == 0) {
                        if (h == 0) ptx::mbar_wait(p_full, ii & 1);
                    } else if (h == 0) {  // rows -1..32 of the previous layer: tiles 0, 1 and the first row of tile 2
                        wait_ar(0);
                        wait_ar(1);
                        wait_ar(2);
                    } else {              // rows 31..64: tile 3 (tiles 1 and 2 were waited for above)
                        wait_ar(3);
                    }
                    ptx::tc_fence_after();
                    for (int ks = 0; ks < L.slabs_in; ++ks) {
                        const uint32_t a_lo0 = ptx::sw128_lo(a_src + ks * SLAB_BYTES);
                        int dy = ntaps == 9 ? -1 : 0, dx = ntaps == 9 ? -1 : 0;
                        for (int tap = 0; tap < ntaps; ++tap) {
                            const uint32_t a_tap = a_lo0 + (uint32_t)((dy * TALL_PITCH + dx + 1) * (LINE_BYTES / 16));
                            ptx::mbar_wait(b_full(stage), sphase);
                            ptx::tc_fence_after();
                            const uint32_t b_lo0 = ptx::sw128_lo(ring_s + stage * FZ_STAGE);
                            if (ptx::elect_one()) {
                                uint32_t first = (ks | tap) == 0 ? 0u : 1u;
                                for (int kk = 0; kk < ksteps; ++kk) {
                                    const uint64_t bdesc = ptx::desc_pack(b_lo0 + kk * 2, b_hi);
#pragma unroll
                                    for (int m2 = 0; m2 < 2; ++m2) {
                                        const int mt = 2 * h + m2;
                                        const uint64_t adesc = ptx::desc_pack(a_tap + kk * 2 + mt * TILE_STEP, a_hi);
                                        ptx::mma_bf16(tmem_base + mt * n, adesc, bdesc, idesc, first);
                                    }
                                    first = 1u;
                                }
                                ptx::mma_commit(b_empty(stage));
                            }
                            __syncwarp();
                            if (++stage == FZ_NSTAGE) {
                                stage = 0;
                                sphase ^= 1;
                            }
                            if (++dx > 1) {
                                dx = -1;
                                ++dy;
                            }
                        }
                    }
                    if (ptx::elect_one()) ptx::mma_commit(tf(h));
                    __syncwarp();
                }
            }
            // ---- 1x1 heads, tile-major with their weight blocks resident in the ring ----
            for (int l = TL; l < P.n_layers; ++l) {
                const FusedLayer& L = P.layer[l];
                const bool logits = L.kind == 1;
                const uint32_t a_src = region_s + L.src_off;
                const int nsub = L.n / L.n_sub, nblk = nsub * L.slabs_in, ksteps = L.ksteps, n = L.n;
                const uint32_t idesc = L.idesc;
                // the layer's blocks, in arrival order (sub-block, slab): wait for all of them once
                const int stage0 = stage;  // block b sits in ring stage (stage0 + b) mod FZ_NSTAGE
                for (int b = 0; b < nblk; ++b) {
                    ptx::mbar_wait(b_full(stage), sphase);
                    if (++stage == FZ_NSTAGE) {
                        stage = 0;
                        sphase ^= 1;
                    }
                }
                ptx::tc_fence_after();
                for (int mt = 0; mt < 4; ++mt) {
                    if (!logits) {
                        // tile mt of the last tower layer written (no halo), and the TMEM columns [mt * n, +n) it overwrites
                        // drained: they held that layer's tiles 2 mt and 2 mt + 1 (64 columns each)
                        if (mt == 0) {
                            wait_ar(0);
                            wait_ar(1);
                        } else if (mt == 1) {
                            wait_ar(2);
                            wait_ar(3);
                        }
                    } else {
                        ptx::mbar_wait(ar5(mt), ii & 1);  // tile mt of H written; policyconv's accumulators up to tile mt drained
                    }
                    ptx::tc_fence_after();
                    if (ptx::elect_one()) {
                        for (int sub = 0; sub < nsub; ++sub)
                            for (int ks = 0; ks < L.slabs_in; ++ks) {
                                int bs = stage0 + sub * L.slabs_in + ks;
                                if (bs >= FZ_NSTAGE) bs -= FZ_NSTAGE;
                                const uint32_t a_lo = ptx::sw128_lo(a_src + ks * SLAB_BYTES) + (uint32_t)(LINE_BYTES / 16) + mt * TILE_STEP;
                                const uint32_t b_lo0 = ptx::sw128_lo(ring_s + bs * FZ_STAGE);
                                for (int kk = 0; kk < ksteps; ++kk)
                                    ptx::mma_bf16(tmem_base + sub * L.n_sub + mt * n, ptx::desc_pack(a_lo + kk * 2, a_hi),
                                                  ptx::desc_pack(b_lo0 + kk * 2, b_hi), idesc, (ks | kk) == 0 ? 0u : 1u);
                            }
                        if (!logits) ptx::mma_commit(tfh(mt));
                    }
                    __syncwarp();
                }
                if (ptx::elect_one()) {
                    if (logits) ptx::mma_commit(tf6);
                    for (int b = 0; b < nblk; ++b) {
                        int bs = stage0 + b;
                        if (bs >= FZ_NSTAGE) bs -= FZ_NSTAGE;
                        ptx::mma_commit(b_empty(bs));
                    }
                }
                __syncwarp();
            }
            if (ii + 1 == my_items) pdl_launch_dependents();
            ptx::mbar_wait(ar6, ii & 1);  // the logits left TMEM: the next item's first layer may overwrite the columns
        }
    } else if (warp == 2 || warp == 3) {
        // ===== value head, second half: Linear(64 -> 256) + tanh (nn.cpp:87-88), off the epilogue warps' critical path =====
        const int j = threadIdx.x - 64;  // 0..63
        pdl_wait();
        for (int ii = 0; ii < my_items; ++ii) {
            const int item = (int)blockIdx.x + ii * (int)gridDim.x;
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(P.fcb) + j);
            float o[NB][4];
#pragma unroll
            for (int s = 0; s < NB; ++s) { o[s][0] = b4.x; o[s][1] = b4.y; o[s][2] = b4.z; o[s][3] = b4.w; }
            const float4* wt = reinterpret_cast<const float4*>(P.fct) + j;  // fct[p][256]
            constexpr int FCB = 8;
            float4 wq[FCB];
#pragma unroll
            for (int p = 0; p < FCB; ++p) wq[p] = __ldg(wt + p * 64);
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
        // ===== epilogue warps: lane quarter q, column group `part` (+4, +8, ...) of every tile =====
        const int e = warp - 4, q = e & 3, part = e >> 2;
        const int et = threadIdx.x - 128;
        uint32_t tf_ph = 0;  // bit h: parity of the next completion of tf(h)
        const bool stamp = P.ts && blockIdx.x == 0 && et == 0;
        int nts = 0;
#define KB_STAMP() do { if (stamp && nts < 60) P.ts[nts++] = clock64(); } while (0)
        // one M tile of layer L: this warp's 32 accumulator rows x its column groups -> bias, ReLU, (skip), store
        auto epi_tile = [&](const FusedLayer& L, int mt) {
            const int ncg = L.n / 16;
            const int r = 32 * q + lane;
            const int R = 16 * mt + (r >> 3), x = r & 7;
            const int slot = (R - 1) / 9, y = (R - 1) - slot * 9;
            const bool valid = R >= 1 && y < 8 && slot < NB;
            const int px = R * TALL_PITCH + 1 + x;
            const uint32_t taddr = tmem_base + ((uint32_t)(32 * q) << 16) + mt * L.n;
            uint32_t va[16], vb[16];
            const int cga = part, cgb = part + 4;  // ncg <= 8: at most two groups per warp
            ptx::tmem_ld16(taddr + cga * 16, va);
            if (cgb < ncg) ptx::tmem_ld16(taddr + cgb * 16, vb);
            ptx::tmem_ld_wait();
            if (!valid) return;
            auto group = [&](const uint32_t (&v)[16], int cg) {
                float f[16];
                const float4* b4 = reinterpret_cast<const float4*>(sbias + L.bias_off + cg * 16);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float4 bb = b4[j];
                    f[4 * j + 0] = __uint_as_float(v[4 * j + 0]) + bb.x;
                    f[4 * j + 1] = __uint_as_float(v[4 * j + 1]) + bb.y;
                    f[4 * j + 2] = __uint_as_float(v[4 * j + 2]) + bb.z;
                    f[4 * j + 3] = __uint_as_float(v[4 * j + 3]) + bb.w;
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
                    for (int hh = 0; hh < 2; ++hh) {
                        const int ch = cg * 16 + 8 * hh;
                        uint4* dst = reinterpret_cast<uint4*>(region + L.dst_off) + (size_t)(ch >> 6) * SLAB_U4 + chunk_u4(px, (ch >> 3) & 7);
                        if (L.skip) {  // x = skip + relu(...), the skip is the destination itself (nn.cpp:31)
                            const uint4 s4 = *dst;
                            const __nv_bfloat162* sb = reinterpret_cast<const __nv_bfloat162*>(&s4);
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const float2 sv = __bfloat1622float2(sb[k]);
                                f[hh * 8 + 2 * k] += sv.x;
                                f[hh * 8 + 2 * k + 1] += sv.y;
                            }
                        }
                        uint32_t w4[4];
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const __nv_bfloat162 b = __floats2bfloat162_rn(f[hh * 8 + 2 * k], f[hh * 8 + 2 * k + 1]);
                            w4[k] = *reinterpret_cast<const uint32_t*>(&b);
                        }
                        *dst = make_uint4(w4[0], w4[1], w4[2], w4[3]);
                    }
                }
            };
            group(va, cga);
            if (cgb < ncg) group(vb, cgb);
        };
        // this warp's part of a tile is in shared memory (visible to tcgen05) and out of TMEM
        auto tile_done = [&](uint32_t bar) {
            fence_async_smem();
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(bar);
        };
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
            if (ii == 0) pdl_wait();
            KB_STAMP();
            for (int l = 0; l < TL; ++l) {
                const FusedLayer& L = P.layer[l];
#pragma unroll 1
                for (int mt = 0; mt < 4; ++mt) {
                    if ((mt & 1) == 0) {  // tiles {0,1} / {2,3} become ready together
                        const int h = mt >> 1;
                        ptx::mbar_wait(tf(h), (tf_ph >> h) & 1u);
                        tf_ph ^= 1u << h;
                        ptx::tc_fence_after();
                        if (h == 1) KB_STAMP();
                    }
                    epi_tile(L, mt);
                    tile_done(ar(mt));
                }
                KB_STAMP();
                if (l == TL - 1) {
                    // ---- value conv 1x1 + ReLU on X (nn.cpp:83-86); the 64 -> 256 Linear + tanh runs on warps 2-3 ----
                    epi_bar();  // X complete
                    if (ii > 0) named_bar_sync<FC_BAR>(3);  // vbuf consumed by the previous item's Linear
                    const uint4* X = reinterpret_cast<const uint4*>(region + L.dst_off);
                    for (int i = et; i < NB * 64; i += FZ_EPI_THREADS) {
                        const int slot = i >> 6, pix = i & 63;
                        const int px = tall_pixel(slot, pix);
                        float acc = P.bv;
#pragma unroll
                        for (int c = 0; c < 8; ++c) {
                            const uint4 a4 = X[chunk_u4(px, c)];
                            const __nv_bfloat162* ab = reinterpret_cast<const __nv_bfloat162*>(&a4);
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const float2 av = __bfloat1622float2(ab[k]);
                                acc = fmaf(av.x, __ldg(P.wv + c * 8 + 2 * k), acc);
                                acc = fmaf(av.y, __ldg(P.wv + c * 8 + 2 * k + 1), acc);
                            }
                        }
                        vbuf[i] = fmaxf(acc, 0.0f);
                    }
                    named_bar_arrive<FC_BAR>(2);  // vbuf ready for warps 2-3
                }
            }
            // ---- policyconv (X -> H, which overlays X | Y): nobody reads X any more once every warp is past the value conv ----
            epi_bar();
            {
                const FusedLayer& L = P.layer[TL];
#pragma unroll 1
                for (int mt = 0; mt < 4; ++mt) {
                    ptx::mbar_wait(tfh(mt), ii & 1);
                    ptx::tc_fence_after();
                    epi_tile(L, mt);
                    tile_done(ar5(mt));
                }
            }
            KB_STAMP();
            // ---- logits conv (H -> fp32 logits over the same shared memory): every MMA of the layer must have read H first ----
            {
                const FusedLayer& L = P.layer[TL + 1];
                ptx::mbar_wait(tf6, ii & 1);
                ptx::tc_fence_after();
                KB_STAMP();
#pragma unroll 1
                for (int mt = 0; mt < 4; ++mt) epi_tile(L, mt);
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(ar6);
            }
            // ---- softmax (nn.cpp:80) ----
            epi_bar();
            if (ii + 1 == my_items) pdl_launch_dependents();
            KB_STAMP();
            if (P.legal_act) {
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
                // every thread owns PER logits of each of the 7 boards; two block-wide reductions in total
                float* red = sred;  // [EW warps][7] maxima, then [EW][7] sums
                const float* lg = reinterpret_cast<const float*>(region);
                constexpr int PER = (KB_PSIZE + FZ_EPI_THREADS - 1) / FZ_EPI_THREADS;
                float m[NB];
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
