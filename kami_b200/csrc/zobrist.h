// Host-side generation of the zobrist key table uploaded to the device at kb_init.
// The reference draws its keys from glibc rand() with the default seed before main()
// (zobrist.c:25-54, env.h:25-39): 793 keys x 8 draws of `rand() & 0xFF`.
// glibc's rand() is the TYPE_3 additive-feedback generator r[i] = r[i-3] + r[i-31]; it is
// restated here so the keys (and therefore repetition detection, even under hash
// collisions) are bit-identical to the reference's without depending on process rand state.
#pragma once
#include <stdint.h>

namespace kb {

inline void make_zobrist(uint64_t* out, int count) {
    uint32_t r[344 + 31];
    int32_t x = 1;
    r[0] = 1;
    for (int i = 1; i < 31; ++i) {
        int64_t v = (16807LL * x) % 2147483647LL;
        if (v < 0) v += 2147483647LL;
        x = (int32_t)v;
        r[i] = (uint32_t)x;
    }
    for (int i = 31; i < 34; ++i) r[i] = r[i - 31];
    for (int i = 34; i < 344; ++i) r[i] = r[i - 31] + r[i - 3];
    uint32_t ring[31];
    for (int i = 0; i < 31; ++i) ring[i] = r[344 - 31 + i];
    int idx = 0;
    for (int k = 0; k < count; ++k) {
        uint64_t key = 0;
        for (int b = 0; b < 8; ++b) {
            uint32_t v = ring[idx] + ring[(idx + 28) % 31];
            ring[idx] = v;
            idx = (idx + 1) % 31;
            key |= (uint64_t)((v >> 1) & 0xFF) << (8 * b);
        }
        out[k] = key;
    }
}

}  // namespace kb
