// Activation layout shared by the encoder, the tcgen05 conv kernels and the heads.
//
// "Tall image", 128-byte swizzled.  Boards are processed in items of NB = 7.  Channels are
// grouped in slabs of 64 (128 B of bf16 per pixel); inside an item a slab is an image of
// TALL_ROWS x TALL_PITCH pixels, ONE 128-BYTE LINE PER PIXEL:
//
//     row 0            zero pad
//     rows 1..8        board 0 (8 pixels at columns 1..8, columns 0 and 9 are zero pad)
//     row 9            zero pad (bottom of board 0 == top of board 1)
//     rows 10..17      board 1
//     ...
//     row 63           zero pad below board 6
//
// A 3x3 tap (dy,dx) of an implicit-GEMM conv is a byte offset of (dy*PITCH + dx)*128 into the
// same slab -- always 128-byte aligned -- and 16 consecutive rows x 8 columns are the 128 rows of
// one UMMA M tile in the canonical K-major SWIZZLE_128B layout (8-row group = 8 consecutive
// pixels = 1024 contiguous bytes, SBO = PITCH*128).  The 16-byte chunk holding channels
// 8j..8j+7 of pixel p sits at chunk slot j ^ (p & 7): tcgen05 applies the 128B swizzle to the
// ABSOLUTE shared-memory address (bits [4,7) ^= bits [7,10)), which was verified on B200 for
// start addresses that are only 128-byte aligned and for SBO = 1280 (scratch experiment,
// DESIGN.md section 4); slabs are 1024-byte aligned in shared memory, so p & 7 is that phase.
// The same bytes live in global memory, so every transfer is a 1-D bulk TMA copy.  Pad pixels
// are zero-initialised once and never written.
#pragma once
#include <stdint.h>

namespace kb {

constexpr int NB = 7;                  // boards per item
constexpr int TALL_ROWS = 64;          // 1 + 7*9
constexpr int TALL_PITCH = 10;         // pixels per tall row
constexpr int SLAB_PIX = TALL_ROWS * TALL_PITCH;  // 640
constexpr int LINE_BYTES = 128;                   // one pixel: 64 bf16 channels
constexpr int SLAB_BYTES = SLAB_PIX * LINE_BYTES; // 81920, a multiple of 1024
constexpr int SLAB_U4 = SLAB_BYTES / 16;
constexpr int IN_SLABS = 1;            // 30 input features live in channels 0..29 of one slab

__host__ __device__ inline int items_for(int boards) { return (boards + NB - 1) / NB; }
// pixel index of board slot s (0..6), square q (rank*8+file) inside a slab
__host__ __device__ inline int tall_pixel(int slot, int q) { return (1 + 9 * slot + (q >> 3)) * TALL_PITCH + 1 + (q & 7); }
// uint4 index (inside a slab) of the 16-byte chunk holding channels 8j..8j+7 of pixel px
__host__ __device__ inline int chunk_u4(int px, int j) { return px * 8 + (j ^ (px & 7)); }
// bytes of an activation tensor with `slabs` 64-channel slabs per item
__host__ __device__ inline size_t act_bytes(int boards, int slabs) { return (size_t)items_for(boards) * slabs * SLAB_BYTES + 1024; }

}  // namespace kb
