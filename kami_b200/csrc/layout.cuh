// Activation layout shared by the encoder, the tcgen05 conv kernels and the heads.
//
// "Tall image": boards are processed in items of NB = 7.  Inside an item every group of 8
// channels (16 B of bf16 per pixel) is one plane of ROWS x PITCH pixels:
//
//     row 0            zero pad
//     rows 1..8        board 0 (8 pixels at columns 1..8, columns 0 and 9 are zero pad)
//     row 9            zero pad (bottom of board 0 == top of board 1)
//     rows 10..17      board 1
//     ...
//     row 63           zero pad below board 6
//
// so a 3x3 tap (dy,dx) of an implicit-GEMM conv is nothing but a byte offset of
// (dy*PITCH + dx)*16 into the same buffer, and 16 consecutive rows x 8 columns are the
// 128 rows of one UMMA M tile in the canonical K-major no-swizzle layout
// (core matrix = 8 pixels x 16 B, SBO = PITCH*16, LBO = plane size).  The same bytes
// live in global memory and in shared memory, so every transfer is a 1-D bulk TMA copy.
// Pad pixels are zero-initialised once and never written.
#pragma once
#include <stdint.h>

namespace kb {

constexpr int NB = 7;                  // boards per item
constexpr int TALL_ROWS = 64;          // 1 + 7*9
constexpr int TALL_PITCH = 10;         // pixels per tall row
constexpr int PLANE_PIX = TALL_ROWS * TALL_PITCH;  // 640
constexpr int PLANE_BYTES = PLANE_PIX * 16;        // 10240: one 8-channel plane
constexpr int IN_CHUNKS = 4;           // 30 input features padded to 32 channels

__host__ __device__ inline int items_for(int boards) { return (boards + NB - 1) / NB; }
// pixel index of board slot s (0..6), square q (rank*8+file) inside a plane
__host__ __device__ inline int tall_pixel(int slot, int q) { return (1 + 9 * slot + (q >> 3)) * TALL_PITCH + 1 + (q & 7); }
// bytes of an activation tensor with `chunks` 8-channel planes per item (+1 plane of slack so
// that the garbage rows an M tile touches past the last plane stay inside the allocation)
__host__ __device__ inline size_t act_bytes(int boards, int chunks) {
    return ((size_t)items_for(boards) * chunks + 1) * PLANE_BYTES + 1024;
}

}  // namespace kb
