// Device-side data layout of the M-path (descriptor matching).  See DESIGN.md §"M-path layout".
#pragma once
#include <cstdint>

namespace msfm {

// Resident descriptor set of one image.
//   sw : [n_pad][128] u8, n_pad = ceil(n/256)*256.  Row r is stored at sw + r*128 with its 16-byte chunk c
//        at chunk position c ^ (r & 7): this IS the tcgen05 K-major SWIZZLE_128B shared-memory layout, so a
//        1-D bulk copy (TMA engine) of 128/256 consecutive rows lands in smem ready for the MMA.
//        Rows >= n are zero.
//   cj : [n_pad] s32 "column constant" = ||row||^2 * 256 + (r & 255); rows >= n hold kPadKey | (r & 255).
struct ImgDev {
    const uint8_t* sw;
    const int32_t* cj;
    int32_t n;
    int32_t n_pad;
};

// One (pair, direction): rows of image q are matched against columns of image t.
struct SegDev {
    int32_t q_slot;
    int32_t t_slot;
    int32_t unit_base;   // first work unit (128-row block) of the segment inside its batch
    int32_t n_units;
};

// One work unit of the tensor kernel: 128 query rows x all train columns.
// Global row index of (unit u, row r) inside the batch scratch arrays is u*128 + r.
struct UnitDev {
    int32_t q_slot;
    int32_t t_slot;
    int32_t row_block;
    int32_t seg;
};

constexpr int32_t kPadKey = 0x7FFFFF00;      // larger than every valid packed key (max 2 130 739 455)
constexpr int32_t kIntInf = 0x7FFFFFFF;
constexpr int32_t kSqrtExactLimit = 1 << 22; // below this, distinct integers have distinct float sqrt

struct MatchOpts {
    double max_distance;  // < 0: off
    float ratio;
    int32_t cross_check;
    int32_t quirks;
    int32_t exact_second; // 1: always rescan the winner's group (knn2 API); 0: only ratio candidates
};

}  // namespace msfm
