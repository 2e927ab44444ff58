// Device-side data layout of the M-path (descriptor matching).  See DESIGN.md §"M-path layout".
#pragma once
#include <cstdint>

namespace msfm {

// Resident descriptor set of one image, stored in "sorted space".
//
// Columns (descriptors) are permuted so that every 32-column GROUP holds descriptors of equal squared-norm parity
// and a bounded norm spread: sort key = (bucket, ||d||^2, original index) with bucket = parity*3 + ||d||^2 / 4e6;
// every bucket is padded with dead columns to a multiple of 32, the whole set to n_pad (a multiple of 256).
// This lets the tensor core emit directly comparable scores: with C_g = max ||d_j||^2 over group g and
// e_j = (C_g - ||d_j||^2) / 2  (an integer in [0, 2 016 029]) encoded in 32 extra K bytes,
//     acc(i, j) = q_i . d_j + e_j        and       C_g - 2 * max_j acc(i, j) = min_j (||d_j||^2 - 2 q_i . d_j),
// so the per-element epilogue is a bare integer max.
//
//   sw   : [n_pad][128] u8.  Row p at sw + p*128, its 16-byte chunk c at chunk position c ^ (p & 7): the tcgen05
//          K-major SWIZZLE_128B shared-memory layout, so a 1-D bulk copy of consecutive rows lands MMA-ready.
//          Dead rows are zero.
//   ext  : [ceil(n_pad/1024)] super-tiles of [256 rows][128 B], same swizzled layout.  Bytes [32q, 32q+32) of row r
//          of super-tile s hold the digits of e_j for column p = (4s+q)*256 + r:  e = b0 + 255 * (b1 + ... + b31).
//          The matching A-side constant row is (1, 255, 255, ..., 255).
//   cg   : [n_pad/32] s32  C_g (kDeadCg for groups without a real column)
//   nrm  : [n_pad] s32     ||d||^2 of the column at sorted position p, -1 for dead columns
//   perm : [n_pad] s32     original index of the column at sorted position p, -1 for dead columns
//   inv  : [n] s32         sorted position of the column with original index j
//   used : [1] s32         number of sorted positions actually occupied (sum of the padded bucket sizes, a multiple
//                          of 32, <= n_pad): written on the device by the formatting kernels, read by K1 to skip the
//                          all-dead tail (n_pad itself is the host-known upper bound n + 186 rounded to 256)
struct ImgDev {
    const uint8_t* sw;
    const uint8_t* ext;
    const int32_t* cg;
    const int32_t* nrm;
    const int32_t* perm;
    const int32_t* inv;
    const int32_t* used;
    int32_t n;
    int32_t n_pad;
};

// One (pair, direction): rows of image q are matched against columns of image t.
struct SegDev {
    int32_t q_slot;
    int32_t t_slot;
    int32_t unit_base;   // first work unit (128-row block of q's sorted space) of the segment inside its batch
    int32_t n_units;     // q.n_pad / 128
};

// One work unit of the tensor kernel: kUnitRows query rows (sorted space) x all train columns.
// K1 results of (unit u, row r) live at u*kUnitRows + r; final per-row results at
// (u - row_block)*kUnitRows + ORIGINAL query index.
struct UnitDev {
    int32_t q_slot;
    int32_t t_slot;
    int32_t row_block;
    int32_t seg;
};

constexpr int32_t kUnitRows = 128;
constexpr int32_t kDeadCg = 0x3FFFFFFF;       // C_g of an all-dead group: its key can never win
constexpr int32_t kDeadKey = 0x30000000;      // keys >= this come from dead groups (real keys are |.| <= 8 323 200)
constexpr int32_t kIntInf = 0x7FFFFFFF;
constexpr int32_t kSqrtExactLimit = 1 << 22;  // below this, distinct integers have distinct float sqrt
constexpr int32_t kBucketSpan = 4000000;      // norm span of one sort bucket; (span / 2) must fit the digit encoding
constexpr int32_t kNumBuckets = 6;            // 2 parities x ceil(8 323 200 / 4e6)
constexpr int32_t kMaxPadPerImage = kNumBuckets * 31;

struct MatchOpts {
    double max_distance;  // < 0: off
    float ratio;
    int32_t cross_check;
    int32_t quirks;
    int32_t exact_second; // 1: always rescan the winner's group (knn2 API); 0: only ratio candidates
};

}  // namespace msfm
