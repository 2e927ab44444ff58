// Device-side data layout of the B-path (bundle adjustment).  See DESIGN.md §"B-path layout".
#pragma once
#include <cstdint>

namespace msfm {
namespace ba {

// Per-camera quantities recomputed once per parameter update (fp64): rotation matrix of Ceres'
// AngleAxisRotatePoint and the four scalar coefficients of its derivative.
struct CamPre {
    double R[9];
    double w[3];
    double t[3];
    double a, b, a1, b1;
    double pad;
};

// A tile = a set of points that together touch at most kTileCams cameras and hold at most kTileObs observations.  One CTA
// linearises one tile at a time and accumulates the tile's share of the reduced camera system in shared memory: the upper
// triangle of 6x6 blocks over the tile's LOCAL camera list, flushed once per tile into the global block-sparse S.
//   normal tile: device points [begin, end) (at most kTilePts), each observed by at most 32 cameras; their observations are
//                the contiguous device range [obs_begin, obs_begin + n_obs);
//   item tile:   items [begin, end) (at most kTileItems, struct Item): a point observed by more than 32 cameras is cut into
//                groups of 16 observations; an item couples group A with group B (or A with itself, which also owns A's
//                diagonal blocks).  Items of neighbouring long points with the same (A, B) group indices are packed together.
struct Tile {
    int32_t begin, end;           // device points (normal) or items (kTileSplit)
    int32_t obs_begin, n_obs;     // normal tiles: device observations [obs_begin, obs_begin + n_obs)
    int32_t cam_begin, w;         // local cameras: tile_cams[cam_begin .. cam_begin + w), ascending global camera index
    int32_t slot_begin;           // tile_slots[slot_begin + lb (lb + 1) / 2 + la] (la <= lb): global block slot or -1
    int32_t flags;                // kTileSplit
    int32_t run_begin, n_runs;    // runs[run_begin .. +n_runs): work items of the pair phase = (run of consecutive units with
                                  // IDENTICAL camera lists, round of 32 camera pairs)
    int32_t pad[2];
};
struct Item {
    int32_t d;                    // device point
    uint16_t a0, a1, b0, b1;      // observation positions [a0, a1) = group A, [b0, b1) = group B (empty: pairs inside A)
    uint16_t pad[2];
    uint8_t lc[32];               // local camera of lane l: lanes 0 .. nA-1 hold group A, nA .. nA+nB-1 group B
};
constexpr int32_t kTileSplit = 1;
// Tile shape.  MSFM_K2_SMALL_TILES=1 (a build-time experiment): half-size tiles with 24 local cameras so that
// TWO 256-thread CTAs of the linearisation kernel fit one SM (102 KB of shared memory each) and overlap each other's barriers.
#ifndef MSFM_K2_SMALL_TILES
#define MSFM_K2_SMALL_TILES 0
#endif
#if MSFM_K2_SMALL_TILES
constexpr int kTileCams = 24;
constexpr int kTileObs = 256;
constexpr int kTilePts = 128;
constexpr int kCtasPerSm = 2;
#else
constexpr int kTileCams = 32;          // local cameras per tile (shared-memory accumulator = 528 blocks)
constexpr int kTileObs = 512;          // observations per tile = threads of the linearisation kernel
constexpr int kTilePts = 256;          // points per normal tile
constexpr int kCtasPerSm = 1;
#endif
constexpr int kLongTrack = kTileCams;  // a point with more observations than this is a long track (cut into item groups)
constexpr int kItemGroup = kTileCams / 2;   // observations per group of a long track: an item couples two groups = at most kTileCams lanes
constexpr int kTileItems = kTileObs / 32;   // items per item tile (one warp of observation lanes each)
constexpr int kBlkStride = 36;         // floats per 6x6 block in the shared-memory accumulator (16-byte aligned rows of 4)

// Offsets (in doubles) into the fp64 "tail" of the system: everything of the reduced system that is not a 6x6 block.
struct TailLayout {
    int32_t scal;      // [8]   0: cost
    int32_t gpm;       // [R]   max |g_p| of every rank (own slot written, the sum all-reduce fills the others)
    int32_t rhs, gc, udiag;   // [n6] each
    int32_t B0, B1, ff;       // shared focal block: two border columns [n6] each, ff[16]
    int32_t total;
};

// SoA view of a BundleData (include/Optimizer/BundleData.h:19-65) flattened for the device.
// "Device order": points are reordered so that neighbours share cameras (pt_order), the observations of a point are
// contiguous and sorted by camera.  pts / cams stay in the caller's order.
struct Problem {
    int32_t n_cams, n_pts, n_obs, n_free;
    double fx, fy;
    const CamPre* pre;          // [n_cams]
    const double* pts;          // [n_pts][3]   caller's order
    const double* obs_uv;       // [n_obs][2]   device order, centred by (cx, cy)
    const int32_t* obs_cam;     // [n_obs]      device order
    const int32_t* obs_pt;      // [n_obs]      device order: the caller's point index
    const int32_t* obs_orig;    // [n_obs]      device order -> the caller's observation index
    const uint8_t* obs_lcam;    // [n_obs]      index of the camera in the local list of the observation's tile
    const uint8_t* obs_lpt;     // [n_obs]      index of the observation's point inside its (normal) tile
    const int32_t* pt_start;    // [n_pts+1]    CSR over device-ordered observations, indexed by DEVICE point
    const int32_t* pt_order;    // [n_pts]      device point -> caller's point index
    const int32_t* cam_free;    // [n_cams] index among the free cameras or -1 (constant pose)
    // ---- tiling + block structure (built once per problem: the sparsity pattern does not change between LM iterations)
    const Tile* tiles;
    const Item* items;
    const uint32_t* runs;           // first unit of the run inside its tile (16 bits) | units (8 bits) | round of 32 pairs (8 bits)
    int32_t n_tiles;
    const int32_t* tile_cams;
    const int32_t* tile_free;       // parallel to tile_cams: cam_free of the local camera (-1: constant pose)
    const int32_t* tile_slots;
    int32_t first_long, n_long;     // device points [first_long, first_long + n_long) are long tracks (more than 32 observations)
    double* long_V;                 // [n_long][9 (+6)]  per long track: V^-1 (6) | g_p (3) (| Wf (6)), written by the pre-pass
    int32_t n_blocks;               // non-empty 6x6 blocks of the upper block triangle (diagonal included)
    const int32_t* blk_row;         // [n_blocks] free-camera row fa
    const int32_t* blk_col;         // [n_blocks] free-camera column fb >= fa
    // ---- the system (per linearisation)
    float* sblk;                    // [n_blocks][36]  fp32 blocks of S = U - sum Y W^T (undamped)
    double* tail;                   // TailLayout
    TailLayout tl;
    int32_t gpm_slot;               // this rank's slot in tail[gpm ..]
    int32_t* tile_counter;          // dynamic tile scheduler
    // ---- shared focal block (BundleAutoDiffCostFunction, CeresBundleOptimizer.cpp:76-121; refine_focal_length)
    int32_t refine_focal;           // 0: (fx, fy) constant; 1: one shared 2-parameter block, bordering the camera system
    double* pt_Wf;                  // [n_pts][6] (device point)  Wf = sum_obs Jf^T Jp (2x3), Jf = d r / d (fx, fy) = diag(xp, yp)
};

}  // namespace ba
}  // namespace msfm
