// Host-side structure analysis of a bundle-adjustment problem (once per sparsity pattern): device order of the points
// and observations, tiles, and the block structure of the reduced camera system.  Plain C++ (no CUDA) so that the CPU test
// suite can check its invariants (tests/test_ba_tiles.py).
//
// What it replaces in the reference: Ceres' Program / block-structure detection for SchurEliminator behind
// ceres::Solve (src/Optimizer/CeresBundleOptimizer.cpp:293) — e-blocks = points, f-blocks = cameras.
#pragma once
#include <algorithm>
#include <cstdint>
#include <numeric>
#include <vector>

#include "ba_types.cuh"

namespace msfm {
namespace ba {

struct TilingParams {
    int w_cap = 32;            // max local cameras per tile (<= kMaxWCap)
    int max_pts = 256;         // max points per tile
    long long max_work = 1 << 15;   // max sum of k (k + 1) / 2 per tile
};

struct Tiling {
    std::vector<int32_t> pt_order;      // device point -> caller's point
    std::vector<int32_t> pt_start;      // [n_pts + 1] CSR over device-ordered observations
    std::vector<int32_t> obs_perm;      // device observation -> caller's observation
    std::vector<uint8_t> obs_lcam;      // device observation -> local camera of its (normal) tile
    std::vector<Tile> tiles;
    std::vector<int32_t> tile_cams;
    std::vector<int32_t> tile_marks;    // per tile, tri(w) entries: 1 if some point of the tile couples the two local cameras
    std::vector<int32_t> tile_slots;    // same shape: global block slot or -1 (assign_slots)
    int w_max = 0;
    // block structure (upper block triangle incl. the diagonal), sorted by (row, col)
    std::vector<int32_t> blk_rowptr, blk_row, blk_col;
};

inline int tri_index(int la, int lb) { return lb * (lb + 1) / 2 + la; }   // la <= lb

// Device order + tiles + per-tile coupling marks.  obs_pt must be non-decreasing (validated by the caller).
// Returns false if a point is observed twice by one camera.
inline bool build_tiling(int n_cams, int n_pts, int n_obs, const int32_t* obs_cam, const int32_t* obs_pt,
                         const int32_t* cam_free, const TilingParams& prm, Tiling& T) {
    std::vector<int32_t> start(size_t(n_pts) + 1, 0);
    for (int i = 0; i < n_obs; ++i) start[size_t(obs_pt[i]) + 1] += 1;
    for (int p = 0; p < n_pts; ++p) start[size_t(p) + 1] += start[p];
    // observations of every point sorted by camera
    std::vector<int32_t> sorted_obs(static_cast<size_t>(n_obs));
    std::iota(sorted_obs.begin(), sorted_obs.end(), 0);
    std::vector<uint64_t> key(static_cast<size_t>(n_pts));
    for (int p = 0; p < n_pts; ++p) {
        int32_t* b = sorted_obs.data() + start[p];
        int32_t* e = sorted_obs.data() + start[size_t(p) + 1];
        std::sort(b, e, [&](int32_t x, int32_t y) { return obs_cam[x] < obs_cam[y]; });
        for (int32_t* q = b; q + 1 < e; ++q)
            if (obs_cam[q[0]] == obs_cam[q[1]]) return false;
        // locality key: the three smallest cameras (21 bits each); points without observations go last
        uint64_t k = ~uint64_t(0);
        if (e > b) {
            const uint64_t c0 = uint64_t(obs_cam[b[0]]) & 0x1FFFFF;
            const uint64_t c1 = e - b > 1 ? uint64_t(obs_cam[b[1]]) & 0x1FFFFF : c0;
            const uint64_t c2 = e - b > 2 ? uint64_t(obs_cam[b[2]]) & 0x1FFFFF : c1;
            k = (c0 << 42) | (c1 << 21) | c2;
        }
        key[p] = k;
    }
    T.pt_order.resize(static_cast<size_t>(n_pts));
    std::iota(T.pt_order.begin(), T.pt_order.end(), 0);
    std::stable_sort(T.pt_order.begin(), T.pt_order.end(), [&](int32_t a, int32_t b) { return key[a] < key[b]; });
    T.pt_start.assign(size_t(n_pts) + 1, 0);
    T.obs_perm.resize(static_cast<size_t>(n_obs));
    T.obs_lcam.assign(static_cast<size_t>(n_obs), 0);
    for (int d = 0; d < n_pts; ++d) {
        const int p = T.pt_order[d];
        const int k = start[size_t(p) + 1] - start[p];
        T.pt_start[size_t(d) + 1] = T.pt_start[d] + k;
        std::copy(sorted_obs.begin() + start[p], sorted_obs.begin() + start[size_t(p) + 1], T.obs_perm.begin() + T.pt_start[d]);
    }
    // ---- greedy tiles over the device order
    const int w_cap = std::max(32, std::min(prm.w_cap, kMaxWCap));   // split tiles hold up to 32 cameras
    std::vector<int32_t> stamp(static_cast<size_t>(std::max(1, n_cams)), -1), lidx(static_cast<size_t>(std::max(1, n_cams)), 0);
    T.tiles.clear(); T.tile_cams.clear(); T.tile_marks.clear();
    T.w_max = 0;
    auto close_tile = [&](int d0, int d1, std::vector<int32_t>& cams) {
        if (d1 <= d0) return;
        std::sort(cams.begin(), cams.end());
        Tile t{};
        t.pt_begin = d0; t.pt_end = d1;
        t.cam_begin = static_cast<int32_t>(T.tile_cams.size());
        t.w = static_cast<int32_t>(cams.size());
        t.slot_begin = static_cast<int32_t>(T.tile_marks.size());
        t.flags = kTilePrimary;
        for (size_t i = 0; i < cams.size(); ++i) lidx[cams[i]] = static_cast<int32_t>(i);
        T.tile_cams.insert(T.tile_cams.end(), cams.begin(), cams.end());
        const size_t nb = size_t(t.w) * (t.w + 1) / 2;
        T.tile_marks.resize(T.tile_marks.size() + nb, 0);
        int32_t* marks = T.tile_marks.data() + t.slot_begin;
        for (int d = d0; d < d1; ++d)
            for (int a = T.pt_start[d]; a < T.pt_start[size_t(d) + 1]; ++a) {
                const int ca = obs_cam[T.obs_perm[a]];
                const int la = lidx[ca];
                T.obs_lcam[a] = static_cast<uint8_t>(la);
                if (cam_free[ca] < 0) continue;
                for (int b = a; b < T.pt_start[size_t(d) + 1]; ++b) {
                    const int cb = obs_cam[T.obs_perm[b]];
                    if (cam_free[cb] >= 0) marks[tri_index(la, lidx[cb])] = 1;
                }
            }
        T.w_max = std::max(T.w_max, int(t.w));
        T.tiles.push_back(t);
        cams.clear();
    };
    // a point with more than 32 observations: one tile per pair of 16-observation groups
    auto split_point = [&](int d) {
        const int beg = T.pt_start[d], k = T.pt_start[size_t(d) + 1] - beg;
        const int ng = (k + 15) / 16;
        for (int gi = 0; gi < ng; ++gi)
            for (int gj = gi; gj < ng; ++gj) {
                Tile t{};
                t.pt_begin = d; t.pt_end = d + 1;
                t.cam_begin = static_cast<int32_t>(T.tile_cams.size());
                t.sub_a0 = gi * 16; t.sub_a1 = std::min(k, gi * 16 + 16);
                if (gj != gi) { t.sub_b0 = gj * 16; t.sub_b1 = std::min(k, gj * 16 + 16); }
                const int na = t.sub_a1 - t.sub_a0, nb_ = t.sub_b1 - t.sub_b0;
                t.w = na + nb_;
                t.slot_begin = static_cast<int32_t>(T.tile_marks.size());
                t.flags = kTileSplit | ((gi == 0 && gj == 0) ? kTilePrimary : 0);
                for (int i = 0; i < na; ++i) T.tile_cams.push_back(obs_cam[T.obs_perm[beg + t.sub_a0 + i]]);
                for (int i = 0; i < nb_; ++i) T.tile_cams.push_back(obs_cam[T.obs_perm[beg + t.sub_b0 + i]]);
                const size_t nb = size_t(t.w) * (t.w + 1) / 2;
                T.tile_marks.resize(T.tile_marks.size() + nb, 0);
                int32_t* marks = T.tile_marks.data() + t.slot_begin;
                const int32_t* lc = T.tile_cams.data() + t.cam_begin;
                if (gj == gi) {
                    for (int x = 0; x < na; ++x)
                        for (int y = x; y < na; ++y)
                            if (cam_free[lc[x]] >= 0 && cam_free[lc[y]] >= 0) marks[tri_index(x, y)] = 1;
                } else {
                    for (int x = 0; x < na; ++x)
                        for (int y = na; y < na + nb_; ++y)
                            if (cam_free[lc[x]] >= 0 && cam_free[lc[y]] >= 0) marks[tri_index(x, y)] = 1;
                }
                T.w_max = std::max(T.w_max, int(t.w));
                T.tiles.push_back(t);
            }
    };
    std::vector<int32_t> cams;
    int d0 = 0, tile_id = 0;
    long long work = 0;
    for (int d = 0; d < n_pts; ++d) {
        const int beg = T.pt_start[d], k = T.pt_start[size_t(d) + 1] - beg;
        if (k == 0) { close_tile(d0, d, cams); d0 = n_pts; break; }     // points without observations are sorted last
        if (k > 32 || k > w_cap) {
            close_tile(d0, d, cams);
            ++tile_id; work = 0; d0 = d + 1;
            split_point(d);
            continue;
        }
        int fresh = 0;
        for (int a = beg; a < beg + k; ++a)
            if (stamp[obs_cam[T.obs_perm[a]]] != tile_id) ++fresh;
        const long long pw = static_cast<long long>(k) * (k + 1) / 2;
        if (d > d0 && (int(cams.size()) + fresh > w_cap || d - d0 >= prm.max_pts || work + pw > prm.max_work)) {
            close_tile(d0, d, cams);
            ++tile_id; work = 0; d0 = d;
        }
        for (int a = beg; a < beg + k; ++a) {
            const int c = obs_cam[T.obs_perm[a]];
            if (stamp[c] != tile_id) { stamp[c] = tile_id; cams.push_back(c); }
        }
        work += pw;
    }
    if (d0 < n_pts) close_tile(d0, n_pts, cams);
    return true;
}

// Marks of all tiles -> presence bitmap [n_free * n_free] (row fa <= column fb), 1 byte per block.
inline void mark_blocks(const Tiling& T, const int32_t* cam_free, int n_free, std::vector<uint8_t>& present) {
    present.assign(size_t(n_free) * n_free, 0);
    for (const Tile& t : T.tiles) {
        const int32_t* lc = T.tile_cams.data() + t.cam_begin;
        const int32_t* marks = T.tile_marks.data() + t.slot_begin;
        for (int lb = 0; lb < t.w; ++lb)
            for (int la = 0; la <= lb; ++la)
                if (marks[tri_index(la, lb)]) present[size_t(cam_free[lc[la]]) * n_free + cam_free[lc[lb]]] = 1;
    }
    for (int f = 0; f < n_free; ++f) present[size_t(f) * n_free + f] = 1;      // every diagonal block exists (damping)
}

// Presence bitmap (already merged over the ranks) -> block CSR and the per-tile slot tables.
inline void assign_slots(Tiling& T, const int32_t* cam_free, int n_free, const std::vector<uint8_t>& present) {
    T.blk_rowptr.assign(size_t(n_free) + 1, 0);
    T.blk_row.clear(); T.blk_col.clear();
    for (int fa = 0; fa < n_free; ++fa) {
        for (int fb = fa; fb < n_free; ++fb)
            if (present[size_t(fa) * n_free + fb]) { T.blk_row.push_back(fa); T.blk_col.push_back(fb); }
        T.blk_rowptr[size_t(fa) + 1] = static_cast<int32_t>(T.blk_col.size());
    }
    T.tile_slots.assign(T.tile_marks.size(), -1);
    for (const Tile& t : T.tiles) {
        const int32_t* lc = T.tile_cams.data() + t.cam_begin;
        for (int lb = 0; lb < t.w; ++lb)
            for (int la = 0; la <= lb; ++la) {
                const int i = t.slot_begin + tri_index(la, lb);
                if (!T.tile_marks[i]) continue;
                const int fa = cam_free[lc[la]], fb = cam_free[lc[lb]];
                const int32_t* b = T.blk_col.data() + T.blk_rowptr[fa];
                const int32_t* e = T.blk_col.data() + T.blk_rowptr[size_t(fa) + 1];
                T.tile_slots[i] = static_cast<int32_t>(std::lower_bound(b, e, fb) - T.blk_col.data());
            }
    }
}

}  // namespace ba
}  // namespace msfm
