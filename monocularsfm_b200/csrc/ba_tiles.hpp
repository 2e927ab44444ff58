// Host-side structure analysis of a bundle-adjustment problem (once per sparsity pattern): device order of the points
// and observations, tiles, and the block structure of the reduced camera system.  Plain C++ (no CUDA) so that the CPU test
// suite can check its invariants (tests/test_ba_tiles.py).
//
// What it replaces in the reference: Ceres' Program / block-structure detection for SchurEliminator behind
// ceres::Solve (src/Optimizer/CeresBundleOptimizer.cpp:293) — e-blocks = points, f-blocks = cameras.
#pragma once
#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstdlib>
#include <functional>
#include <numeric>
#include <thread>
#include <vector>

#include "ba_types.cuh"

namespace msfm {
namespace ba {

struct TilingParams {
    int max_pts = kTilePts;    // max points per normal tile (<= kTilePts)
    int max_obs = kTileObs;    // max observations per normal tile (<= kTileObs)
    int max_items = kTileItems;
    int chunk_pts = 16384;     // device points per independently tiled chunk of normal tiles (a constant of the analysis, not
    int chunk_long = 1024;     // of the machine: the result does not depend on the thread count); long tracks per chunk of item tiles
    int max_run = 16;          // units per run (<= 255).  A run is ONE work item per 32 camera pairs, i.e. the load-balance grain of
                               // the pair phase: 255 -> 16 took 4.5 % off the kernel at configs[4] (profiles/r02_k2_v7_ab.txt)
};

struct Tiling {
    std::vector<int32_t> pt_order;      // device point -> caller's point
    std::vector<int32_t> pt_start;      // [n_pts + 1] CSR over device-ordered observations
    std::vector<int32_t> obs_perm;      // device observation -> caller's observation
    std::vector<uint8_t> obs_lcam;      // device observation -> local camera of its (normal) tile
    std::vector<uint8_t> obs_lpt;       // device observation -> index of its point inside its (normal) tile
    int first_long = 0, n_long = 0;     // device points [first_long, first_long + n_long): more than 32 observations
    std::vector<Tile> tiles;
    std::vector<Item> items;
    std::vector<uint32_t> runs;         // per tile, the work items of the pair phase: first unit of a run of units with identical
                                        // camera lists (16 bits) | units in the run (8 bits) | round of 32 camera pairs (8 bits)
    std::vector<int32_t> tile_cams;
    std::vector<int32_t> tile_free;     // parallel to tile_cams: cam_free of the local camera (assign_slots)
    std::vector<int32_t> tile_marks;    // per tile, tri(w) entries: 1 if some point of the tile couples the two local cameras
    std::vector<int32_t> tile_slots;    // same shape: global block slot or -1 (assign_slots)
    int w_max = 0;
    // block structure (upper block triangle incl. the diagonal), sorted by (row, col)
    std::vector<int32_t> blk_rowptr, blk_row, blk_col;
};

inline int tri_index(int la, int lb) { return lb * (lb + 1) / 2 + la; }   // la <= lb

constexpr int kMaxHostThreads = 16;
// host threads of the structure analysis: hardware_concurrency, at most kMaxHostThreads; MSFM_HOST_THREADS=<n> lowers it
// (the result does not depend on it: tests/test_ba_tiles.py::test_analysis_does_not_depend_on_the_thread_count)
inline int host_threads() {
    int nt = static_cast<int>(std::thread::hardware_concurrency());
    nt = std::max(1, std::min(kMaxHostThreads, nt));
    if (const char* e = std::getenv("MSFM_HOST_THREADS")) {
        const int v = std::atoi(e);
        if (v > 0) nt = std::min(nt, v);
    }
    return nt;
}

// f(begin, end) over [0, n) on up to kMaxHostThreads host threads (the analysis of a 5 M-observation problem is ~0.5 s on one)
template <class F>
inline void parallel_ranges(int n, F f) {
    const int nt = host_threads();
    if (n < 20000 || nt == 1) { f(0, n); return; }
    std::vector<std::thread> th;
    const int step = (n + nt - 1) / nt;
    for (int t = 0; t < nt; ++t) {
        const int b = t * step, e = std::min(n, b + step);
        if (b < e) th.emplace_back([=] { f(b, e); });
    }
    for (auto& x : th) x.join();
}

// Stable order of v under a strict total order `less`, on the host threads: sorted chunks, then rounds of pairwise merges.
template <class T, class Less>
inline void parallel_sort(std::vector<T>& v, Less less) {
    const int nt = host_threads();
    const size_t n = v.size();
    if (n < 50000 || nt == 1) { std::sort(v.begin(), v.end(), less); return; }
    int parts = 1;
    while (parts * 2 <= nt) parts *= 2;
    std::vector<size_t> cut(size_t(parts) + 1);
    for (int i = 0; i <= parts; ++i) cut[i] = n * size_t(i) / size_t(parts);
    {
        std::vector<std::thread> th;
        for (int i = 0; i < parts; ++i) th.emplace_back([&, i] { std::sort(v.begin() + cut[i], v.begin() + cut[i + 1], less); });
        for (auto& x : th) x.join();
    }
    for (int width = 1; width < parts; width *= 2) {
        std::vector<std::thread> th;
        for (int i = 0; i + width < parts; i += 2 * width)
            th.emplace_back([&, i, width] {
                std::inplace_merge(v.begin() + cut[i], v.begin() + cut[i + width], v.begin() + cut[std::min(parts, i + 2 * width)], less);
            });
        for (auto& x : th) x.join();
    }
}

// f(chunk) for chunk in [0, n_chunks), chunks handed out dynamically to up to 8 host threads
template <class F>
inline void parallel_chunks(int n_chunks, F f) {
    const int nt = std::max(1, std::min(host_threads(), n_chunks));
    if (nt <= 1) { for (int c = 0; c < n_chunks; ++c) f(c); return; }
    std::atomic<int> next{0};
    std::vector<std::thread> th;
    for (int t = 0; t < nt; ++t)
        th.emplace_back([&] { for (int c = next.fetch_add(1); c < n_chunks; c = next.fetch_add(1)) f(c); });
    for (auto& x : th) x.join();
}

// Tiles of one chunk of the device order (the chunks are tiled independently on the host threads and concatenated in order;
// the chunk size is a constant, so the result does not depend on the number of threads).  Offsets are chunk-local.
struct TileChunk {
    std::vector<Tile> tiles;
    std::vector<Item> items;
    std::vector<uint32_t> runs;
    std::vector<int32_t> tile_cams, tile_marks;
    int w_max = 0;
};

// Device order + tiles + per-tile coupling marks.  obs_pt must be non-decreasing (validated by the caller).
// Returns false if a point is observed twice by one camera.
inline bool build_tiling(int n_cams, int n_pts, int n_obs, const int32_t* obs_cam, const int32_t* obs_pt,
                         const int32_t* cam_free, const TilingParams& prm, Tiling& T) {
    std::vector<int32_t> start(size_t(n_pts) + 1, 0);
    for (int i = 0; i < n_obs; ++i) start[size_t(obs_pt[i]) + 1] += 1;
    for (int p = 0; p < n_pts; ++p) start[size_t(p) + 1] += start[p];
    // observations of every point sorted by camera; cs = the sorted camera lists, contiguous (the comparisons below read them).
    // Device order: class (key bit 62: long tracks behind the others, points without observations last) and the three smallest
    // cameras — neighbours share cameras (few cameras per tile) — then track length and a hash of the whole camera list, so
    // that points with IDENTICAL lists end up adjacent (runs: their 6x6 products are summed in registers before they touch the
    // shared-memory accumulator; run detection compares the lists themselves, a hash collision only costs adjacency).
    // All-scalar keys in one contiguous array: no list walks in the sort.
    struct SortKey { uint64_t key, hash; int32_t k, idx; };
    std::vector<int32_t> sorted_obs(static_cast<size_t>(n_obs)), cs(static_cast<size_t>(n_obs));
    std::vector<SortKey> sk(static_cast<size_t>(n_pts));
    std::atomic<int> dup{0};
    parallel_ranges(n_pts, [&](int p0, int p1) {
        for (int p = p0; p < p1; ++p) {
            int32_t* b = sorted_obs.data() + start[p];
            int32_t* e = sorted_obs.data() + start[size_t(p) + 1];
            std::iota(b, e, start[p]);
            bool sorted = true;
            for (int32_t* q = b; q + 1 < e && sorted; ++q) sorted = obs_cam[q[0]] < obs_cam[q[1]];
            if (!sorted) {
                std::sort(b, e, [&](int32_t x, int32_t y) { return obs_cam[x] < obs_cam[y]; });
                for (int32_t* q = b; q + 1 < e; ++q)
                    if (obs_cam[q[0]] == obs_cam[q[1]]) dup.store(1, std::memory_order_relaxed);
            }
            int32_t* c = cs.data() + start[p];
            uint64_t h = 1469598103934665603ull;
            for (int32_t* q = b; q < e; ++q) {
                const int32_t cam = obs_cam[*q];
                *c++ = cam;
                h ^= static_cast<uint64_t>(static_cast<uint32_t>(cam)); h *= 1099511628211ull;
            }
            uint64_t k = ~uint64_t(0);
            if (e > b) {
                const int32_t* cc = cs.data() + start[p];
                const uint64_t c0 = uint64_t(cc[0]) & 0xFFFFF;
                const uint64_t c1 = e - b > 1 ? uint64_t(cc[1]) & 0xFFFFF : c0;
                const uint64_t c2 = e - b > 2 ? uint64_t(cc[2]) & 0xFFFFF : c1;
                k = (uint64_t(e - b > kLongTrack ? 1 : 0) << 62) | (c0 << 40) | (c1 << 20) | c2;
            }
            sk[p] = SortKey{k, h, static_cast<int32_t>(e - b), p};
        }
    });
    if (dup.load()) return false;
    parallel_sort(sk, [](const SortKey& a, const SortKey& b) {
        if (a.key != b.key) return a.key < b.key;
        if (a.k != b.k) return a.k < b.k;
        if (a.hash != b.hash) return a.hash < b.hash;
        return a.idx < b.idx;
    });
    T.pt_order.resize(static_cast<size_t>(n_pts));
    T.pt_start.assign(size_t(n_pts) + 1, 0);
    int first_long = n_pts, end_long = n_pts;          // classes are sorted: normal | long | unobserved
    for (int d = 0; d < n_pts; ++d) {
        T.pt_order[d] = sk[d].idx;
        T.pt_start[size_t(d) + 1] = T.pt_start[d] + sk[d].k;
        if (first_long == n_pts && (sk[d].k == 0 || sk[d].k > kLongTrack)) first_long = d;
        if (end_long == n_pts && sk[d].k == 0) end_long = d;
    }
    sk.clear(); sk.shrink_to_fit();
    T.obs_perm.resize(static_cast<size_t>(n_obs));
    T.obs_lcam.assign(static_cast<size_t>(n_obs), 0);
    T.obs_lpt.assign(static_cast<size_t>(n_obs), 0);
    std::vector<int32_t> dev_cam(static_cast<size_t>(n_obs));       // camera of every observation in device order
    parallel_ranges(n_pts, [&](int d0, int d1) {
        for (int d = d0; d < d1; ++d) {
            const int p = T.pt_order[d];
            std::copy(sorted_obs.begin() + start[p], sorted_obs.begin() + start[size_t(p) + 1], T.obs_perm.begin() + T.pt_start[d]);
            std::copy(cs.begin() + start[p], cs.begin() + start[size_t(p) + 1], dev_cam.begin() + T.pt_start[d]);
        }
    });
    const int w_cap = kTileCams;
    const int max_pts = std::max(1, std::min(prm.max_pts, kTilePts)), max_obs = std::max(32, std::min(prm.max_obs, kTileObs));
    const int max_items = std::max(1, std::min(prm.max_items, kTileItems));
    const int max_run = std::max(1, std::min(prm.max_run, 255));
    const int kChunkPoints = std::max(1, prm.chunk_pts), kChunkLong = std::max(1, prm.chunk_long);
    const int n_cam_slots = std::max(1, n_cams);
    auto cam_of = [&](int dev_obs) { return dev_cam[dev_obs]; };

    // ---- normal tiles: greedy over the device order, chunk by chunk
    const int n_chunks_a = (first_long + kChunkPoints - 1) / kChunkPoints;
    std::vector<TileChunk> chunk_a(static_cast<size_t>(n_chunks_a));
    parallel_chunks(n_chunks_a, [&](int ci) {
        TileChunk& C = chunk_a[ci];
        std::vector<int32_t> stamp(static_cast<size_t>(n_cam_slots), -1), lidx(static_cast<size_t>(n_cam_slots), 0);
        std::vector<int32_t> cams;
        auto close_tile = [&](int d0, int d1) {
            if (d1 <= d0) return;
            std::sort(cams.begin(), cams.end());
            Tile t{};
            t.begin = d0; t.end = d1;
            t.obs_begin = T.pt_start[d0]; t.n_obs = T.pt_start[d1] - T.pt_start[d0];
            t.cam_begin = static_cast<int32_t>(C.tile_cams.size());
            t.w = static_cast<int32_t>(cams.size());
            t.slot_begin = static_cast<int32_t>(C.tile_marks.size());
            for (size_t i = 0; i < cams.size(); ++i) lidx[cams[i]] = static_cast<int32_t>(i);
            C.tile_cams.insert(C.tile_cams.end(), cams.begin(), cams.end());
            C.tile_marks.resize(C.tile_marks.size() + size_t(t.w) * (t.w + 1) / 2, 0);
            int32_t* marks = C.tile_marks.data() + t.slot_begin;
            for (int a = T.pt_start[d0]; a < T.pt_start[d1]; ++a) T.obs_lcam[a] = static_cast<uint8_t>(lidx[cam_of(a)]);
            for (int d = d0; d < d1; ++d)
                for (int a = T.pt_start[d]; a < T.pt_start[size_t(d) + 1]; ++a) T.obs_lpt[a] = static_cast<uint8_t>(d - d0);
            // runs of consecutive points with identical camera lists (at most max_run points each), one work item per 32 camera
            // pairs; the coupling marks are those of the run's first point
            t.run_begin = static_cast<int32_t>(C.runs.size());
            for (int d = d0; d < d1;) {
                int e = d + 1;
                const int kd = T.pt_start[size_t(d) + 1] - T.pt_start[d];
                const int32_t* cd = dev_cam.data() + T.pt_start[d];
                while (e < d1 && e - d < max_run && T.pt_start[size_t(e) + 1] - T.pt_start[e] == kd &&
                       std::equal(cd, cd + kd, dev_cam.data() + T.pt_start[e])) ++e;
                for (int a = 0; a < kd; ++a) {
                    if (cam_free[cd[a]] < 0) continue;
                    const int la = lidx[cd[a]];
                    for (int b = a; b < kd; ++b)
                        if (cam_free[cd[b]] >= 0) marks[tri_index(la, lidx[cd[b]])] = 1;
                }
                const int rounds = (kd * (kd - 1) / 2 + 31) / 32;
                for (int r = 0; r < rounds; ++r)
                    C.runs.push_back(static_cast<uint32_t>(d - d0) | (static_cast<uint32_t>(e - d) << 16) | (static_cast<uint32_t>(r) << 24));
                d = e;
            }
            t.n_runs = static_cast<int32_t>(C.runs.size()) - t.run_begin;
            C.w_max = std::max(C.w_max, int(t.w));
            C.tiles.push_back(t);
            cams.clear();
        };
        const int c0 = ci * kChunkPoints, c1 = std::min(first_long, c0 + kChunkPoints);
        int d0 = c0, tile_id = 0, tile_obs = 0;
        for (int d = c0; d < c1; ++d) {
            const int beg = T.pt_start[d], k = T.pt_start[size_t(d) + 1] - beg;
            int fresh = 0;
            for (int a = beg; a < beg + k; ++a)
                if (stamp[cam_of(a)] != tile_id) ++fresh;
            if (d > d0 && (int(cams.size()) + fresh > w_cap || d - d0 >= max_pts || tile_obs + k > max_obs)) {
                close_tile(d0, d);
                ++tile_id; tile_obs = 0; d0 = d;
            }
            for (int a = beg; a < beg + k; ++a) {
                const int c = cam_of(a);
                if (stamp[c] != tile_id) { stamp[c] = tile_id; cams.push_back(c); }
            }
            tile_obs += k;
        }
        close_tile(d0, c1);
    });

    // ---- item tiles: long tracks cut into groups of 16 observations; one open tile per (group A, group B) index pair, so
    //      that the items of neighbouring long points (nearly the same cameras) are packed together
    T.first_long = first_long;
    T.n_long = end_long - first_long;
    const int n_chunks_b = (T.n_long + kChunkLong - 1) / kChunkLong;
    std::vector<TileChunk> chunk_b(static_cast<size_t>(n_chunks_b));
    parallel_chunks(n_chunks_b, [&](int ci) {
        TileChunk& C = chunk_b[ci];
        std::vector<int32_t> lidx(static_cast<size_t>(n_cam_slots), 0);
        struct Open { std::vector<int32_t> cams; std::vector<Item> items; };
        std::vector<Open> open;                    // index gi * ng_max + gj, grown on demand
        int ng_max = 0;
        auto close_items = [&](Open& o) {
            if (o.items.empty()) return;
            Tile t{};
            t.flags = kTileSplit;
            t.begin = static_cast<int32_t>(C.items.size());
            t.end = t.begin + static_cast<int32_t>(o.items.size());
            t.cam_begin = static_cast<int32_t>(C.tile_cams.size());
            t.w = static_cast<int32_t>(o.cams.size());
            t.slot_begin = static_cast<int32_t>(C.tile_marks.size());
            for (size_t i = 0; i < o.cams.size(); ++i) lidx[o.cams[i]] = static_cast<int32_t>(i);       // o.cams is kept sorted
            C.tile_cams.insert(C.tile_cams.end(), o.cams.begin(), o.cams.end());
            C.tile_marks.resize(C.tile_marks.size() + size_t(t.w) * (t.w + 1) / 2, 0);
            int32_t* marks = C.tile_marks.data() + t.slot_begin;
            for (Item& it : o.items) {
                const int beg = T.pt_start[it.d];
                const int na = it.a1 - it.a0, nb = it.b1 - it.b0;
                int cam_l[32];
                for (int l = 0; l < na + nb; ++l) {
                    cam_l[l] = cam_of(beg + (l < na ? it.a0 + l : it.b0 + l - na));
                    it.lc[l] = static_cast<uint8_t>(lidx[cam_l[l]]);
                }
                for (int x = 0; x < na; ++x) {
                    if (cam_free[cam_l[x]] < 0) continue;
                    if (nb == 0) {
                        for (int y = x; y < na; ++y)
                            if (cam_free[cam_l[y]] >= 0) marks[tri_index(it.lc[x], it.lc[y])] = 1;
                    } else {
                        for (int y = na; y < na + nb; ++y)
                            if (cam_free[cam_l[y]] >= 0) marks[tri_index(it.lc[x], it.lc[y])] = 1;
                    }
                }
                C.items.push_back(it);
            }
            t.run_begin = static_cast<int32_t>(C.runs.size());
            for (int u = 0; u < t.end - t.begin; ++u) {                          // an item is a run of one; a work item per 32 pairs
                const Item& it = C.items[size_t(t.begin) + u];
                const int na = it.a1 - it.a0, nb = it.b1 - it.b0;
                const int rounds = ((nb ? na * nb : na * (na - 1) / 2) + 31) / 32;
                for (int r = 0; r < rounds; ++r) C.runs.push_back(static_cast<uint32_t>(u) | (1u << 16) | (static_cast<uint32_t>(r) << 24));
            }
            t.n_runs = static_cast<int32_t>(C.runs.size()) - t.run_begin;
            C.w_max = std::max(C.w_max, int(t.w));
            C.tiles.push_back(t);
            o = Open();
        };
        const int c0 = first_long + ci * kChunkLong, c1 = std::min(end_long, c0 + kChunkLong);
        std::vector<int32_t> merged;
        for (int d = c0; d < c1; ++d) {
            const int beg = T.pt_start[d], k = T.pt_start[size_t(d) + 1] - beg;
            const int ng = (k + kItemGroup - 1) / kItemGroup;
            if (ng > ng_max) {                         // re-index the open tiles for the larger group count
                std::vector<Open> grown(size_t(ng) * ng);
                for (int gi = 0; gi < ng_max; ++gi)
                    for (int gj = gi; gj < ng_max; ++gj) grown[size_t(gi) * ng + gj] = std::move(open[size_t(gi) * ng_max + gj]);
                open.swap(grown);
                ng_max = ng;
            }
            for (int gi = 0; gi < ng; ++gi)
                for (int gj = gi; gj < ng; ++gj) {
                    Item it{};
                    it.d = d;
                    it.a0 = static_cast<uint16_t>(gi * kItemGroup); it.a1 = static_cast<uint16_t>(std::min(k, gi * kItemGroup + kItemGroup));
                    if (gj != gi) { it.b0 = static_cast<uint16_t>(gj * kItemGroup); it.b1 = static_cast<uint16_t>(std::min(k, gj * kItemGroup + kItemGroup)); }
                    const int na = it.a1 - it.a0, nb = it.b1 - it.b0;
                    Open& o = open[size_t(gi) * ng_max + gj];
                    // the item's cameras are ascending (group A, then group B); the open tile keeps its list sorted: one merge walk
                    int32_t ic[32];
                    for (int l = 0; l < na + nb; ++l) ic[l] = cam_of(beg + (l < na ? it.a0 + l : it.b0 + l - na));
                    auto count_fresh = [&](const std::vector<int32_t>& have) {
                        int fresh = 0;
                        size_t h = 0;
                        for (int l = 0; l < na + nb; ++l) {
                            while (h < have.size() && have[h] < ic[l]) ++h;
                            if (h == have.size() || have[h] != ic[l]) ++fresh;
                        }
                        return fresh;
                    };
                    if (!o.items.empty() && (int(o.cams.size()) + count_fresh(o.cams) > w_cap || int(o.items.size()) >= max_items)) close_items(o);
                    merged.resize(o.cams.size() + size_t(na + nb));
                    merged.resize(static_cast<size_t>(std::set_union(o.cams.begin(), o.cams.end(), ic, ic + na + nb, merged.begin()) - merged.begin()));
                    o.cams.swap(merged);
                    o.items.push_back(it);
                }
        }
        for (Open& o : open) close_items(o);
    });

    // ---- concatenate the chunks in order (normal tiles, then item tiles); chunk-local offsets become global
    size_t n_tiles = 0, n_items = 0, n_runs = 0, n_tc = 0, n_marks = 0;
    for (const std::vector<TileChunk>* cv : {&chunk_a, &chunk_b})
        for (const TileChunk& C : *cv) {
            n_tiles += C.tiles.size(); n_items += C.items.size(); n_runs += C.runs.size(); n_tc += C.tile_cams.size(); n_marks += C.tile_marks.size();
        }
    T.tiles.clear(); T.items.clear(); T.runs.clear(); T.tile_cams.clear(); T.tile_marks.clear();
    T.tiles.reserve(n_tiles); T.items.reserve(n_items); T.runs.reserve(n_runs); T.tile_cams.reserve(n_tc); T.tile_marks.reserve(n_marks);
    T.w_max = 0;
    for (std::vector<TileChunk>* cv : {&chunk_a, &chunk_b})
        for (TileChunk& C : *cv) {
            const int32_t item0 = static_cast<int32_t>(T.items.size()), run0 = static_cast<int32_t>(T.runs.size());
            const int32_t cam0 = static_cast<int32_t>(T.tile_cams.size()), mark0 = static_cast<int32_t>(T.tile_marks.size());
            for (Tile t : C.tiles) {
                if (t.flags & kTileSplit) { t.begin += item0; t.end += item0; }
                t.cam_begin += cam0; t.slot_begin += mark0; t.run_begin += run0;
                T.tiles.push_back(t);
            }
            T.items.insert(T.items.end(), C.items.begin(), C.items.end());
            T.runs.insert(T.runs.end(), C.runs.begin(), C.runs.end());
            T.tile_cams.insert(T.tile_cams.end(), C.tile_cams.begin(), C.tile_cams.end());
            T.tile_marks.insert(T.tile_marks.end(), C.tile_marks.begin(), C.tile_marks.end());
            T.w_max = std::max(T.w_max, C.w_max);
            C = TileChunk();
        }
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
    T.tile_free.resize(T.tile_cams.size());
    for (size_t i = 0; i < T.tile_cams.size(); ++i) T.tile_free[i] = cam_free[T.tile_cams[i]];
    const int n_tiles = static_cast<int>(T.tiles.size());
    parallel_ranges(n_tiles, [&](int t0, int t1) {
        for (int ti = t0; ti < t1; ++ti) {
            const Tile& t = T.tiles[ti];
            const int32_t* lc = T.tile_cams.data() + t.cam_begin;
            for (int la = 0; la < t.w; ++la) {
                const int fa = cam_free[lc[la]];
                if (fa < 0) continue;
                // local cameras and the block row are both ascending in the free-camera index: one forward walk
                const int32_t* b = T.blk_col.data() + T.blk_rowptr[fa];
                const int32_t* e = T.blk_col.data() + T.blk_rowptr[size_t(fa) + 1];
                for (int lb = la; lb < t.w; ++lb) {
                    const int i = t.slot_begin + tri_index(la, lb);
                    if (!T.tile_marks[i]) continue;
                    const int fb = cam_free[lc[lb]];
                    while (b < e && *b < fb) ++b;
                    T.tile_slots[i] = static_cast<int32_t>(b - T.blk_col.data());
                }
            }
        }
    });
}

}  // namespace ba
}  // namespace msfm
