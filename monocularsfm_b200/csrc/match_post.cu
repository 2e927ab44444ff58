// M-path support kernels around the tensor kernel (match_k1.cu):
//   desc_norm_key / desc_rank / desc_scatter / desc_groups
//                          raw [n][128] u8  ->  resident sorted-space layout of match_types.cuh
//   build_units_kernel     segment table    ->  unit table of one batch
//   resolve_rows_kernel    K1 row results   ->  ratio test; rescans the winner's 32-column group (best column +
//                                               in-group runner-up) when the row can still pass; defers tie /
//                                               sqrt-collapse rows
//   exact_rows_kernel      exact CUDA-core scan in OpenCV's (sqrtf(d2), index) order for deferred rows
//                          (and for every row in msfm_match_knn2_u8 mode 1)
//   count/scan/write       CrossCheck (FeatureUtils.cpp:281-310) + FilterMatchesByDistance (:208-218) +
//                          ordered compaction to the CSR output
// All integer arithmetic is exact; the float steps reproduce the reference's: distance = sqrtf((float)d2)
// (OpenCV batchDistL2_), ratio test `d0 < ratio * d1` in float (FeatureUtils.cpp:152).
#include "match_types.cuh"
#include "launch_count.hpp"
#include <cuda_runtime.h>

namespace msfm {

// ------------------------------------------------------------------------------------------------ upload formatting
// sort key of a descriptor: (bucket, squared norm, original index); bucket = parity * 3 + norm / kBucketSpan
__device__ __forceinline__ int bucket_of(int32_t nrm) { return (nrm & 1) * 3 + nrm / kBucketSpan; }

// one warp per descriptor: squared norm and 64-bit sort key; bucket population counts
__global__ void desc_norm_key_kernel(const uint8_t* __restrict__ raw, int n, int32_t* __restrict__ nrm_orig,
                                     unsigned long long* __restrict__ keys, int32_t* __restrict__ bucket_cnt) {
    const int wpb = blockDim.x >> 5, lane = threadIdx.x & 31;
    for (int j = blockIdx.x * wpb + (threadIdx.x >> 5); j < n; j += gridDim.x * wpb) {
        const uint32_t w = reinterpret_cast<const uint32_t*>(raw + static_cast<size_t>(j) * 128)[lane];
        int32_t s = static_cast<int32_t>(__dp4a(w, w, 0u));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) {
            nrm_orig[j] = s;
            const int b = bucket_of(s);
            keys[j] = (static_cast<unsigned long long>(b) << 56) | (static_cast<unsigned long long>(s) << 24) |
                      static_cast<unsigned long long>(j);
            atomicAdd(&bucket_cnt[b], 1);
        }
    }
}

// rank of every key among all keys: O(n^2) compares (n is a few thousand), tiled through shared memory and split over
// gridDim.y slices of the comparison range so that all SMs take part; partial ranks are summed with integer atomics
__global__ void __launch_bounds__(256)
desc_rank_kernel(const unsigned long long* __restrict__ keys, int n, const int32_t* __restrict__ bucket_cnt,
                 int32_t* __restrict__ rank_of, int32_t* __restrict__ used) {
    __shared__ unsigned long long tile[1024];
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned long long mine = j < n ? keys[j] : ~0ull;
    const int per = (n + gridDim.y - 1) / gridDim.y;
    const int lo = blockIdx.y * per, hi = min(n, lo + per);
    int rank = 0;
    for (int base = lo; base < hi; base += 1024) {
        for (int k = threadIdx.x; k < 1024; k += blockDim.x) tile[k] = (base + k < hi) ? keys[base + k] : ~0ull;
        __syncthreads();
        const int lim = min(1024, hi - base);
#pragma unroll 8
        for (int k = 0; k < lim; ++k) rank += tile[k] < mine ? 1 : 0;
        __syncthreads();
    }
    if (j == 0 && blockIdx.y == 0) {
        int tot = 0;
        for (int k = 0; k < kNumBuckets; ++k) tot += (bucket_cnt[k] + 31) & ~31;
        used[0] = tot;
    }
    if (j < n && rank) atomicAdd(&rank_of[j], rank);
}

// one warp per descriptor: write the swizzled row at its sorted position, its norm and original index
__global__ void desc_scatter_kernel(const uint8_t* __restrict__ raw, int n, const int32_t* __restrict__ rank_of,
                                    const int32_t* __restrict__ nrm_orig, const int32_t* __restrict__ bucket_cnt,
                                    uint8_t* __restrict__ sw, int32_t* __restrict__ nrm, int32_t* __restrict__ perm,
                                    int32_t* __restrict__ inv) {
    const int wpb = blockDim.x >> 5, lane = threadIdx.x & 31;
    for (int j = blockIdx.x * wpb + (threadIdx.x >> 5); j < n; j += gridDim.x * wpb) {
        // sorted-space position: rank + dead columns inserted before this key's bucket
        const int b = bucket_of(nrm_orig[j]);
        int pad = 0;
        for (int k = 0; k < b; ++k) pad += (32 - (bucket_cnt[k] & 31)) & 31;
        const int p = rank_of[j] + pad;
        const uint32_t w = reinterpret_cast<const uint32_t*>(raw + static_cast<size_t>(j) * 128)[lane];
        const int chunk = lane >> 2;                          // 16-byte chunk of this lane's word
        const int slot = ((chunk ^ (p & 7)) << 2) | (lane & 3);
        reinterpret_cast<uint32_t*>(sw + static_cast<size_t>(p) * 128)[slot] = w;
        if (lane == 0) { nrm[p] = nrm_orig[j]; perm[p] = j; inv[j] = p; }
    }
}

// one warp per 32-column group: C_g and the extension digits of every column
__global__ void desc_groups_kernel(const int32_t* __restrict__ nrm, int n_pad, int32_t* __restrict__ cg,
                                   uint8_t* __restrict__ ext) {
    const int wpb = blockDim.x >> 5, lane = threadIdx.x & 31;
    for (int g = blockIdx.x * wpb + (threadIdx.x >> 5); g < n_pad / 32; g += gridDim.x * wpb) {
        const int p = g * 32 + lane;
        const int32_t v = nrm[p];
        int32_t c = v;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) c = max(c, __shfl_xor_sync(0xffffffffu, c, o));
        if (lane == 0) cg[g] = c >= 0 ? c : kDeadCg;
        // digits of e = (C_g - ||d||^2) / 2 = b0 + 255 * (b1 + ... + b31); dead columns get e = 0
        uint32_t e = (v >= 0) ? static_cast<uint32_t>(c - v) >> 1 : 0u;
        uint32_t words[8];
        uint32_t t = e / 255u;
        const uint32_t b0 = e - t * 255u;
#pragma unroll
        for (int wi = 0; wi < 8; ++wi) {
            uint32_t word = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                uint32_t digit;
                if (wi == 0 && k == 0) digit = b0;
                else { digit = t < 255u ? t : 255u; t -= digit; }
                word |= digit << (8 * k);
            }
            words[wi] = word;
        }
        // column p = (4s+q)*256 + r  ->  bytes [32q, 32q+32) of row r of super-tile s (swizzled 16-byte chunks)
        const int s = p >> 10, q = (p >> 8) & 3, r = p & 255;
        uint8_t* rowp = ext + (static_cast<size_t>(s) * 256 + r) * 128;
        uint4* c0 = reinterpret_cast<uint4*>(rowp + (((2 * q) ^ (r & 7)) << 4));
        uint4* c1 = reinterpret_cast<uint4*>(rowp + (((2 * q + 1) ^ (r & 7)) << 4));
        *c0 = make_uint4(words[0], words[1], words[2], words[3]);
        *c1 = make_uint4(words[4], words[5], words[6], words[7]);
    }
}

// ------------------------------------------------------------------------------------------------
__global__ void build_units_kernel(const SegDev* __restrict__ segs, int nseg, int num_units, UnitDev* __restrict__ units) {
    const int u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= num_units) return;
    int lo = 0, hi = nseg - 1;                 // last segment with unit_base <= u
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (segs[mid].unit_base <= u) lo = mid; else hi = mid - 1;
    }
    const SegDev s = segs[lo];
    UnitDev o;
    o.q_slot = s.q_slot;
    o.t_slot = s.t_slot;
    o.row_block = u - s.unit_base;
    o.seg = lo;
    units[u] = o;
}

// ------------------------------------------------------------------------------------------------
// exact squared distance between swizzled rows (qi of image q, tj of image t)
__device__ __forceinline__ int32_t sqdist_rows(const uint8_t* __restrict__ qsw, int qi, const uint8_t* __restrict__ tsw,
                                               int tj) {
    const uint4* qa = reinterpret_cast<const uint4*>(qsw + static_cast<size_t>(qi) * 128);
    const uint4* tb = reinterpret_cast<const uint4*>(tsw + static_cast<size_t>(tj) * 128);
    const int qx = qi & 7, tx = tj & 7;
    uint32_t acc = 0;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const uint4 a = __ldg(qa + (c ^ qx));
        const uint4 b = __ldg(tb + (c ^ tx));
        uint32_t d;
        d = __vabsdiffu4(a.x, b.x); acc = __dp4a(d, d, acc);
        d = __vabsdiffu4(a.y, b.y); acc = __dp4a(d, d, acc);
        d = __vabsdiffu4(a.z, b.z); acc = __dp4a(d, d, acc);
        d = __vabsdiffu4(a.w, b.w); acc = __dp4a(d, d, acc);
    }
    return static_cast<int32_t>(acc);
}

// The same, but a caller that only needs min(limit, distance) lets the row stop early: after 64 and after 96 of the 128 bytes
// the partial sum is compared with `limit` and returned as it is once it has reached it (partial >= limit implies distance >=
// limit, so min(limit, .) is unchanged — the results stay bit-exact).  The rescan of resolve_rows_kernel is bound by L2
// bandwidth (74 % of the L2's peak, profiles/r02_resolve_ncu_summary.txt); a column that is not a neighbour usually passes the
// best distance outside its group well before its last bytes.
__device__ __forceinline__ int32_t sqdist_rows_bounded(const uint8_t* __restrict__ qsw, int qi, const uint8_t* __restrict__ tsw,
                                                       int tj, int32_t limit) {
    const uint4* qa = reinterpret_cast<const uint4*>(qsw + static_cast<size_t>(qi) * 128);
    const uint4* tb = reinterpret_cast<const uint4*>(tsw + static_cast<size_t>(tj) * 128);
    const int qx = qi & 7, tx = tj & 7;
    uint32_t acc = 0;
    auto chunk = [&](int c) {
        const uint4 a = __ldg(qa + (c ^ qx));
        const uint4 b = __ldg(tb + (c ^ tx));
        uint32_t d;
        d = __vabsdiffu4(a.x, b.x); acc = __dp4a(d, d, acc);
        d = __vabsdiffu4(a.y, b.y); acc = __dp4a(d, d, acc);
        d = __vabsdiffu4(a.z, b.z); acc = __dp4a(d, d, acc);
        d = __vabsdiffu4(a.w, b.w); acc = __dp4a(d, d, acc);
    };
    chunk(0); chunk(1); chunk(2); chunk(3);
    if (static_cast<int32_t>(acc) >= limit) return static_cast<int32_t>(acc);
    chunk(4); chunk(5);
    if (static_cast<int32_t>(acc) >= limit) return static_cast<int32_t>(acc);
    chunk(6); chunk(7);
    return static_cast<int32_t>(acc);
}

__device__ __forceinline__ float dist_of(int32_t d2) { return __fsqrt_rn(static_cast<float>(d2)); }
__device__ __forceinline__ bool ratio_pass(int32_t d1, int32_t d2, float ratio) {
    return dist_of(d1) < __fmul_rn(ratio, dist_of(d2));
}
__device__ __forceinline__ bool dist_filter_ok(int32_t d1, double max_distance) {
    // FilterMatchesByDistance: drop when (double)distance > max_distance
    return max_distance < 0.0 || !(static_cast<double>(dist_of(d1)) > max_distance);
}

// ------------------------------------------------------------------------------------------------
// One block of kUnitRows threads per unit; thread r owns sorted-space row r of the unit.
// Final per-row outputs are indexed by the ORIGINAL query index: o = (unit - row_block)*kUnitRows + perm_q[row]
//   m_j   original train index of the accepted match or -1     m_d1  exact d2 of the best column (kIntInf = none)
//   m_d2  exact d2 of the runner-up when it was computed, else an upper bound (kIntInf = none)
//   m_j0  [2o] best column regardless of the ratio test (knn2 API), [2o+1] runner-up column (exact path only) or -1
#ifndef MSFM_RESOLVE_MIN_CTAS
#define MSFM_RESOLVE_MIN_CTAS 12
#endif
__global__ void __launch_bounds__(kUnitRows, MSFM_RESOLVE_MIN_CTAS)
resolve_rows_kernel(const ImgDev* __restrict__ imgs, const UnitDev* __restrict__ units, int unit0, int num_units,
                    const int32_t* __restrict__ res_g, const int32_t* __restrict__ res_d1,
                    const int32_t* __restrict__ res_u, MatchOpts opt, int32_t* __restrict__ m_j,
                    int32_t* __restrict__ m_d1, int32_t* __restrict__ m_d2, int32_t* __restrict__ m_j0,
                    int32_t* __restrict__ exact_list, unsigned int* __restrict__ counters /*[0]=exact,[1]=rescans*/) {
    if (static_cast<int>(blockIdx.x) >= num_units) return;
    const int u = unit0 + blockIdx.x;
    const UnitDev unit = units[u];
    const ImgDev q = imgs[unit.q_slot];
    const ImgDev t = imgs[unit.t_slot];
    const int lane = threadIdx.x & 31;
    const int r = threadIdx.x;
    if (unit.row_block * kUnitRows >= (q.used ? __ldg(q.used) : 0)) return;     // all-dead rows (K1 skipped the unit too)
    const int qp = unit.row_block * kUnitRows + r;           // sorted-space row of the query image
    const size_t g = static_cast<size_t>(u) * kUnitRows + r; // K1 result slot
    const int qorig = q.perm[qp];                            // -1 for dead rows
    const bool valid = qorig >= 0;
    const size_t o = static_cast<size_t>(u - unit.row_block) * kUnitRows + (valid ? qorig : 0);

    int32_t g1 = -1, d1 = kIntInf, uu = kIntInf;
    if (valid) { g1 = res_g[g]; d1 = res_d1[g]; uu = res_u[g]; }
    const bool have = valid && g1 >= 0 && t.n >= 1;
    // rows that must be redone exactly: cross-group tie of the best distance, or float-sqrt collapse range
    const bool need_exact = have && t.n >= 2 && (uu == d1 || d1 >= kSqrtExactLimit);
    // the best column / runner-up only matter if the row passes against the upper bound uu (or the caller wants them)
    const bool need_rescan = have && !need_exact &&
                             (opt.exact_second || t.n < 2 || uu == kIntInf || ratio_pass(d1, uu, opt.ratio));

    int32_t d2 = uu;
    int32_t j1 = -1;
    // ---- warp-cooperative rescan of the winner's 32-column group: best column (lowest original index among equal
    //      distances, OpenCV's order) and the in-group runner-up distance
    unsigned todo = __ballot_sync(0xffffffffu, need_rescan);
    while (todo) {
        const int src = __ffs(todo) - 1;
        todo &= todo - 1;
        const int sg1 = __shfl_sync(0xffffffffu, g1, src);
        const int sqp = __shfl_sync(0xffffffffu, qp, src);
        // a column at or beyond the best distance outside the group (uu) changes neither the winner (its distance d1 < uu) nor
        // d2 = min(uu, in-group runner-up): it may stop early — unless the caller wants the exact second distance (knn2)
        const int32_t lim = opt.exact_second ? kIntInf : __shfl_sync(0xffffffffu, uu, src);
        const int col = sg1 * 32 + lane;                     // sorted-space column (< t.n_pad)
        const int corig = t.perm[col];
        int32_t dd = kIntInf;
        if (corig >= 0) dd = sqdist_rows_bounded(q.sw, sqp, t.sw, col, lim);
        // (dd, corig) lexicographic minimum over the warp: two REDUX instead of five shuffle / compare rounds
        const int32_t bd = __reduce_min_sync(0xffffffffu, dd);
        const int32_t bj = __reduce_min_sync(0xffffffffu, (corig >= 0 && dd == bd) ? corig : kIntInf);
        // runner-up inside the group: minimum over the lanes that are not the winner
        const int32_t second = __reduce_min_sync(0xffffffffu, (corig >= 0 && corig != bj) ? dd : kIntInf);
        if (lane == src) {
            j1 = bj;
            d2 = min(d2, second);
        }
    }
    const unsigned nres = __popc(__ballot_sync(0xffffffffu, need_rescan));
    if (lane == 0 && nres) atomicAdd(&counters[1], nres);

    if (need_exact) {
        const unsigned slot = atomicAdd(&counters[0], 1u);
        exact_list[slot] = static_cast<int32_t>(g);
    }
    if (valid && !need_exact) {
        int32_t mj = -1;
        if (j1 >= 0 && t.n >= 2 && d2 != kIntInf && ratio_pass(d1, d2, opt.ratio)) mj = j1;
        m_j[o] = mj;
        m_d1[o] = have ? d1 : kIntInf;
        m_d2[o] = (have && t.n >= 2) ? d2 : kIntInf;
        m_j0[2 * o] = j1;
        m_j0[2 * o + 1] = -1;      // the tensor path tracks the runner-up's distance, not its column
    }
}

// ------------------------------------------------------------------------------------------------
// Exact scan of listed rows over all train columns, OpenCV order = (sqrtf(d2), ORIGINAL column) lexicographic.
struct Top2 {
    int32_t d0, j0, d1, j1;
};
__device__ __forceinline__ bool lex_less(int32_t da, int32_t ja, int32_t db, int32_t jb) {
    const float fa = dist_of(da), fb = dist_of(db);
    return fa < fb || (fa == fb && ja < jb);
}
__device__ __forceinline__ void top2_insert(Top2& s, int32_t d, int32_t j) {
    if (j < 0) return;
    if (s.j1 >= 0 && !lex_less(d, j, s.d1, s.j1)) return;
    if (s.j0 < 0 || lex_less(d, j, s.d0, s.j0)) {
        s.d1 = s.d0; s.j1 = s.j0; s.d0 = d; s.j0 = j;
    } else {
        s.d1 = d; s.j1 = j;
    }
}

// One CTA per listed row (8 warps split the columns; a lone warp per row is latency-bound on ~260 dependent
// load rounds).
__global__ void __launch_bounds__(256)
exact_rows_kernel(const ImgDev* __restrict__ imgs, const UnitDev* __restrict__ units,
                  const int32_t* __restrict__ row_list, const unsigned int* __restrict__ row_count_dev,
                  int row_count_host /* >=0: use this instead of the device counter */, MatchOpts opt,
                  int32_t* __restrict__ m_j, int32_t* __restrict__ m_d1, int32_t* __restrict__ m_d2,
                  int32_t* __restrict__ m_j0) {
    const int nrows = row_count_host >= 0 ? row_count_host : static_cast<int>(*row_count_dev);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __shared__ Top2 part[8];
    for (int w = blockIdx.x; w < nrows; w += gridDim.x) {
        const size_t g = row_list ? static_cast<size_t>(row_list[w]) : static_cast<size_t>(w);
        const int u = static_cast<int>(g / kUnitRows);
        const UnitDev unit = units[u];
        const ImgDev q = imgs[unit.q_slot];
        const ImgDev t = imgs[unit.t_slot];
        const int qp = unit.row_block * kUnitRows + static_cast<int>(g % kUnitRows);
        const int qorig = (q.used && qp < __ldg(q.used)) ? q.perm[qp] : -1;      // block-uniform
        if (qorig < 0) continue;
        const size_t o = static_cast<size_t>(u - unit.row_block) * kUnitRows + qorig;
        const int tcols = t.used ? __ldg(t.used) : 0;
        Top2 s{kIntInf, -1, kIntInf, -1};
        for (int col = threadIdx.x; col < tcols; col += blockDim.x) {
            const int corig = t.perm[col];
            if (corig >= 0) top2_insert(s, sqdist_rows(q.sw, qp, t.sw, col), corig);
        }
#pragma unroll
        for (int ofs = 16; ofs > 0; ofs >>= 1) {
            Top2 other;
            other.d0 = __shfl_xor_sync(0xffffffffu, s.d0, ofs);
            other.j0 = __shfl_xor_sync(0xffffffffu, s.j0, ofs);
            other.d1 = __shfl_xor_sync(0xffffffffu, s.d1, ofs);
            other.j1 = __shfl_xor_sync(0xffffffffu, s.j1, ofs);
            top2_insert(s, other.d0, other.j0);
            top2_insert(s, other.d1, other.j1);
        }
        __syncthreads();                       // part[] of the previous row has been consumed
        if (lane == 0) part[warp] = s;
        __syncthreads();
        if (threadIdx.x == 0) {
            Top2 r = part[0];
            for (int k = 1; k < (blockDim.x >> 5); ++k) {
                top2_insert(r, part[k].d0, part[k].j0);
                top2_insert(r, part[k].d1, part[k].j1);
            }
            int32_t mj = -1;
            if (r.j0 >= 0 && r.j1 >= 0 && ratio_pass(r.d0, r.d1, opt.ratio)) mj = r.j0;
            m_j[o] = mj;
            m_d1[o] = r.j0 >= 0 ? r.d0 : kIntInf;
            m_d2[o] = r.j1 >= 0 ? r.d1 : kIntInf;
            m_j0[2 * o] = r.j0;
            m_j0[2 * o + 1] = r.j1;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// CrossCheck + distance filter predicate for query row i of pair segment(s).
struct PairView {
    int n1, n2;
    size_t base12, base21;   // global row base of the two directions (base21 unused without cross-check)
};
// segment table of a batch: [forward segments of all pairs][reverse segments of all pairs]
__device__ __forceinline__ PairView pair_view(const ImgDev* imgs, const SegDev* segs, int p, int npairs, int cross) {
    const SegDev s12 = segs[p];
    PairView v;
    v.n1 = imgs[s12.q_slot].n;
    v.n2 = imgs[s12.t_slot].n;
    v.base12 = static_cast<size_t>(s12.unit_base) * kUnitRows;
    v.base21 = cross ? static_cast<size_t>(segs[npairs + p].unit_base) * kUnitRows : 0;
    return v;
}
__device__ __forceinline__ bool keep_match(const PairView& v, int i, const int32_t* m_j, const int32_t* m_d1,
                                           const MatchOpts& opt, int32_t& j_out) {
    const int32_t j = m_j[v.base12 + i];
    j_out = j;
    if (j < 0) return false;
    if (opt.cross_check) {
        const int32_t rj = m_j[v.base21 + j];
        // reference CrossCheck: vis[train] (default-inserted 0 when the reverse direction has no match) == query
        const bool ok = (rj == i) || (opt.quirks && rj < 0 && i == 0);
        if (!ok) return false;
    }
    return dist_filter_ok(m_d1[v.base12 + i], opt.max_distance);
}

__global__ void __launch_bounds__(256)
count_matches_kernel(const ImgDev* __restrict__ imgs, const SegDev* __restrict__ segs, int npairs, MatchOpts opt,
                     const int32_t* __restrict__ m_j, const int32_t* __restrict__ m_d1, int32_t* __restrict__ counts) {
    const int p = blockIdx.x;
    if (p >= npairs) return;
    const PairView v = pair_view(imgs, segs, p, npairs, opt.cross_check);
    int c = 0;
    for (int i = threadIdx.x; i < v.n1; i += blockDim.x) {
        int32_t j;
        c += keep_match(v, i, m_j, m_d1, opt, j) ? 1 : 0;
    }
    __shared__ int sh[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int s = 0;
        for (int k = 0; k < (blockDim.x >> 5); ++k) s += sh[k];
        counts[p] = s;
    }
}

// Single-block exclusive scan of counts[0..n) into offsets[first .. first+n], continuing from *running_total.
// offsets are global (CSR over the whole call); running_total carries across batches on the device.
__global__ void __launch_bounds__(1024)
scan_counts_kernel(const int32_t* __restrict__ counts, int n, long long* __restrict__ offsets /* already offset to this batch */,
                   long long* __restrict__ running_total) {
    __shared__ long long warp_sums[32];
    __shared__ long long carry;
    if (threadIdx.x == 0) carry = *running_total;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int base = 0; base < n; base += blockDim.x) {
        const int i = base + threadIdx.x;
        long long v = (i < n) ? counts[i] : 0;
        long long incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const long long y = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += y;
        }
        if (lane == 31) warp_sums[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            long long ws = warp_sums[lane];
            long long wincl = ws;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const long long y = __shfl_up_sync(0xffffffffu, wincl, o);
                if (lane >= o) wincl += y;
            }
            warp_sums[lane] = wincl - ws;   // exclusive
        }
        __syncthreads();
        const long long excl = carry + warp_sums[warp] + incl - v;
        if (i < n) offsets[i] = excl;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        offsets[n] = carry;
        *running_total = carry;
    }
}

__global__ void __launch_bounds__(256)
write_matches_kernel(const ImgDev* __restrict__ imgs, const SegDev* __restrict__ segs, int npairs, MatchOpts opt,
                     const int32_t* __restrict__ m_j, const int32_t* __restrict__ m_d1,
                     const long long* __restrict__ offsets /* this batch */, long long capacity,
                     int32_t* __restrict__ out_matches, float* __restrict__ out_dist) {
    const int p = blockIdx.x;
    if (p >= npairs) return;
    const PairView v = pair_view(imgs, segs, p, npairs, opt.cross_check);
    const long long off = offsets[p];
    __shared__ int warp_cnt[8];
    __shared__ int chunk_base;
    if (threadIdx.x == 0) chunk_base = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int base = 0; base < v.n1; base += blockDim.x) {
        const int i = base + threadIdx.x;
        int32_t j = -1;
        const bool keep = (i < v.n1) && keep_match(v, i, m_j, m_d1, opt, j);
        const unsigned mask = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) warp_cnt[warp] = __popc(mask);
        __syncthreads();
        int pre = chunk_base;
        for (int k = 0; k < warp; ++k) pre += warp_cnt[k];
        if (keep) {
            const long long pos = off + pre + __popc(mask & ((1u << lane) - 1u));
            if (pos < capacity) {
                out_matches[2 * pos] = i;
                out_matches[2 * pos + 1] = j;
                if (out_dist) out_dist[pos] = dist_of(m_d1[v.base12 + i]);
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int s = 0;
            for (int k = 0; k < (blockDim.x >> 5); ++k) s += warp_cnt[k];
            chunk_base += s;
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// Cross-check needs the reverse direction only for train descriptors that some query row actually matched
// (FeatureUtils::CrossCheck looks up vis[m12.trainIdx] only, FeatureUtils.cpp:296-302).  After the forward pass the
// matched train rows of every pair are gathered into a compact temporary "query image"; the reverse pass then runs on
// ~n/10 rows instead of n.  Its perm[] holds the ORIGINAL train index, so the reverse results land where the full
// reverse pass would have put them and the downstream kernels do not change.
struct TempImgs {
    uint8_t* sw;        // [pairs][n_pad_t][128]
    int32_t* nrm;       // [pairs][n_pad_t]
    int32_t* perm;      // [pairs][n_pad_t]
    int32_t* used;      // [pairs]
    int32_t* flags;     // [pairs][n_pad_t]
    int32_t n_pad_t;    // rows reserved per pair (multiple of kUnitRows)
};

// table entries of the temporary query images: slot = first_temp_slot + pair
__global__ void setup_temp_imgs_kernel(ImgDev* __restrict__ imgs, int first_temp_slot, const SegDev* __restrict__ segs,
                                       int npairs, TempImgs T) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npairs) return;
    const int n2 = imgs[segs[p].t_slot].n;
    ImgDev e;
    e.sw = T.sw + static_cast<size_t>(p) * T.n_pad_t * 128;
    e.ext = nullptr; e.cg = nullptr; e.inv = nullptr;
    e.nrm = T.nrm + static_cast<size_t>(p) * T.n_pad_t;
    e.perm = T.perm + static_cast<size_t>(p) * T.n_pad_t;
    e.used = T.used + p;
    e.n = n2;
    e.n_pad = (n2 + 2 * kUnitRows - 1) / (2 * kUnitRows) * (2 * kUnitRows);     // unit pairs (match_k1.cu)
    imgs[first_temp_slot + p] = e;
}

// one CTA per pair: flag the train rows matched by the forward pass, compact them in ascending original index,
// copy (re-swizzle) their descriptor rows.  1024 threads: every phase is a short chain of dependent global loads per
// thread, so the kernel's duration is (rows per thread) x (load latency); the 256-thread version with one row per warp
// iteration spent 0.33 ms per launch on ~1000 sequential iterations per warp.
constexpr int kGatherThreads = 1024;
__global__ void __launch_bounds__(kGatherThreads)
gather_candidates_kernel(const ImgDev* __restrict__ imgs, const SegDev* __restrict__ segs, int npairs,
                         const int32_t* __restrict__ m_j, TempImgs T) {
    const int p = blockIdx.x;
    if (p >= npairs) return;
    const SegDev s12 = segs[p];
    const ImgDev img1 = imgs[s12.q_slot], img2 = imgs[s12.t_slot];
    const int n1 = img1.n, n2 = img2.n;
    int32_t* flags = T.flags + static_cast<size_t>(p) * T.n_pad_t;
    int32_t* t_nrm = T.nrm + static_cast<size_t>(p) * T.n_pad_t;
    int32_t* t_perm = T.perm + static_cast<size_t>(p) * T.n_pad_t;
    uint8_t* t_sw = T.sw + static_cast<size_t>(p) * T.n_pad_t * 128;
    const size_t base12 = static_cast<size_t>(s12.unit_base) * kUnitRows;
    constexpr int kWarps = kGatherThreads / 32;
    for (int j = threadIdx.x; j < n2; j += kGatherThreads) flags[j] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < n1; i += kGatherThreads) {
        const int32_t j = m_j[base12 + i];
        if (j >= 0) flags[j] = 1;
    }
    __syncthreads();
    // ordered compaction: flags[j] becomes the slot of row j (or -1)
    __shared__ int warp_cnt[kWarps];
    __shared__ int warp_pre[kWarps];
    __shared__ int chunk_base;
    if (threadIdx.x == 0) chunk_base = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int base = 0; base < n2; base += kGatherThreads) {
        const int j = base + threadIdx.x;
        const bool f = j < n2 && flags[j] != 0;
        const unsigned mask = __ballot_sync(0xffffffffu, f);
        if (lane == 0) warp_cnt[warp] = __popc(mask);
        __syncthreads();
        if (warp == 0) {                                       // exclusive scan of the 32 warp counts
            const int cnt = warp_cnt[lane], cb = chunk_base;
            int incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int y = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += y;
            }
            warp_pre[lane] = cb + incl - cnt;
            if (lane == 31) chunk_base = cb + incl;            // every lane read cb before the shuffles above
        }
        __syncthreads();
        if (j < n2) flags[j] = f ? warp_pre[warp] + __popc(mask & ((1u << lane) - 1u)) : -1;
    }
    __syncthreads();
    const int K = chunk_base;
    const int Kpad = (K + kUnitRows - 1) / kUnitRows * kUnitRows;
    if (threadIdx.x == 0) T.used[p] = Kpad;
    // copy rows: a warp takes 32 consecutive train rows, reads their slots with one coalesced load and moves the rows
    // that have one, four at a time (the loads of the four rows in flight together)
    for (int j0 = warp * 32; j0 < n2; j0 += kWarps * 32) {
        const int j = j0 + lane;
        const int slot = j < n2 ? flags[j] : -1;
        const int pos = slot >= 0 ? __ldg(img2.inv + j) : 0;
        const int nrm = slot >= 0 ? __ldg(img2.nrm + pos) : 0;
        if (slot >= 0) { t_nrm[slot] = nrm; t_perm[slot] = j; }
        unsigned todo = __ballot_sync(0xffffffffu, slot >= 0);
        while (todo) {
            int rs[4], rp[4];
            uint32_t w[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int src_lane = todo ? __ffs(todo) - 1 : -1;
                if (todo) todo &= todo - 1;
                rs[k] = src_lane >= 0 ? __shfl_sync(0xffffffffu, slot, src_lane) : -1;
                rp[k] = src_lane >= 0 ? __shfl_sync(0xffffffffu, pos, src_lane) : 0;
            }
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (rs[k] >= 0) w[k] = __ldg(reinterpret_cast<const uint32_t*>(img2.sw + static_cast<size_t>(rp[k]) * 128) + lane);
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (rs[k] >= 0) {
                    const int chunk = (lane >> 2) ^ (rp[k] & 7);     // logical 16-byte chunk held at this lane's source word
                    reinterpret_cast<uint32_t*>(t_sw + static_cast<size_t>(rs[k]) * 128)[((chunk ^ (rs[k] & 7)) << 2) | (lane & 3)] = w[k];
                }
        }
    }
    // dead padding rows of the last unit
    for (int k = K + warp; k < Kpad; k += kWarps) {
        reinterpret_cast<uint32_t*>(t_sw + static_cast<size_t>(k) * 128)[lane] = 0u;
        if (lane == 0) { t_nrm[k] = -1; t_perm[k] = -1; }
    }
}

// ------------------------------------------------------------------------------------------------ launchers
cudaError_t launch_setup_temp_imgs(ImgDev* imgs, int first_temp_slot, const SegDev* segs, int npairs, TempImgs T, cudaStream_t st) {
    if (npairs <= 0) return cudaSuccess;
    { setup_temp_imgs_kernel<<<(npairs + 127) / 128, 128, 0, st>>>(imgs, first_temp_slot, segs, npairs, T); MSFM_COUNT_LAUNCH(); }
    return cudaGetLastError();
}
cudaError_t launch_gather_candidates(const ImgDev* imgs, const SegDev* segs, int npairs, const int32_t* m_j, TempImgs T,
                                     cudaStream_t st) {
    if (npairs <= 0) return cudaSuccess;
    { gather_candidates_kernel<<<npairs, kGatherThreads, 0, st>>>(imgs, segs, npairs, m_j, T); MSFM_COUNT_LAUNCH(); }
    return cudaGetLastError();
}
// ---- float32 descriptors (what the reference's Database stores, src/Database/Database.cpp:174-199) -> uint8.
// The bridge documented in INTEGRATION.md: a set whose values are all integers in [0,255] (un-normalised SIFT) converts
// exactly; any other set (L1-root / L2 normalised, FeatureExtraction.cpp:260-281) is quantised as
// clamp(rint(512 v), 0, 255) with round-half-to-even.  flag[0] must be 1 on entry; mode 0 = decide per set, 1 = always quantise.
__global__ void desc_f32_check_kernel(const float* __restrict__ src, size_t count, int32_t* __restrict__ flag) {
    bool ok = true;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < count; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const float v = __ldg(src + i);
        ok = ok && (v >= 0.f && v <= 255.f && v == floorf(v));
    }
    if (!__all_sync(0xffffffffu, ok) && (threadIdx.x & 31) == 0) atomicAnd(flag, 0);
}
__global__ void desc_f32_quantize_kernel(const float* __restrict__ src, size_t count4, int mode, const int32_t* __restrict__ flag,
                                         uint32_t* __restrict__ dst) {
    const bool integral = mode == 0 && flag[0] != 0;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < count4; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(src) + i);
        const float in[4] = {v.x, v.y, v.z, v.w};
        uint32_t packed = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float q = integral ? in[k] : rintf(__fmul_rn(in[k], 512.0f));
            const float c = fminf(255.f, fmaxf(0.f, q));          // NaN -> 0, like std::max(0.f, NaN)
            packed |= static_cast<uint32_t>(static_cast<int>(c)) << (8 * k);
        }
        dst[i] = packed;
    }
}
// Extraction-time normalisation of raw SIFT rows (FeatureExtraction.cpp:143-160, 260-281), in place, one warp per row, with
// OpenCV's arithmetic: the norm accumulated in double (cv::norm), every element multiplied IN FLOAT by float(1 / norm)
// (Mat /= double is convertTo(alpha = 1 / norm), which scales 32F data in float), then — L1_ROOT — a correctly rounded sqrtf.
//   kind 1: L1RootNormalized   row /= |row|_1 ; sqrt(row)          kind 2: L2Normalized   row /= |row|_2
__global__ void desc_f32_normalize_kernel(float* __restrict__ data, int n, int kind) {
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    for (int row = blockIdx.x * wpb + (threadIdx.x >> 5); row < n; row += gridDim.x * wpb) {
        float4* p = reinterpret_cast<float4*>(data + static_cast<size_t>(row) * 128) + lane;
        const float4 v = *p;
        double s = kind == 1 ? static_cast<double>(fabsf(v.x)) + fabsf(v.y) + fabsf(v.z) + fabsf(v.w)
                             : static_cast<double>(v.x) * v.x + static_cast<double>(v.y) * v.y + static_cast<double>(v.z) * v.z +
                                   static_cast<double>(v.w) * v.w;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        const double norm = kind == 1 ? s : sqrt(s);
        const float a = static_cast<float>(1.0 / norm);
        float4 r = make_float4(__fmul_rn(v.x, a), __fmul_rn(v.y, a), __fmul_rn(v.z, a), __fmul_rn(v.w, a));
        if (kind == 1) r = make_float4(__fsqrt_rn(r.x), __fsqrt_rn(r.y), __fsqrt_rn(r.z), __fsqrt_rn(r.w));
        *p = r;
    }
}
cudaError_t launch_desc_normalize(float* data, int n, int kind, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    int grid = (n + 7) / 8;
    if (grid > 148 * 8) grid = 148 * 8;
    { desc_f32_normalize_kernel<<<grid, 256, 0, st>>>(data, n, kind); MSFM_COUNT_LAUNCH(); }
    return cudaGetLastError();
}

cudaError_t launch_desc_quantize(const float* src, int n, int mode, int32_t* flag, uint8_t* dst, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    const size_t count = static_cast<size_t>(n) * 128;
    cudaError_t e;
    const int32_t one = 1;
    if ((e = cudaMemcpyAsync(flag, &one, sizeof(one), cudaMemcpyHostToDevice, st)) != cudaSuccess) return e;
    int grid = static_cast<int>((count / 4 + 255) / 256);
    if (grid > 148 * 8) grid = 148 * 8;
    if (mode == 0) { desc_f32_check_kernel<<<grid, 256, 0, st>>>(src, count, flag); MSFM_COUNT_LAUNCH(); }
    { desc_f32_quantize_kernel<<<grid, 256, 0, st>>>(src, count / 4, mode, flag, reinterpret_cast<uint32_t*>(dst)); MSFM_COUNT_LAUNCH(); }
    return cudaGetLastError();
}

// raw [n][128] (device) -> resident layout.  block = one allocation laid out by img_layout() (msfm_api.cu);
// scratch: keys [n] u64 | nrm_orig [n] | pos_of [n] | bucket_cnt [8]
cudaError_t launch_desc_format(const uint8_t* raw, int n, int n_pad, uint8_t* sw, uint8_t* ext, int32_t* cg,
                               int32_t* nrm, int32_t* perm, int32_t* inv, int32_t* used, unsigned long long* keys, int32_t* nrm_orig,
                               int32_t* pos_of, int32_t* bucket_cnt, cudaStream_t st) {
    if (n_pad <= 0) return cudaSuccess;
    const size_t ext_bytes = static_cast<size_t>((n_pad + 1023) / 1024) * 256 * 128;
    cudaError_t e;
    if ((e = cudaMemsetAsync(sw, 0, static_cast<size_t>(n_pad) * 128, st)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(ext, 0, ext_bytes, st)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(nrm, 0xFF, static_cast<size_t>(n_pad) * 4, st)) != cudaSuccess) return e;     // -1: dead
    if ((e = cudaMemsetAsync(perm, 0xFF, static_cast<size_t>(n_pad) * 4, st)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(bucket_cnt, 0, 8 * 4, st)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(used, 0, 4, st)) != cudaSuccess) return e;
    if (n > 0) {
        const int wpb = 8;
        int grid = (n + wpb - 1) / wpb;
        if (grid > 148 * 8) grid = 148 * 8;
        { desc_norm_key_kernel<<<grid, wpb * 32, 0, st>>>(raw, n, nrm_orig, keys, bucket_cnt); MSFM_COUNT_LAUNCH(); }
        if ((e = cudaMemsetAsync(pos_of, 0, static_cast<size_t>(n) * 4, st)) != cudaSuccess) return e;
        const int slices = n >= 2048 ? 8 : 1;
        { desc_rank_kernel<<<dim3((n + 255) / 256, slices), 256, 0, st>>>(keys, n, bucket_cnt, pos_of, used); MSFM_COUNT_LAUNCH(); }
        { desc_scatter_kernel<<<grid, wpb * 32, 0, st>>>(raw, n, pos_of, nrm_orig, bucket_cnt, sw, nrm, perm, inv); MSFM_COUNT_LAUNCH(); }
    }
    int ggrid = (n_pad / 32 + 7) / 8;
    { desc_groups_kernel<<<ggrid, 256, 0, st>>>(nrm, n_pad, cg, ext); MSFM_COUNT_LAUNCH(); }
    return cudaGetLastError();
}
cudaError_t launch_build_units(const SegDev* segs, int nseg, int num_units, UnitDev* units, cudaStream_t st) {
    if (num_units <= 0) return cudaSuccess;
    { build_units_kernel<<<(num_units + 255) / 256, 256, 0, st>>>(segs, nseg, num_units, units); MSFM_COUNT_LAUNCH(); }
    return cudaGetLastError();
}
cudaError_t launch_resolve_rows(const ImgDev* imgs, const UnitDev* units, int unit0, int num_units, const int32_t* res_g,
                                const int32_t* res_d1, const int32_t* res_u, MatchOpts opt, int32_t* m_j, int32_t* m_d1,
                                int32_t* m_d2, int32_t* m_j0, int32_t* exact_list, unsigned int* counters,
                                cudaStream_t st) {
    if (num_units <= 0) return cudaSuccess;
    { resolve_rows_kernel<<<num_units, kUnitRows, 0, st>>>(imgs, units, unit0, num_units, res_g, res_d1, res_u, opt, m_j, m_d1, m_d2,
                                                   m_j0, exact_list, counters); MSFM_COUNT_LAUNCH(); }
    return cudaGetLastError();
}
cudaError_t launch_exact_rows(const ImgDev* imgs, const UnitDev* units, const int32_t* row_list,
                              const unsigned int* row_count_dev, int row_count_host, MatchOpts opt, int32_t* m_j,
                              int32_t* m_d1, int32_t* m_d2, int32_t* m_j0, int num_sms, cudaStream_t st) {
    { exact_rows_kernel<<<num_sms * 8, 256, 0, st>>>(imgs, units, row_list, row_count_dev, row_count_host, opt, m_j, m_d1,
                                                   m_d2, m_j0); MSFM_COUNT_LAUNCH(); }
    return cudaGetLastError();
}
cudaError_t launch_count_scan_write(const ImgDev* imgs, const SegDev* segs, int npairs, MatchOpts opt,
                                    const int32_t* m_j, const int32_t* m_d1, int32_t* counts, long long* offsets,
                                    long long* running_total, long long capacity, int32_t* out_matches, float* out_dist,
                                    cudaStream_t st) {
    if (npairs <= 0) return cudaSuccess;
    { count_matches_kernel<<<npairs, 256, 0, st>>>(imgs, segs, npairs, opt, m_j, m_d1, counts); MSFM_COUNT_LAUNCH(); }
    { scan_counts_kernel<<<1, 1024, 0, st>>>(counts, npairs, offsets, running_total); MSFM_COUNT_LAUNCH(); }
    { write_matches_kernel<<<npairs, 256, 0, st>>>(imgs, segs, npairs, opt, m_j, m_d1, offsets, capacity, out_matches,
                                                 out_dist); MSFM_COUNT_LAUNCH(); }
    return cudaGetLastError();
}

}  // namespace msfm
