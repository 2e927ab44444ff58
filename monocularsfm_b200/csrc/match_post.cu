// M-path support kernels around the tensor kernel (match_k1.cu):
//   desc_format_kernel     raw [n][128] u8  ->  swizzled resident layout + column constants
//   build_units_kernel     segment table    ->  unit table of one batch
//   resolve_rows_kernel    K1 row results   ->  ratio test; rescans the winner's 32-column group when the
//                                               runner-up could hide there; defers tie / sqrt-collapse rows
//   exact_rows_kernel      exact CUDA-core scan in OpenCV's (sqrtf(d2), index) order for deferred rows
//                          (and for every row in msfm_match_knn2_u8 mode 1)
//   count/scan/write       CrossCheck (FeatureUtils.cpp:281-310) + FilterMatchesByDistance (:208-218) +
//                          ordered compaction to the CSR output
// All integer arithmetic is exact; the float steps reproduce the reference's: distance = sqrtf((float)d2)
// (OpenCV batchDistL2_), ratio test `d0 < ratio * d1` in float (FeatureUtils.cpp:152).
#include "match_types.cuh"
#include <cuda_runtime.h>

namespace msfm {

// ------------------------------------------------------------------------------------------------
__global__ void desc_format_kernel(const uint8_t* __restrict__ raw, int n, int n_pad, uint8_t* __restrict__ sw,
                                   int32_t* __restrict__ cj) {
    const int warps_per_block = blockDim.x >> 5;
    const int lane = threadIdx.x & 31;
    for (int r = blockIdx.x * warps_per_block + (threadIdx.x >> 5); r < n_pad; r += gridDim.x * warps_per_block) {
        uint32_t w = 0;
        if (r < n) w = reinterpret_cast<const uint32_t*>(raw + static_cast<size_t>(r) * 128)[lane];
        int32_t s = __dp4a(w, w, 0u);                      // sum of the 4 squared bytes (unsigned)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        const int chunk = lane >> 2;                          // 16-byte chunk of this lane's word
        const int pos = ((chunk ^ (r & 7)) << 2) | (lane & 3);
        reinterpret_cast<uint32_t*>(sw + static_cast<size_t>(r) * 128)[pos] = w;
        if (lane == 0) cj[r] = (r < n) ? (s * 256 + (r & 255)) : (kPadKey | (r & 255));
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

__device__ __forceinline__ float dist_of(int32_t d2) { return __fsqrt_rn(static_cast<float>(d2)); }
__device__ __forceinline__ bool ratio_pass(int32_t d1, int32_t d2, float ratio) {
    return dist_of(d1) < __fmul_rn(ratio, dist_of(d2));
}
__device__ __forceinline__ bool dist_filter_ok(int32_t d1, double max_distance) {
    // FilterMatchesByDistance: drop when (double)distance > max_distance
    return max_distance < 0.0 || !(static_cast<double>(dist_of(d1)) > max_distance);
}

// ------------------------------------------------------------------------------------------------
// One block of 128 threads per unit; thread r owns row r of the unit.
// Outputs (global row index = unit*128 + r):
//   m_j   train index of the accepted match or -1          m_d1  exact d2 of the best column (kIntInf = none)
//   m_d2  exact d2 of the runner-up when it was computed, else an upper bound (kIntInf = none)
//   m_j0  [2g] best column regardless of the ratio test (knn2 API), [2g+1] runner-up column (exact path only) or -1
__global__ void __launch_bounds__(128)
resolve_rows_kernel(const ImgDev* __restrict__ imgs, const UnitDev* __restrict__ units, int num_units,
                    const int32_t* __restrict__ res_j, const int32_t* __restrict__ res_d1,
                    const int32_t* __restrict__ res_u, MatchOpts opt, int32_t* __restrict__ m_j,
                    int32_t* __restrict__ m_d1, int32_t* __restrict__ m_d2, int32_t* __restrict__ m_j0,
                    int32_t* __restrict__ exact_list, unsigned int* __restrict__ counters /*[0]=exact,[1]=rescans*/) {
    const int u = blockIdx.x;
    if (u >= num_units) return;
    const UnitDev unit = units[u];
    const ImgDev q = imgs[unit.q_slot];
    const ImgDev t = imgs[unit.t_slot];
    const int lane = threadIdx.x & 31;
    const int r = threadIdx.x;
    const int qi = unit.row_block * 128 + r;
    const size_t g = static_cast<size_t>(u) * 128 + r;

    int32_t j1 = -1, d1 = kIntInf, uu = kIntInf;
    bool valid = qi < q.n;
    if (valid) { j1 = res_j[g]; d1 = res_d1[g]; uu = res_u[g]; }
    if (valid && (j1 < 0 || j1 >= t.n)) { j1 = -1; d1 = kIntInf; uu = kIntInf; }
    // rows that must be redone exactly: cross-group tie of the best distance, or float-sqrt collapse range
    const bool need_exact = valid && j1 >= 0 && t.n >= 2 && (uu == d1 || d1 >= kSqrtExactLimit);
    // the runner-up can only matter if the row passes against the upper bound uu (or the caller wants it exactly)
    bool need_rescan = valid && j1 >= 0 && t.n >= 2 && !need_exact &&
                       (opt.exact_second || uu == kIntInf || ratio_pass(d1, uu, opt.ratio));

    int32_t d2 = uu;
    // ---- warp-cooperative rescan of the winner's 32-column group
    unsigned todo = __ballot_sync(0xffffffffu, need_rescan);
    while (todo) {
        const int src = __ffs(todo) - 1;
        todo &= todo - 1;
        const int sj1 = __shfl_sync(0xffffffffu, j1, src);
        const int sqi = __shfl_sync(0xffffffffu, qi, src);
        const int col = (sj1 & ~31) + lane;
        int32_t dd = kIntInf;
        if (col < t.n && col != sj1) dd = sqdist_rows(q.sw, sqi, t.sw, col);
        // only the runner-up's VALUE matters for the ratio test
        const int32_t best = __reduce_min_sync(0xffffffffu, dd);
        if (lane == src) d2 = min(d2, best);
    }
    const unsigned nres = __popc(__ballot_sync(0xffffffffu, need_rescan));
    if (lane == 0 && nres) atomicAdd(&counters[1], nres);

    if (need_exact) {
        const unsigned slot = atomicAdd(&counters[0], 1u);
        exact_list[slot] = static_cast<int32_t>(g);
    }
    if (valid) {
        int32_t mj = -1;
        if (!need_exact && j1 >= 0 && t.n >= 2 && d2 != kIntInf && ratio_pass(d1, d2, opt.ratio)) mj = j1;
        m_j[g] = mj;
        m_d1[g] = d1;
        m_d2[g] = (t.n >= 2) ? d2 : kIntInf;
        m_j0[2 * g] = j1;
        m_j0[2 * g + 1] = -1;      // the tensor path tracks the runner-up's distance, not its column
    }
}

// ------------------------------------------------------------------------------------------------
// Exact scan: one warp per listed row, all train columns, OpenCV order = (sqrtf(d2), column) lexicographic.
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

__global__ void __launch_bounds__(256)
exact_rows_kernel(const ImgDev* __restrict__ imgs, const UnitDev* __restrict__ units,
                  const int32_t* __restrict__ row_list, const unsigned int* __restrict__ row_count_dev,
                  int row_count_host /* >=0: use this instead of the device counter */, MatchOpts opt,
                  int32_t* __restrict__ m_j, int32_t* __restrict__ m_d1, int32_t* __restrict__ m_d2,
                  int32_t* __restrict__ m_j0) {
    const int nrows = row_count_host >= 0 ? row_count_host : static_cast<int>(*row_count_dev);
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    for (int w = blockIdx.x * wpb + (threadIdx.x >> 5); w < nrows; w += gridDim.x * wpb) {
        const size_t g = row_list ? static_cast<size_t>(row_list[w]) : static_cast<size_t>(w);
        const int u = static_cast<int>(g >> 7);
        const UnitDev unit = units[u];
        const ImgDev q = imgs[unit.q_slot];
        const ImgDev t = imgs[unit.t_slot];
        const int qi = unit.row_block * 128 + static_cast<int>(g & 127);
        if (qi >= q.n) continue;
        Top2 s{kIntInf, -1, kIntInf, -1};
        for (int col = lane; col < t.n; col += 32) top2_insert(s, sqdist_rows(q.sw, qi, t.sw, col), col);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            Top2 other;
            other.d0 = __shfl_xor_sync(0xffffffffu, s.d0, o);
            other.j0 = __shfl_xor_sync(0xffffffffu, s.j0, o);
            other.d1 = __shfl_xor_sync(0xffffffffu, s.d1, o);
            other.j1 = __shfl_xor_sync(0xffffffffu, s.j1, o);
            top2_insert(s, other.d0, other.j0);
            top2_insert(s, other.d1, other.j1);
        }
        if (lane == 0) {
            int32_t mj = -1;
            if (s.j0 >= 0 && s.j1 >= 0 && ratio_pass(s.d0, s.d1, opt.ratio)) mj = s.j0;
            m_j[g] = mj;
            m_d1[g] = s.j0 >= 0 ? s.d0 : kIntInf;
            m_d2[g] = s.j1 >= 0 ? s.d1 : kIntInf;
            m_j0[2 * g] = s.j0;
            m_j0[2 * g + 1] = s.j1;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// CrossCheck + distance filter predicate for query row i of pair segment(s).
struct PairView {
    int n1, n2;
    size_t base12, base21;   // global row base of the two directions (base21 unused without cross-check)
};
__device__ __forceinline__ PairView pair_view(const ImgDev* imgs, const SegDev* segs, int p, int cross) {
    const SegDev s12 = segs[cross ? 2 * p : p];
    PairView v;
    v.n1 = imgs[s12.q_slot].n;
    v.n2 = imgs[s12.t_slot].n;
    v.base12 = static_cast<size_t>(s12.unit_base) * 128;
    v.base21 = cross ? static_cast<size_t>(segs[2 * p + 1].unit_base) * 128 : 0;
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
    const PairView v = pair_view(imgs, segs, p, opt.cross_check);
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
    const PairView v = pair_view(imgs, segs, p, opt.cross_check);
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

// ------------------------------------------------------------------------------------------------ launchers
cudaError_t launch_desc_format(const uint8_t* raw, int n, int n_pad, uint8_t* sw, int32_t* cj, cudaStream_t st) {
    if (n_pad <= 0) return cudaSuccess;
    const int wpb = 8;
    int grid = (n_pad + wpb - 1) / wpb;
    if (grid > 148 * 8) grid = 148 * 8;
    desc_format_kernel<<<grid, wpb * 32, 0, st>>>(raw, n, n_pad, sw, cj);
    return cudaGetLastError();
}
cudaError_t launch_build_units(const SegDev* segs, int nseg, int num_units, UnitDev* units, cudaStream_t st) {
    if (num_units <= 0) return cudaSuccess;
    build_units_kernel<<<(num_units + 255) / 256, 256, 0, st>>>(segs, nseg, num_units, units);
    return cudaGetLastError();
}
cudaError_t launch_resolve_rows(const ImgDev* imgs, const UnitDev* units, int num_units, const int32_t* res_j,
                                const int32_t* res_d1, const int32_t* res_u, MatchOpts opt, int32_t* m_j, int32_t* m_d1,
                                int32_t* m_d2, int32_t* m_j0, int32_t* exact_list, unsigned int* counters,
                                cudaStream_t st) {
    if (num_units <= 0) return cudaSuccess;
    resolve_rows_kernel<<<num_units, 128, 0, st>>>(imgs, units, num_units, res_j, res_d1, res_u, opt, m_j, m_d1, m_d2,
                                                   m_j0, exact_list, counters);
    return cudaGetLastError();
}
cudaError_t launch_exact_rows(const ImgDev* imgs, const UnitDev* units, const int32_t* row_list,
                              const unsigned int* row_count_dev, int row_count_host, MatchOpts opt, int32_t* m_j,
                              int32_t* m_d1, int32_t* m_d2, int32_t* m_j0, int num_sms, cudaStream_t st) {
    exact_rows_kernel<<<num_sms * 4, 256, 0, st>>>(imgs, units, row_list, row_count_dev, row_count_host, opt, m_j, m_d1,
                                                   m_d2, m_j0);
    return cudaGetLastError();
}
cudaError_t launch_count_scan_write(const ImgDev* imgs, const SegDev* segs, int npairs, MatchOpts opt,
                                    const int32_t* m_j, const int32_t* m_d1, int32_t* counts, long long* offsets,
                                    long long* running_total, long long capacity, int32_t* out_matches, float* out_dist,
                                    cudaStream_t st) {
    if (npairs <= 0) return cudaSuccess;
    count_matches_kernel<<<npairs, 256, 0, st>>>(imgs, segs, npairs, opt, m_j, m_d1, counts);
    scan_counts_kernel<<<1, 1024, 0, st>>>(counts, npairs, offsets, running_total);
    write_matches_kernel<<<npairs, 256, 0, st>>>(imgs, segs, npairs, opt, m_j, m_d1, offsets, capacity, out_matches,
                                                 out_dist);
    return cudaGetLastError();
}

}  // namespace msfm
