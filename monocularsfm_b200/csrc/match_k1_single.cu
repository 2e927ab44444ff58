// K1, single-CTA variant (diagnostic: MSFM_K1_SINGLE=1; the default is the CTA-pair kernel of match_k1.cu) —
// match_tile_kernel: tcgen05 u8 x u8 -> s32 score tiles with a fused per-row top-2 reduction.
//
// Replaces the arithmetic of OpenCV's batchDistance + k-NN update that the reference reaches through
// FeatureUtils::ComputeMatches (src/Feature/FeatureUtils.cpp:146-149, knnMatch k=2).
//
// Work unit = 128 query rows (one UMMA M=128 tile, resident in smem) against ALL columns of the train image,
// streamed as 256-column tiles (UMMA N=256).  Per tile the tensor core runs FIVE tcgen05.mma.kind::i8 (K = 32 bytes
// each): four over the 128 descriptor bytes and one over the 32 "extension" bytes that add e_j = (C_g - ||d_j||^2)/2
// to every column (match_types.cuh), so the accumulator is  acc(i,j) = q_i.d_j + e_j  and, inside a 32-column group,
//     C_g - 2 * max_j acc(i,j) = min_j (||d_j||^2 - 2 q_i.d_j) = min_j d2(i,j) - ||q_i||^2 .
// The epilogue therefore needs ONE integer max per two elements (VIMNMX3) and no per-column constants.
// Persistent CTAs (one per SM) walk the unit table with stride gridDim.x.
//
// Warp roles (384 threads):
//   warp 0      bulk-copy producer (cp.async.bulk = TMA engine, one elected lane): query tile, train tiles,
//               extension super-tiles (one per 4 train tiles)
//   warps 1-2   TWO MMA issuers (one elected lane each): issuer w owns the tiles with (tile counter & 1) == w, i.e. the
//               accumulator stage w.  Measured on B200 (profiles/r01_microbench_mma_patterns.log): one pass of an
//               issuer's loop (mbarrier waits + 5 MMAs + commits) has ~1000-1250 cycles of latency that only overlaps
//               with >= ~1300 cycles of queued tensor work; a single issuer therefore caps a 640-cycle tile at ~1030
//               cycles (versions 1-4 of this kernel all sat there).  Several issuers interleave their latency chains.
//   warp 2      also the TMEM allocator
//   warps 4-11  epilogue: two warpgroups, each owns 128 of a tile's 256 columns; thread = query row
//               (tcgen05.ld 32x32b: lane <-> row), so the row-wise reduction is thread-local.
//
// Output per row (sorted space of the query image): g1 = group holding the best column, d1 = exact squared distance of
// the best column, u = exact squared distance of the best column of any OTHER group.  match_post.cu finds the best
// column and the in-group runner-up by rescanning the 32 columns of g1 — only for rows that can still pass the ratio
// test — and re-does rows whose best distance ties across groups (u == d1) or may collapse under float sqrt
// (d1 >= 2^22) with the exact scan.
#include "match_types.cuh"
#include "ptx.cuh"

namespace msfm {
namespace k1s {

constexpr int BM = kUnitRows;                // 128 query rows per unit (UMMA M)
constexpr int BN = 256;                      // train columns per tile (UMMA N)
constexpr int KBYTES = 128;
constexpr int UMMA_KB = 32;                  // bytes of K per tcgen05.mma.kind::i8
constexpr int NS = 3;                        // smem stages of the train-tile ring
constexpr int NE = 2;                        // smem stages of the extension super-tile ring (one per 4 tiles)
constexpr uint32_t A_BYTES = BM * KBYTES;    // 16 KiB
constexpr uint32_t B_BYTES = BN * KBYTES;    // 32 KiB
constexpr uint32_t E_BYTES = BN * 128;       // 32 KiB: extension bytes of 4 consecutive tiles
constexpr int NUM_THREADS = 384;
constexpr int NUM_ISSUERS = 2;               // warps 1, 2.  MUST equal the number of accumulator stages: with more
                                             // issuers than stages two of them race for the same stage's parity-tracked
                                             // t_empty barrier (seen as a deadlock, caught by the mbarrier watchdog)
constexpr int EPI_WARP0 = 4;
constexpr int EPI_WARPS = 8;
constexpr int EPI_THREADS = EPI_WARPS * 32;
static_assert(BM == 128, "one UMMA M=128 tile per unit");
static_assert(NUM_ISSUERS == 2, "one issuer per accumulator stage");

constexpr uint32_t OFF_A = 0;
constexpr uint32_t OFF_AEXT = OFF_A + 2 * A_BYTES;
constexpr uint32_t OFF_B = OFF_AEXT + A_BYTES;
constexpr uint32_t OFF_E = OFF_B + NS * B_BYTES;
constexpr uint32_t OFF_MERGE = OFF_E + NE * E_BYTES;            // [2][128] int4 (k1,k2,group,-)
constexpr uint32_t OFF_BAR = OFF_MERGE + 2 * 128 * 16;
constexpr uint32_t NUM_BARS = 2 + 2 + NS + NS + NE + NE + 2 + 2;
constexpr uint32_t OFF_TMEMPTR = OFF_BAR + NUM_BARS * 8;
constexpr uint32_t SMEM_USED = OFF_TMEMPTR + 16;
constexpr uint32_t SMEM_BYTES = SMEM_USED + 1024;               // slack for manual 1024-B alignment

static_assert(OFF_B % 1024 == 0 && OFF_E % 1024 == 0 && OFF_AEXT % 1024 == 0, "swizzle-128B tiles need 1 KiB alignment");
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");

// occupied sorted positions of an image (0 for an empty image, which has no device block)
__device__ __forceinline__ int img_used(const ImgDev& im) { return im.used ? __ldg(im.used) : 0; }

__global__ void __launch_bounds__(NUM_THREADS, 1)
match_tile_kernel(const ImgDev* __restrict__ imgs, const UnitDev* __restrict__ units, int unit0, int num_units,
                  int32_t* __restrict__ res_g, int32_t* __restrict__ res_d1, int32_t* __restrict__ res_u) {
    extern __shared__ uint8_t smem_raw[];
    // manual 1 KiB alignment (dynamic smem is only guaranteed 16-B aligned)
    const uint32_t raw_addr = ptx::smem_u32(smem_raw);
    uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);

    uint8_t* smA = smem + OFF_A;
    uint8_t* smAext = smem + OFF_AEXT;
    uint8_t* smB = smem + OFF_B;
    uint8_t* smE = smem + OFF_E;
    int4* smMerge = reinterpret_cast<int4*>(smem + OFF_MERGE);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
    uint64_t* a_full = bars;            // [2]
    uint64_t* a_empty = a_full + 2;     // [2]   count NUM_ISSUERS
    uint64_t* b_full = a_empty + 2;     // [NS]
    uint64_t* b_empty = b_full + NS;    // [NS]  count 1: the owning issuer's commit
    uint64_t* e_full = b_empty + NS;    // [NE]
    uint64_t* e_empty = e_full + NE;    // [NE]  count NUM_ISSUERS
    uint64_t* t_full = e_empty + NE;    // [2]
    uint64_t* t_empty = t_full + 2;     // [2]   count 8: epilogue warps
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + OFF_TMEMPTR);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    // constant A-side extension tile: every row is (1, 255 x 31, 0 ...) in the swizzled layout
    for (int i = threadIdx.x; i < BM * 8; i += NUM_THREADS) {
        const int r = i >> 3, pos = i & 7;
        const int chunk = pos ^ (r & 7);
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (chunk == 0) v = make_uint4(0xFFFFFF01u, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
        else if (chunk == 1) v = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
        reinterpret_cast<uint4*>(smAext)[i] = v;
    }
    ptx::fence_proxy_async();            // generic-proxy writes -> visible to the tensor core (async proxy)

    if (warp == 1 && lane == 0) {
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(&a_full[i], 1);
            ptx::mbar_init(&a_empty[i], NUM_ISSUERS);
            ptx::mbar_init(&t_full[i], 1);
            ptx::mbar_init(&t_empty[i], EPI_WARPS);
        }
        for (int i = 0; i < NS; ++i) {
            ptx::mbar_init(&b_full[i], 1);
            ptx::mbar_init(&b_empty[i], 1);
        }
        for (int i = 0; i < NE; ++i) {
            ptx::mbar_init(&e_full[i], 1);
            ptx::mbar_init(&e_empty[i], NUM_ISSUERS);
        }
        ptx::fence_mbar_init();
    }
    if (warp == 2) ptx::tmem_alloc<512>(tmem_ptr);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        // ===================================================================== producer
        if (lane == 0) {
            uint32_t it = 0, un = 0, et = 0;
            for (int u = unit0 + blockIdx.x; u < unit0 + num_units; u += gridDim.x) {
                const UnitDev unit = units[u];
                const ImgDev q = imgs[unit.q_slot];
                const ImgDev t = imgs[unit.t_slot];
                if (unit.row_block * BM >= img_used(q)) continue;      // all-dead query rows: every role skips the unit
                const uint32_t cu = un++;                              // index among the units this CTA really processes
                const uint32_t ab = cu & 1;
                ptx::mbar_wait(&a_empty[ab], ((cu >> 1) & 1) ^ 1);
                ptx::mbar_arrive_expect_tx(&a_full[ab], A_BYTES);
                ptx::bulk_g2s(smA + ab * A_BYTES, q.sw + static_cast<size_t>(unit.row_block) * A_BYTES, A_BYTES,
                              &a_full[ab]);
                const int ntiles = (img_used(t) + BN - 1) / BN;          // tiles beyond the occupied columns are skipped
                for (int tile = 0; tile < ntiles; ++tile, ++it) {
                    if ((tile & 3) == 0) {
                        const uint32_t es = et % NE;
                        ptx::mbar_wait(&e_empty[es], ((et / NE) & 1) ^ 1);
                        ptx::mbar_arrive_expect_tx(&e_full[es], E_BYTES);
                        ptx::bulk_g2s(smE + es * E_BYTES, t.ext + static_cast<size_t>(tile >> 2) * E_BYTES, E_BYTES, &e_full[es]);
                        ++et;
                    }
                    const uint32_t s = it % NS;
                    const uint32_t ph = (it / NS) & 1;
                    ptx::mbar_wait(&b_empty[s], ph ^ 1);
                    ptx::mbar_arrive_expect_tx(&b_full[s], B_BYTES);
                    ptx::bulk_g2s(smB + s * B_BYTES, t.sw + static_cast<size_t>(tile) * B_BYTES, B_BYTES, &b_full[s]);
                }
            }
        }
    } else if (warp >= 1 && warp <= NUM_ISSUERS) {
        // ===================================================================== MMA issuers
        if (lane == 0) {
            const uint32_t me = static_cast<uint32_t>(warp - 1);   // owns the tiles with it % NUM_ISSUERS == me
            constexpr uint32_t idesc = ptx::make_idesc_u8(BM, BN);
            const uint64_t aext_desc = ptx::make_smem_desc_sw128(ptx::smem_u32(smAext));
            uint32_t it = 0, un = 0, et = 0;
            for (int u = unit0 + blockIdx.x; u < unit0 + num_units; u += gridDim.x) {
                const UnitDev unit = units[u];
                const ImgDev t = imgs[unit.t_slot];
                if (unit.row_block * BM >= img_used(imgs[unit.q_slot])) continue;
                const uint32_t cu = un++;
                const uint32_t ab = cu & 1;
                const int tcols = img_used(t);
                const int ntiles = (tcols + BN - 1) / BN;
                bool have_a = false, have_e = false;
                uint64_t a_desc0 = 0;
                for (int tile = 0; tile < ntiles; ++tile, ++it) {
                    const uint32_t es = et % NE;
                    if (it % NUM_ISSUERS == me) {
                        if (!have_a) {
                            ptx::mbar_wait(&a_full[ab], (cu >> 1) & 1);
                            a_desc0 = ptx::make_smem_desc_sw128(ptx::smem_u32(smA + ab * A_BYTES));
                            have_a = true;
                        }
                        const uint32_t s = it % NS;
                        const uint32_t ph = (it / NS) & 1;
                        const uint32_t acc = it & 1;
                        const uint32_t aph = (it >> 1) & 1;
                        if (!have_e) {                     // first owned tile of this super-tile
                            ptx::mbar_wait(&e_full[es], (et / NE) & 1);
                            have_e = true;
                        }
                        ptx::mbar_wait(&b_full[s], ph);
                        ptx::mbar_wait(&t_empty[acc], aph ^ 1);
                        ptx::tc_fence_after();
                        const uint64_t b_desc0 = ptx::make_smem_desc_sw128(ptx::smem_u32(smB + s * B_BYTES));
                        const uint64_t e_desc0 = ptx::make_smem_desc_sw128(ptx::smem_u32(smE + es * E_BYTES));
                        const uint32_t d_tmem = tmem_base + acc * BN;
                        // the last tile of an image may hold fewer than 256 occupied columns (a multiple of 32): narrower N
                        const int ncols = min(BN, tcols - tile * BN);
                        const uint32_t idesc_t = ncols == BN ? idesc : ptx::make_idesc_u8(BM, static_cast<uint32_t>(ncols));
#pragma unroll
                        for (int k = 0; k < KBYTES / UMMA_KB; ++k) {
                            // advancing K by 32 bytes inside the 128-B swizzle atom = +2 in the (addr >> 4) field
                            ptx::mma_i8_ss(d_tmem, a_desc0 + 2 * k, b_desc0 + 2 * k, idesc_t, k > 0 ? 1u : 0u);
                        }
                        // + e_j : the tile's 32 extension bytes sit at K offset 32*(tile%4) of the super-tile rows
                        ptx::mma_i8_ss(d_tmem, aext_desc, e_desc0 + 2 * (tile & 3), idesc_t, 1u);
                        ptx::mma_commit(&b_empty[s]);
                        ptx::mma_commit(&t_full[acc]);
                    }
                    if ((tile & 3) == 3 || tile == ntiles - 1) {
                        ptx::mma_commit(&e_empty[es]);     // this issuer's MMAs on the super-tile (if any) are done
                        ++et;
                        have_e = false;
                    }
                }
                ptx::mma_commit(&a_empty[ab]);
            }
        }
    } else if (warp >= EPI_WARP0) {
        // ===================================================================== epilogue
        const int wg = (warp - EPI_WARP0) >> 2;          // column half of the tile
        const int quarter = warp & 3;                    // TMEM lane quarter this warp may access
        const int row = quarter * 32 + lane;             // query row inside the unit
        uint32_t it = 0, un = 0;
        for (int u = unit0 + blockIdx.x; u < unit0 + num_units; u += gridDim.x) {
            const UnitDev unit = units[u];
            const ImgDev q = imgs[unit.q_slot];
            const ImgDev t = imgs[unit.t_slot];
            if (unit.row_block * BM >= img_used(q)) continue;
            const uint32_t cu = un++;
            const int tcols = img_used(t);
            const int ntiles = (tcols + BN - 1) / BN;
            int32_t k1 = kIntInf, k2 = kIntInf, g1 = 0;
            for (int tile = 0; tile < ntiles; ++tile, ++it) {
                const uint32_t acc = it & 1;
                const uint32_t aph = (it >> 1) & 1;
                const int4 cgv = __ldg(reinterpret_cast<const int4*>(t.cg) + tile * 2 + wg);   // C_g of this half's 4 groups
                ptx::mbar_wait(&t_full[acc], aph);
                ptx::tc_fence_after();
                const uint32_t taddr0 = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BN + wg * 128;
                // occupied 32-column groups of this warpgroup's half (4 except in the narrow last tile of an image)
                const int nch = min(4, max(0, (min(BN, tcols - tile * BN) - wg * 128) >> 5));
                uint32_t va[32], vb[32];
                ptx::tmem_ld_32x32(taddr0, va);
                ptx::tmem_ld_wait();
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    // software pipeline: chunk c+1 is in flight while chunk c is reduced
                    uint32_t(&cur)[32] = (c & 1) ? vb : va;
                    uint32_t(&nxt)[32] = (c & 1) ? va : vb;
                    if (c < 3) ptx::tmem_ld_32x32(taddr0 + (c + 1) * 32, nxt);
                    // max over the 32 columns of the group: four independent chains, two elements per VIMNMX3
                    int32_t m0 = static_cast<int32_t>(cur[0]), m1 = static_cast<int32_t>(cur[1]);
                    int32_t m2 = static_cast<int32_t>(cur[2]), m3 = static_cast<int32_t>(cur[3]);
#pragma unroll
                    for (int e = 4; e < 28; e += 8) {
                        m0 = __vimax3_s32(m0, static_cast<int32_t>(cur[e + 0]), static_cast<int32_t>(cur[e + 1]));
                        m1 = __vimax3_s32(m1, static_cast<int32_t>(cur[e + 2]), static_cast<int32_t>(cur[e + 3]));
                        m2 = __vimax3_s32(m2, static_cast<int32_t>(cur[e + 4]), static_cast<int32_t>(cur[e + 5]));
                        m3 = __vimax3_s32(m3, static_cast<int32_t>(cur[e + 6]), static_cast<int32_t>(cur[e + 7]));
                    }
                    m0 = __vimax3_s32(m0, static_cast<int32_t>(cur[28]), static_cast<int32_t>(cur[29]));
                    m1 = __vimax3_s32(m1, static_cast<int32_t>(cur[30]), static_cast<int32_t>(cur[31]));
                    const int32_t m = max(__vimax3_s32(m0, m1, m2), m3);
                    const int32_t cgc = (c == 0) ? cgv.x : (c == 1) ? cgv.y : (c == 2) ? cgv.z : cgv.w;
                    const int32_t key = cgc - 2 * m;      // = min over the group of (||d_j||^2 - 2 q.d_j)
                    // insert the group minimum into the running (best, best-of-other-groups)
                    if (c < nch) {
                        k2 = min(k2, max(k1, key));
                        if (key < k1) g1 = tile * 8 + wg * 4 + c;
                        k1 = min(k1, key);
                    }
                    if (c < 3) {
                        ptx::tmem_ld_wait();
                        if (c == 2) {
                            // the last chunk is in registers: hand the accumulator stage back to the tensor core
                            ptx::tc_fence_before();
                            __syncwarp();
                            if (lane == 0) ptx::mbar_arrive(&t_empty[acc]);
                        }
                    }
                }
            }
            // ---- unit end: fold the two column halves, write the row results
            int4* mbuf = smMerge + (cu & 1) * 128;
            if (wg == 1) mbuf[row] = make_int4(k1, k2, g1, 0);
            ptx::bar_sync(1, EPI_THREADS);
            if (wg == 0) {
                const int4 o = mbuf[row];
                k2 = __vimin3_s32(k2, o.y, max(k1, o.x));
                if (o.x < k1) g1 = o.z;
                k1 = min(k1, o.x);
                const int grow = unit.row_block * BM + row;           // sorted-space row of the query image (< n_pad)
                const int32_t ni = q.nrm[grow];                        // ||q_i||^2, -1 for dead rows
                const size_t out = static_cast<size_t>(u) * BM + row;
                const bool have1 = ni >= 0 && k1 < kDeadKey;
                const bool have2 = have1 && k2 < kDeadKey;
                res_g[out] = have1 ? g1 : -1;
                res_d1[out] = have1 ? (k1 + ni) : kIntInf;
                res_u[out] = have2 ? (k2 + ni) : kIntInf;
            }
        }
    }

    // ---- teardown
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 2) ptx::tmem_dealloc<512>(tmem_base);
}

}  // namespace k1s

// host launcher (called from msfm_api.cu)
cudaError_t launch_match_tile_single(const ImgDev* imgs, const UnitDev* units, int unit0, int num_units, int32_t* res_g,
                                     int32_t* res_d1, int32_t* res_u, int num_sms, cudaStream_t stream) {
    cudaError_t e = cudaFuncSetAttribute(k1s::match_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(k1s::SMEM_BYTES));
    if (e != cudaSuccess) return e;
    if (num_units <= 0) return cudaSuccess;
    const int grid = num_units < num_sms ? num_units : num_sms;
    k1s::match_tile_kernel<<<grid, k1s::NUM_THREADS, k1s::SMEM_BYTES, stream>>>(imgs, units, unit0, num_units, res_g, res_d1,
                                                                            res_u);
    return cudaGetLastError();
}

}  // namespace msfm
