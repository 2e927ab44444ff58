// K1 — match_tile_kernel: tcgen05 u8 x u8 -> s32 distance tiles with a fused per-row top-2 reduction.
//
// Replaces the arithmetic of OpenCV's batchDistance + k-NN update that the reference reaches through
// FeatureUtils::ComputeMatches (src/Feature/FeatureUtils.cpp:146-149, knnMatch k=2).
//
// Work unit = 128 query rows (one UMMA M=128 tile, resident in smem) against ALL columns of the train image,
// streamed as 256-column tiles (UMMA N=256, K = 128 bytes = 4 x tcgen05.mma.kind::i8 of K=32).
// Persistent CTAs (one per SM) walk the unit table with stride gridDim.x.
//
// Warp roles (384 threads):
//   warp 0      bulk-copy producer (cp.async.bulk = TMA engine, one elected lane)
//   warp 1      MMA issuer (one elected lane), accumulators double-buffered in TMEM (2 x 256 columns)
//   warp 2      TMEM allocator
//   warps 4-11  epilogue: two warpgroups, each owns 128 of a tile's 256 columns; thread = query row
//               (tcgen05.ld 32x32b: lane <-> row), so the row-wise reduction is thread-local.
//
// Epilogue arithmetic (all exact integers):
//   key(i,j) = cj[j] - 512*dot(i,j) = 256*(||t_j||^2 - 2 q_i.t_j) + (j & 255)          (one IMAD)
//   d2(i,j)  = (key >> 8) + ||q_i||^2
//   per 32-column group: g = min key (VIMNMX3, two keys per instruction); the packed low byte makes the
//   minimum unique inside a tile and prefers the lowest column on equal distance.
//   across groups: running (k1 = best key, its tile, k2 = best key of any OTHER group).
// Output per row: j1 = best column, d1 = its exact d2, u = exact d2 of the best column outside the winner's
// 32-column group.  The true runner-up is min(u, runner-up inside the winner's group); match_post.cu
// rescans those <=31 columns only for rows that can still pass the ratio test.  Rows whose best distance
// ties across groups (u == d1) or may collapse under float sqrt (d1 >= 2^22) are re-done exactly there.
#include "match_types.cuh"
#include "ptx.cuh"

namespace msfm {
namespace k1 {

constexpr int BM = 128;
constexpr int BN = 256;
constexpr int KBYTES = 128;
constexpr int UMMA_KB = 32;                  // bytes of K per tcgen05.mma.kind::i8
constexpr int NS = 4;                        // smem stages of the train-tile ring
constexpr uint32_t A_BYTES = BM * KBYTES;    // 16 KiB
constexpr uint32_t B_BYTES = BN * KBYTES;    // 32 KiB
constexpr uint32_t CJ_BYTES = BN * 4;        // 1 KiB
constexpr int NUM_THREADS = 384;
constexpr int EPI_WARP0 = 4;
constexpr int EPI_THREADS = 256;

constexpr uint32_t OFF_A = 0;
constexpr uint32_t OFF_B = OFF_A + 2 * A_BYTES;
constexpr uint32_t OFF_CJ = OFF_B + NS * B_BYTES;
constexpr uint32_t OFF_MERGE = OFF_CJ + NS * CJ_BYTES;          // [2][128] int4 (k1,k2,tile,-)
constexpr uint32_t OFF_BAR = OFF_MERGE + 2 * 128 * 16;
constexpr uint32_t NUM_BARS = 2 + 2 + NS + NS + 2 + 2;
constexpr uint32_t OFF_TMEMPTR = OFF_BAR + NUM_BARS * 8;
constexpr uint32_t SMEM_USED = OFF_TMEMPTR + 16;
constexpr uint32_t SMEM_BYTES = SMEM_USED + 1024;               // slack for manual 1024-B alignment

static_assert(OFF_B % 1024 == 0 && B_BYTES % 1024 == 0 && A_BYTES % 1024 == 0, "swizzle-128B tiles need 1 KiB alignment");

__global__ void __launch_bounds__(NUM_THREADS, 1)
match_tile_kernel(const ImgDev* __restrict__ imgs, const UnitDev* __restrict__ units, int num_units,
                  int32_t* __restrict__ res_j, int32_t* __restrict__ res_d1, int32_t* __restrict__ res_u) {
    extern __shared__ uint8_t smem_raw[];
    // manual 1 KiB alignment (dynamic smem is only guaranteed 16-B aligned)
    const uint32_t raw_addr = ptx::smem_u32(smem_raw);
    uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);

    uint8_t* smA = smem + OFF_A;
    uint8_t* smB = smem + OFF_B;
    int32_t* smCJ = reinterpret_cast<int32_t*>(smem + OFF_CJ);
    int4* smMerge = reinterpret_cast<int4*>(smem + OFF_MERGE);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
    uint64_t* a_full = bars;            // [2]
    uint64_t* a_empty = bars + 2;       // [2]
    uint64_t* b_full = bars + 4;        // [NS]
    uint64_t* b_empty = b_full + NS;    // [NS]
    uint64_t* t_full = b_empty + NS;    // [2]
    uint64_t* t_empty = t_full + 2;     // [2]
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + OFF_TMEMPTR);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 1 && lane == 0) {
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(&a_full[i], 1);
            ptx::mbar_init(&a_empty[i], 1);
            ptx::mbar_init(&t_full[i], 1);
            ptx::mbar_init(&t_empty[i], EPI_THREADS / 32);
        }
        for (int i = 0; i < NS; ++i) {
            ptx::mbar_init(&b_full[i], 1);
            ptx::mbar_init(&b_empty[i], 1 + EPI_THREADS / 32);   // MMA commit + one lane per epilogue warp (cj reads)
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
            uint32_t it = 0, un = 0;
            for (int u = blockIdx.x; u < num_units; u += gridDim.x, ++un) {
                const UnitDev unit = units[u];
                const ImgDev q = imgs[unit.q_slot];
                const ImgDev t = imgs[unit.t_slot];
                const uint32_t ab = un & 1;
                ptx::mbar_wait(&a_empty[ab], ((un >> 1) & 1) ^ 1);
                ptx::mbar_arrive_expect_tx(&a_full[ab], A_BYTES);
                ptx::bulk_g2s(smA + ab * A_BYTES, q.sw + static_cast<size_t>(unit.row_block) * A_BYTES, A_BYTES,
                              &a_full[ab]);
                const int ntiles = t.n_pad / BN;
                for (int tile = 0; tile < ntiles; ++tile, ++it) {
                    const uint32_t s = it % NS;
                    const uint32_t ph = (it / NS) & 1;
                    ptx::mbar_wait(&b_empty[s], ph ^ 1);
                    ptx::mbar_arrive_expect_tx(&b_full[s], B_BYTES + CJ_BYTES);
                    ptx::bulk_g2s(smB + s * B_BYTES, t.sw + static_cast<size_t>(tile) * B_BYTES, B_BYTES, &b_full[s]);
                    ptx::bulk_g2s(smCJ + s * BN, t.cj + static_cast<size_t>(tile) * BN, CJ_BYTES, &b_full[s]);
                }
            }
        }
    } else if (warp == 1) {
        // ===================================================================== MMA issuer
        if (lane == 0) {
            constexpr uint32_t idesc = ptx::make_idesc_u8(BM, BN);
            uint32_t it = 0, un = 0;
            for (int u = blockIdx.x; u < num_units; u += gridDim.x, ++un) {
                const UnitDev unit = units[u];
                const ImgDev t = imgs[unit.t_slot];
                const uint32_t ab = un & 1;
                ptx::mbar_wait(&a_full[ab], (un >> 1) & 1);
                const uint64_t a_desc0 = ptx::make_smem_desc_sw128(ptx::smem_u32(smA + ab * A_BYTES));
                const int ntiles = t.n_pad / BN;
                for (int tile = 0; tile < ntiles; ++tile, ++it) {
                    const uint32_t s = it % NS;
                    const uint32_t ph = (it / NS) & 1;
                    const uint32_t acc = it & 1;
                    const uint32_t aph = (it >> 1) & 1;
                    ptx::mbar_wait(&b_full[s], ph);
                    ptx::mbar_wait(&t_empty[acc], aph ^ 1);
                    ptx::tc_fence_after();
                    const uint64_t b_desc0 = ptx::make_smem_desc_sw128(ptx::smem_u32(smB + s * B_BYTES));
                    const uint32_t d_tmem = tmem_base + acc * BN;
#pragma unroll
                    for (int k = 0; k < KBYTES / UMMA_KB; ++k) {
                        // advancing K by 32 bytes inside the 128-B swizzle atom = +2 in the (addr >> 4) field
                        ptx::mma_i8_ss(d_tmem, a_desc0 + 2 * k, b_desc0 + 2 * k, idesc, k > 0 ? 1u : 0u);
                    }
                    ptx::mma_commit(&b_empty[s]);
                    ptx::mma_commit(&t_full[acc]);
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
        for (int u = blockIdx.x; u < num_units; u += gridDim.x, ++un) {
            const UnitDev unit = units[u];
            const ImgDev q = imgs[unit.q_slot];
            const ImgDev t = imgs[unit.t_slot];
            const int ntiles = t.n_pad / BN;
            int32_t k1 = kIntInf, k2 = kIntInf, tk = 0;
            for (int tile = 0; tile < ntiles; ++tile, ++it) {
                const uint32_t s = it % NS;
                const uint32_t ph = (it / NS) & 1;
                const uint32_t acc = it & 1;
                const uint32_t aph = (it >> 1) & 1;
                ptx::mbar_wait(&b_full[s], ph);           // acquire the bulk-copied cj[] of this stage
                ptx::mbar_wait(&t_full[acc], aph);
                ptx::tc_fence_after();
                const int4* cj4 = reinterpret_cast<const int4*>(smCJ + s * BN + wg * 128);
                const uint32_t taddr0 = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BN + wg * 128;
#pragma unroll 1
                for (int c = 0; c < 4; ++c) {
                    uint32_t v[32];
                    ptx::tmem_ld_32x32(taddr0 + c * 32, v);
                    ptx::tmem_ld_wait();
                    int32_t g = kIntInf;
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const int4 cc = cj4[c * 8 + e];
                        const int32_t ka = cc.x - (static_cast<int32_t>(v[4 * e + 0]) << 9);
                        const int32_t kb = cc.y - (static_cast<int32_t>(v[4 * e + 1]) << 9);
                        const int32_t kc = cc.z - (static_cast<int32_t>(v[4 * e + 2]) << 9);
                        const int32_t kd = cc.w - (static_cast<int32_t>(v[4 * e + 3]) << 9);
                        g = __vimin3_s32(g, ka, kb);
                        g = __vimin3_s32(g, kc, kd);
                    }
                    // insert the group minimum into the running (best, best-of-other-groups)
                    k2 = min(k2, max(k1, g));
                    if (g < k1) tk = tile;
                    k1 = min(k1, g);
                }
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    ptx::mbar_arrive(&t_empty[acc]);
                    ptx::mbar_arrive(&b_empty[s]);
                }
            }
            // ---- unit end: fold the two column halves, write the row results
            int4* mbuf = smMerge + (un & 1) * 128;
            if (wg == 1) mbuf[row] = make_int4(k1, k2, tk, 0);
            ptx::bar_sync(1, EPI_THREADS);
            if (wg == 0) {
                const int4 o = mbuf[row];
                k2 = __vimin3_s32(k2, o.y, max(k1, o.x));
                if (o.x < k1) tk = o.z;
                k1 = min(k1, o.x);
                const int grow = unit.row_block * BM + row;           // row inside the query image (< n_pad)
                const int32_t ni = q.cj[grow] >> 8;                     // ||q_i||^2 (padding rows: garbage, ignored)
                const size_t out = static_cast<size_t>(u) * BM + row;
                const bool have1 = k1 < kPadKey;
                const bool have2 = k2 < kPadKey;
                res_j[out] = have1 ? (tk * BN + (k1 & 255)) : -1;
                res_d1[out] = have1 ? ((k1 >> 8) + ni) : kIntInf;
                res_u[out] = have2 ? ((k2 >> 8) + ni) : kIntInf;
            }
        }
    }

    // ---- teardown
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 2) ptx::tmem_dealloc<512>(tmem_base);
}

}  // namespace k1

// host launcher (called from match_api.cu)
cudaError_t launch_match_tile_kernel(const ImgDev* imgs, const UnitDev* units, int num_units, int32_t* res_j,
                                     int32_t* res_d1, int32_t* res_u, int num_sms, cudaStream_t stream) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(k1::match_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             static_cast<int>(k1::SMEM_BYTES));
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    if (num_units <= 0) return cudaSuccess;
    const int grid = num_units < num_sms ? num_units : num_sms;
    k1::match_tile_kernel<<<grid, k1::NUM_THREADS, k1::SMEM_BYTES, stream>>>(imgs, units, num_units, res_j, res_d1,
                                                                            res_u);
    return cudaGetLastError();
}

}  // namespace msfm
