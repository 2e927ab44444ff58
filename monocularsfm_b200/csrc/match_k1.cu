// K1 — match_pair_kernel: tcgen05 u8 x u8 -> s32 score tiles with a fused per-row top-2 reduction, CTA-pair version.
//
// Replaces the arithmetic of OpenCV's batchDistance + k-NN update that the reference reaches through
// FeatureUtils::ComputeMatches (src/Feature/FeatureUtils.cpp:146-149, knnMatch k=2).
//
// Work item = TWO consecutive 128-row units of a segment (256 query rows) against ALL columns of the train image,
// streamed as 256-column tiles, executed by a cluster of two CTAs on the two SMs of a TPC with
// tcgen05.mma.cta_group::2 (M = 256, N = 256, K = 32 bytes): CTA r keeps the query rows of unit u+r and loads only
// columns 128r..128r+127 of every train tile; the tensor cores of the pair exchange the halves, and CTA r receives
// the 128 x 256 accumulator block of its own rows in its own TMEM.  Per tile FIVE MMAs: four over the 128 descriptor
// bytes and one over the 32 "extension" bytes that add e_j = (C_g - ||d_j||^2)/2 to every column (match_types.cuh),
// so the accumulator is  acc(i,j) = q_i.d_j + e_j  and, inside a 32-column group,
//     C_g - 2 * max_j acc(i,j) = min_j (||d_j||^2 - 2 q_i.d_j) = min_j d2(i,j) - ||q_i||^2 .
// The epilogue therefore needs ONE integer max per two elements (VIMNMX3) and no per-column constants.
//
// Why a CTA pair: the single-CTA predecessor of round 1 (128 x 256 tiles, removed; measured in profiles/r01b_bench_single_cta_k1.json) moved, per 640 tensor cycles and SM,
// 40 KB L2 -> smem plus 60 KB smem -> tensor core, i.e. 156 B/clk against the 128 B/clk of an SM's shared memory, and
// 9.4 KB/clk chip-wide against the ~6.3-6.9 KB/clk the L2 delivers (profiles/r01_k1_bench_launch_ncu_raw.csv:
// 12.9 TB/s at 73 % tensor-pipe utilisation).  Sharing every train tile between two SMs halves both.
//
// Warp roles per CTA (384 threads).  The producer, relay and issuer warps run their loops with all 32 lanes and
// warp-uniform operands; one elected lane issues the asynchronous instruction (see match_pair_kernel).
//   warp 0      bulk-copy producer (cp.async.bulk = TMA engine): this CTA's query rows, its half of every train
//               tile (8-stage ring) and of every extension super-tile (one per 4 train tiles, 2-stage ring)
//   warp 1      leader CTA (rank 0): THE MMA issuer, tiles in order, accumulator stage = tile counter & 1.
//               peer CTA (rank 1): relays "my half has landed" from its own full-barriers to the leader's
//               (a 1-D bulk copy can only signal an mbarrier of the CTA it writes to).
//   warp 2      TMEM allocator (both CTAs, cta_group::2 form)
//   warps 4-11  epilogue: two warpgroups, warpgroup w serves accumulator stage w (the even / the odd tiles), so that
//               one warpgroup reduces while the other one reads; thread = query row (tcgen05.ld 32x32b: lane <-> row),
//               the row-wise reduction is thread-local.  A tile is read in two passes of 128 columns (128 registers);
//               the stage goes back to the tensor cores after the second read.
//
// Barriers (every CTA has the full set at the same shared-memory offsets; tcgen05.commit multicasts to both):
//   a_full/b_full/e_full   leader: 2 arrivals (own producer + peer relay); peer: 1 (own producer)
//   done[tile % 16]        ONE commit per tile, multicast to both CTAs: it publishes the accumulator stage to the
//                          epilogues and tells the producers that the tile's smem stage (and, after a super-tile's last
//                          tile, the extension stage) may be refilled
//   a_empty                commit after a unit pair's last tile, multicast
//   t_empty[stage]         leader only: 8 arrivals (the 4 warps of the stage's warpgroup in each CTA)
//
// Output per row (sorted space of the query image): g1 = group holding the best column, d1 = exact squared distance of
// the best column, u = exact squared distance of the best column of any OTHER group.  match_post.cu finds the best
// column and the in-group runner-up by rescanning the 32 columns of g1 — only for rows that can still pass the ratio
// test — and re-does rows whose best distance ties across groups (u == d1) or may collapse under float sqrt
// (d1 >= 2^22) with the exact scan.
#include "match_types.cuh"
#include "ptx.cuh"
#include "launch_count.hpp"
#include <cstdlib>

namespace msfm {


namespace k1 {

constexpr int BM = kUnitRows;                // 128 query rows per CTA (one unit); UMMA M = 256 over the pair
constexpr int BN = 256;                      // train columns per tile (UMMA N)
constexpr int HN = BN / 2;                   // columns of a tile this CTA loads
constexpr int KBYTES = 128;
constexpr int UMMA_KB = 32;                  // bytes of K per tcgen05.mma.kind::i8
constexpr int NS = 8;                        // smem stages of the train-tile ring (power of two: ring index = it & 7)
constexpr int NE = 2;                        // smem stages of the extension super-tile ring (one per 4 tiles)
constexpr uint32_t A_BYTES = BM * KBYTES;    // 16 KiB
constexpr uint32_t B_BYTES = HN * KBYTES;    // 16 KiB: this CTA's half of a train tile
constexpr uint32_t E_BYTES = HN * 128;       // 16 KiB: this CTA's half of an extension super-tile
constexpr int NUM_THREADS = 384;
constexpr int EPI_WARP0 = 4;
constexpr int EPI_WARPS = 8;
constexpr int EPI_THREADS = EPI_WARPS * 32;
constexpr uint16_t BOTH_CTAS = 0b11;
static_assert(BM == 128, "one UMMA M=128 half per CTA");

constexpr uint32_t OFF_A = 0;
constexpr uint32_t OFF_AEXT = OFF_A + 2 * A_BYTES;
constexpr uint32_t OFF_B = OFF_AEXT + A_BYTES;
constexpr uint32_t OFF_E = OFF_B + NS * B_BYTES;
constexpr uint32_t OFF_MERGE = OFF_E + NE * E_BYTES;            // [2][128] int4 (k1,k2,group,-)
constexpr uint32_t OFF_BAR = OFF_MERGE + 2 * 128 * 16;
constexpr int ND = 2 * NS;                   // ring of per-tile "MMAs done" barriers
constexpr uint32_t NUM_BARS = 2 + 2 + NS + NE + ND + 2;
constexpr uint32_t OFF_TMEMPTR = OFF_BAR + NUM_BARS * 8;
constexpr uint32_t SMEM_USED = OFF_TMEMPTR + 16;
constexpr uint32_t SMEM_BYTES = SMEM_USED + 1024;               // slack for manual 1024-B alignment

static_assert(OFF_B % 1024 == 0 && OFF_E % 1024 == 0 && OFF_AEXT % 1024 == 0, "swizzle-128B tiles need 1 KiB alignment");
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");

// occupied sorted positions of an image (0 for an empty image, which has no device block)
__device__ __forceinline__ int img_used(const ImgDev& im) { return im.used ? __ldg(im.used) : 0; }

// Flat work item of one unit pair, written by build_items_kernel right before the tensor kernel.  Every role prefetches
// the NEXT item while it works on the current one: resolving units[u] -> imgs[slot] -> *used on the fly is a chain of
// three dependent global loads (~2000 cycles) per pair in every role — with only two accumulator stages (1280 tensor
// cycles) queued behind the issuers, that chain was a bubble per pair that no per-tile tuning could remove.
struct __align__(16) ItemSrc {     // producer
    const uint8_t* q_rows;         // sorted-space rows of the pair's first unit (the second follows at +A_BYTES)
    const uint8_t* t_sw;
    const uint8_t* t_ext;
    uint64_t pad;
};
struct __align__(16) ItemEpi {     // epilogue
    const int32_t* t_cg;
    const int32_t* q_nrm;          // ||q||^2 of the first unit's rows (second at +BM)
};
struct __align__(16) ItemCtl {     // every role
    int32_t tcols;                 // occupied columns of the train image
    int32_t live;                  // live units of the pair: 0 (skip), 1, 2
    int32_t unit;                  // index of the first unit (K1 result slot = unit * BM + row)
    int32_t pad;
};
struct __align__(16) Item {
    ItemSrc src;
    ItemEpi epi;
    ItemCtl ctl;
};
static_assert(sizeof(Item) == 64 && sizeof(ItemSrc) == 32 && sizeof(ItemEpi) == 16 && sizeof(ItemCtl) == 16, "item layout");

template <typename T>
__device__ __forceinline__ T load_part(const T* p) {     // 16-byte read-only loads
    T out;
    const int4* s = reinterpret_cast<const int4*>(p);
    int4* d = reinterpret_cast<int4*>(&out);
#pragma unroll
    for (int i = 0; i < static_cast<int>(sizeof(T) / 16); ++i) d[i] = __ldg(s + i);
    return out;
}

__global__ void build_items_kernel(const ImgDev* __restrict__ imgs, const UnitDev* __restrict__ units, int unit0,
                                   int num_items, Item* __restrict__ items) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= num_items) return;
    const int u = unit0 + 2 * i;
    const UnitDev unit = units[u];
    const ImgDev q = imgs[unit.q_slot];
    const ImgDev t = imgs[unit.t_slot];
    const int qused = img_used(q);
    Item it;
    it.src.q_rows = q.sw + static_cast<size_t>(unit.row_block) * A_BYTES;
    it.src.t_sw = t.sw;
    it.src.t_ext = t.ext;
    it.src.pad = 0;
    it.epi.t_cg = t.cg;
    it.epi.q_nrm = q.nrm + static_cast<size_t>(unit.row_block) * BM;
    it.ctl.tcols = img_used(t);
    it.ctl.live = unit.row_block * BM >= qused ? 0 : ((unit.row_block + 1) * BM < qused ? 2 : 1);
    it.ctl.unit = u;
    it.ctl.pad = 0;
    items[i] = it;
}

// 32-bit / 64-bit values that every lane of the warp holds identically (loaded from one address): the shuffle tells
// the compiler so, which lets it keep descriptors, addresses and loop state in UNIFORM registers.
__device__ __forceinline__ uint32_t uni(uint32_t v) { return __shfl_sync(0xffffffffu, v, 0); }
__device__ __forceinline__ int32_t uni(int32_t v) { return __shfl_sync(0xffffffffu, v, 0); }
__device__ __forceinline__ const uint8_t* uni(const uint8_t* p) {
    const unsigned long long v = reinterpret_cast<unsigned long long>(p);
    const uint32_t lo = __shfl_sync(0xffffffffu, static_cast<uint32_t>(v), 0);
    const uint32_t hi = __shfl_sync(0xffffffffu, static_cast<uint32_t>(v >> 32), 0);
    return reinterpret_cast<const uint8_t*>((static_cast<unsigned long long>(hi) << 32) | lo);
}

// Units come in pairs (u, u+1) of the same segment: the host makes every segment an even number of units.
// A pair is skipped by every role of both CTAs when its FIRST unit is all dead rows.
//
// The producer, relay and issuer loops are executed by ALL 32 lanes of their warp with warp-uniform control flow and
// operands; only the asynchronous instructions themselves (bulk copy, tcgen05.mma, tcgen05.commit, remote arrive) are
// predicated on one elected lane.  Written as `if (lane == 0) { whole loop }` the same loop compiles to a
// ELECT / 5 x R2UR.BROADCAST / BRA.U.ANY sequence around every UTCIMMA plus ~200 scalar instructions per tile, which one
// thread needs ~1400 cycles to get through — twice the 640 tensor cycles of a tile (tools/microbench_pair.cu,
// profiles/r01b_microbench_pair.log).
template <bool kDbg>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
match_pair_kernel(const Item* __restrict__ items, int num_items, int32_t* __restrict__ res_g,
                  int32_t* __restrict__ res_d1, int32_t* __restrict__ res_u, long long* __restrict__ dbg, int dbg_mode) {
    extern __shared__ uint8_t smem_raw[];
    // manual 1 KiB alignment (dynamic smem is only guaranteed 16-B aligned); both CTAs of the pair get the same offset
    const uint32_t raw_addr = ptx::smem_u32(smem_raw);
    uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);

    uint8_t* smA = smem + OFF_A;
    uint8_t* smAext = smem + OFF_AEXT;
    uint8_t* smB = smem + OFF_B;
    uint8_t* smE = smem + OFF_E;
    int4* smMerge = reinterpret_cast<int4*>(smem + OFF_MERGE);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
    uint64_t* a_full = bars;            // [2]
    uint64_t* a_empty = a_full + 2;     // [2]
    uint64_t* b_full = a_empty + 2;     // [NS]
    uint64_t* e_full = b_full + NS;     // [NE]
    uint64_t* done = e_full + NE;       // [ND]  tile `it` -> done[it % ND]
    uint64_t* t_empty = done + ND;      // [2]   used in the leader only
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + OFF_TMEMPTR);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = ptx::cluster_ctarank();          // 0 = leader
    const bool leader = rank == 0;

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
        const uint32_t full_count = leader ? 2u : 1u;       // own producer (+ the peer's relay)
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(&a_full[i], full_count);
            ptx::mbar_init(&a_empty[i], 1);
            ptx::mbar_init(&t_empty[i], EPI_WARPS);          // the 4 warps of the stage's warpgroup in each CTA
        }
        for (int i = 0; i < NS; ++i) ptx::mbar_init(&b_full[i], full_count);
        for (int i = 0; i < NE; ++i) ptx::mbar_init(&e_full[i], full_count);
        for (int i = 0; i < ND; ++i) ptx::mbar_init(&done[i], 1);
        ptx::fence_mbar_init();
    }
    if (warp == 2) ptx::tmem_alloc_pair<512>(tmem_ptr);
    ptx::tc_fence_before();
    ptx::cluster_sync();                 // barriers of BOTH CTAs are initialised before anyone arrives remotely
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    const int i_first = static_cast<int>(ptx::cluster_id_x());
    const int i_step = static_cast<int>(ptx::num_clusters_x());

    if (warp == 0) {
        // ===================================================================== producer (both CTAs, own halves)
        uint32_t it = 0, un = 0, et = 0;
        uint32_t ext_last0 = 0, ext_last1 = 0;                     // last tile (global counter) that reads each ext stage
        ItemSrc nsrc{};
        ItemCtl nctl{};
        if (i_first < num_items) { nsrc = load_part(&items[i_first].src); nctl = load_part(&items[i_first].ctl); }
        for (int i = i_first; i < num_items; i += i_step) {
            const uint8_t* q_rows = uni(nsrc.q_rows);
            const uint8_t* t_sw = uni(nsrc.t_sw);
            const uint8_t* t_ext = uni(nsrc.t_ext);
            const int tcols = uni(nctl.tcols), live = uni(nctl.live);
            if (i + i_step < num_items) {                          // in flight while this pair is fed
                nsrc = load_part(&items[i + i_step].src);
                nctl = load_part(&items[i + i_step].ctl);
            }
            // all-dead query rows: every role skips the pair.  An empty train image needs no operands either (only the
            // epilogue acts, writing "no match"): loading A for it would let this producer lap the peer's relay, which
            // is otherwise throttled through the issuer's a_full waits
            if (live == 0 || tcols == 0) continue;
            const uint32_t cu = un++;                              // index among the pairs this cluster really processes
            const uint32_t ab = cu & 1;
            ptx::mbar_wait(&a_empty[ab], ((cu >> 1) & 1) ^ 1);
            if (ptx::elect_one()) {
                ptx::mbar_arrive_expect_tx(&a_full[ab], A_BYTES);
                ptx::bulk_g2s(smA + ab * A_BYTES, q_rows + rank * A_BYTES, A_BYTES, &a_full[ab]);
            }
            __syncwarp();
            const int ntiles = (tcols + BN - 1) / BN;              // tiles beyond the occupied columns are skipped
            for (int tile = 0; tile < ntiles; ++tile, ++it) {
                if ((tile & 3) == 0) {
                    const uint32_t es = et & 1;
                    if (et >= NE) {                    // the stage's previous super-tile: wait for its last tile's MMAs
                        const uint32_t l = es ? ext_last1 : ext_last0;
                        ptx::mbar_wait(&done[l % ND], (l / ND) & 1);
                    }
                    const uint32_t last = it + static_cast<uint32_t>(min(4, ntiles - tile)) - 1;
                    if (es) ext_last1 = last; else ext_last0 = last;
                    if (ptx::elect_one()) {
                        ptx::mbar_arrive_expect_tx(&e_full[es], E_BYTES);
                        ptx::bulk_g2s(smE + es * E_BYTES, t_ext + (static_cast<size_t>(tile >> 2) * 2 + rank) * E_BYTES, E_BYTES,
                                      &e_full[es]);
                    }
                    __syncwarp();
                    ++et;
                }
                const uint32_t s = it % NS;
                if (it >= NS) {                        // the stage's previous tile, it - NS: wait for its MMAs
                    const uint32_t l = it - NS;
                    ptx::mbar_wait(&done[l % ND], (l / ND) & 1);
                }
                if (ptx::elect_one()) {
                    ptx::mbar_arrive_expect_tx(&b_full[s], B_BYTES);
                    ptx::bulk_g2s(smB + s * B_BYTES, t_sw + (static_cast<size_t>(tile) * 2 + rank) * B_BYTES, B_BYTES, &b_full[s]);
                }
                __syncwarp();
            }
        }
        // drain: every multicast commit aimed at this CTA's a_empty barriers has landed before the CTA may exit
        // (the epilogue waits for every done[] arrival)
        for (uint32_t i = 0; i < 2; ++i, ++un) ptx::mbar_wait(&a_empty[un & 1], ((un >> 1) & 1) ^ 1);
    } else if (warp == 1 && !leader) {
        // ===================================================================== relay (peer CTA): my half has landed
        const uint32_t a_full_leader[2] = {ptx::map_to_cta(ptx::smem_u32(&a_full[0]), 0), ptx::map_to_cta(ptx::smem_u32(&a_full[1]), 0)};
        const uint32_t e_full_leader[2] = {ptx::map_to_cta(ptx::smem_u32(&e_full[0]), 0), ptx::map_to_cta(ptx::smem_u32(&e_full[1]), 0)};
        const uint32_t b_full_leader0 = ptx::map_to_cta(ptx::smem_u32(&b_full[0]), 0);
        uint32_t it = 0, un = 0, et = 0;
        ItemCtl nctl{};
        if (i_first < num_items) nctl = load_part(&items[i_first].ctl);
        for (int i = i_first; i < num_items; i += i_step) {
            const int tcols = uni(nctl.tcols), live = uni(nctl.live);
            if (i + i_step < num_items) nctl = load_part(&items[i + i_step].ctl);
            if (live == 0 || tcols == 0) continue;
            const uint32_t cu = un++;
            const uint32_t ab = cu & 1;
            ptx::mbar_wait(&a_full[ab], (cu >> 1) & 1);
            if (ptx::elect_one()) ptx::mbar_arrive_cluster(a_full_leader[ab]);
            __syncwarp();
            const int ntiles = (tcols + BN - 1) / BN;
            for (int tile = 0; tile < ntiles; ++tile, ++it) {
                if ((tile & 3) == 0) {
                    const uint32_t es = et & 1;
                    ptx::mbar_wait(&e_full[es], (et >> 1) & 1);
                    if (ptx::elect_one()) ptx::mbar_arrive_cluster(e_full_leader[es]);
                    __syncwarp();
                    ++et;
                }
                const uint32_t s = it % NS;
                ptx::mbar_wait(&b_full[s], (it / NS) & 1);
                if (ptx::elect_one()) ptx::mbar_arrive_cluster(b_full_leader0 + s * 8);
                __syncwarp();
            }
        }
    } else if (warp == 1 && leader) {
        // ===================================================================== MMA issuer (leader CTA), in tile order
        constexpr uint32_t idesc = ptx::make_idesc_u8(2 * BM, BN);
        const uint32_t sbase = uni(ptx::smem_u32(smem));
        const uint32_t tbase = uni(tmem_base);
        const uint64_t aext_desc = ptx::make_smem_desc_sw128(sbase + OFF_AEXT);
        uint32_t it = 0, un = 0, et = 0;
        long long dbg_c0 = 0, dbg_t0 = 0;
        if (kDbg) {
            dbg_c0 = clock64();
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(dbg_t0));
        }
        ItemCtl nctl{};
        if (i_first < num_items) nctl = load_part(&items[i_first].ctl);
        for (int i = i_first; i < num_items; i += i_step) {
            const int tcols = uni(nctl.tcols), live = uni(nctl.live);
            if (i + i_step < num_items) nctl = load_part(&items[i + i_step].ctl);
            if (live == 0 || tcols == 0) continue;
            const uint32_t cu = un++;
            const uint32_t ab = cu & 1;
            const int ntiles = (tcols + BN - 1) / BN;
            ptx::mbar_wait(&a_full[ab], (cu >> 1) & 1);
            const uint64_t a_desc0 = ptx::make_smem_desc_sw128(sbase + OFF_A + ab * A_BYTES);
            for (int tile = 0; tile < ntiles; ++tile, ++it) {
                const uint32_t es = et & 1;
                const long long tsp = kDbg ? clock64() : 0;
                if ((tile & 3) == 0) ptx::mbar_wait(&e_full[es], (et >> 1) & 1);
                const uint32_t s = it % NS;
                const uint32_t acc = it & 1;
                const long long ts0 = kDbg ? clock64() : 0;
                ptx::mbar_wait2(&b_full[s], (it / NS) & 1, &t_empty[acc], ((it >> 1) & 1) ^ 1);
                ptx::tc_fence_after();
                const long long ts1 = kDbg ? clock64() : 0;
                const uint64_t b_desc0 = ptx::make_smem_desc_sw128(sbase + OFF_B + s * B_BYTES);
                // the tile's 32 extension bytes sit at K offset 32*(tile%4) of the super-tile rows
                const uint64_t e_desc = ptx::make_smem_desc_sw128(sbase + OFF_E + es * E_BYTES) + 2 * (tile & 3);
                const uint32_t d_tmem = tbase + acc * BN;
                if (ptx::elect_one()) {
#pragma unroll
                    for (int k = 0; k < KBYTES / UMMA_KB; ++k) {
                        // advancing K by 32 bytes inside the 128-B swizzle atom = +2 in the (addr >> 4) field
                        ptx::mma_i8_ss_pair(d_tmem, a_desc0 + 2 * k, b_desc0 + 2 * k, idesc, k > 0 ? 1u : 0u);
                    }
                    ptx::mma_i8_ss_pair(d_tmem, aext_desc, e_desc, idesc, 1u);     // + e_j
                    // ONE commit per tile: it frees the train-tile stage (producers), the extension stage when this was
                    // the super-tile's last tile (producers) and publishes the accumulator (epilogues)
                    ptx::mma_commit_pair(&done[it % ND], BOTH_CTAS);
                }
                __syncwarp();
                const long long ts2 = kDbg ? clock64() : 0;
                long long ts3 = 0;
                if (kDbg && dbg_mode == 5) {       // diagnostic: the issuer itself waits for the tile's MMAs (serialises the tiles)
                    ptx::mbar_wait(&done[it % ND], (it / ND) & 1);
                    ts3 = clock64();
                }
                if (kDbg && lane == 0 && ptx::cluster_id_x() == 0 && it >= 1024 && it < 1024 + 64) {
                    long long* d = dbg + 4 * 128 + (it - 1024) * 8;
                    d[0] = ts0; d[1] = ts1; d[2] = ts2; d[5] = tsp; d[6] = ts3;
                }
                if ((tile & 3) == 3 || tile == ntiles - 1) ++et;
            }
            if (ptx::elect_one()) ptx::mma_commit_pair(&a_empty[ab], BOTH_CTAS);
            __syncwarp();
        }
        if (kDbg && lane == 0) {       // diagnostic (MSFM_K1_DEBUG=1): issue-loop cycles, wall time, tiles of this cluster
            long long t1;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            long long* d = dbg + 4 * ptx::cluster_id_x();
            d[0] = clock64() - dbg_c0;
            d[1] = t1 - dbg_t0;
            d[2] = it;
            d[3] = un;
        }
    } else if (warp >= EPI_WARP0) {
        // ===================================================================== epilogue (both CTAs, own rows)
        const int wg = (warp - EPI_WARP0) >> 2;          // accumulator stage (tile parity) this warpgroup serves
        const int quarter = warp & 3;                    // TMEM lane quarter this warp may access
        const int row = quarter * 32 + lane;             // query row inside the unit
        const uint32_t t_empty_leader0 = ptx::map_to_cta(ptx::smem_u32(&t_empty[0]), 0);
        uint32_t it = 0, un = 0;
        long long ph_work = 0, ph_ld = 0, ph_all = 0, ph_z = 0, ph_t = kDbg ? clock64() : 0;     // kDbg: cycles per phase of this warp
        ItemEpi nepi{};
        ItemCtl nctl{};
        if (i_first < num_items) { nepi = load_part(&items[i_first].epi); nctl = load_part(&items[i_first].ctl); }
        for (int i = i_first; i < num_items; i += i_step) {
            const ItemEpi epi = nepi;
            const ItemCtl ctl = nctl;
            if (i + i_step < num_items) {
                nepi = load_part(&items[i + i_step].epi);
                nctl = load_part(&items[i + i_step].ctl);
            }
            if (ctl.live == 0) continue;
            const uint32_t cu = un++;
            const bool active = static_cast<int>(rank) < ctl.live;                // CTA-uniform; this CTA's unit is unit + rank
            const int tcols = ctl.tcols;
            const int ntiles = (tcols + BN - 1) / BN;
            // ||q_i||^2 (-1 for dead rows): requested now, needed when the pair is finished
            const int32_t ni = (active && wg == 0) ? __ldg(epi.q_nrm + rank * BM + row) : -1;
            int32_t k1 = kIntInf, k2 = kIntInf, g1 = 0;
            // C_g of the 8 groups of a tile, requested one own tile (= two tiles) ahead: an L2 round trip (~800 cycles) is
            // longer than a tile
            const int4* cg8 = reinterpret_cast<const int4*>(epi.t_cg);
            const int tfirst = ((it & 1) == static_cast<uint32_t>(wg)) ? 0 : 1;          // first tile of this pair in my stage
            int4 cg_next0 = make_int4(0, 0, 0, 0), cg_next1 = cg_next0;
            if (active && tfirst < ntiles) { cg_next0 = __ldg(cg8 + tfirst * 2); cg_next1 = __ldg(cg8 + tfirst * 2 + 1); }
            for (int tile = 0; tile < ntiles; ++tile, ++it) {
                const uint32_t acc = it & 1;
                if (acc != static_cast<uint32_t>(wg)) continue;       // the other warpgroup's accumulator stage
                if (!active || (kDbg && dbg_mode == 1)) {          // all-dead second unit: keep the barrier protocol going
                    ptx::mbar_wait(&done[it % ND], (it / ND) & 1);
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive_cluster(t_empty_leader0 + acc * 8);
                    continue;
                }
                const int4 cgv0 = cg_next0, cgv1 = cg_next1;
                if (tile + 2 < ntiles) { cg_next0 = __ldg(cg8 + (tile + 2) * 2); cg_next1 = __ldg(cg8 + (tile + 2) * 2 + 1); }
                ptx::mbar_wait(&done[it % ND], (it / ND) & 1);
                ptx::tc_fence_after();
                const long long te0 = kDbg ? clock64() : 0;
                if (kDbg) { ph_all += te0 - ph_t; ph_t = te0; }      // previous tile's integer work + the wait for this tile
                const uint32_t taddr0 = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BN;
                uint32_t v[4][32];
                // group maxima of the four 32-column groups in v, inserted into the running (best, best-of-other-groups);
                // all-dead groups (also the ones behind the occupied columns of the last tile) carry C_g = kDeadCg and
                // can never win.  Every max chain starts from `seed` (a neutral element).
                auto reduce1 = [&](uint32_t(&cur)[32], const int32_t cgc, const int gid, const int32_t seed) -> int32_t {
                    // max over the 32 columns of the group: two independent chains, two elements per VIMNMX3
                    int32_t m0 = __vimax3_s32(seed, static_cast<int32_t>(cur[0]), static_cast<int32_t>(cur[1]));
                    int32_t m1 = __vimax3_s32(seed, static_cast<int32_t>(cur[2]), static_cast<int32_t>(cur[3]));
#pragma unroll
                    for (int e = 4; e < 32; e += 4) {
                        m0 = __vimax3_s32(m0, static_cast<int32_t>(cur[e + 0]), static_cast<int32_t>(cur[e + 1]));
                        m1 = __vimax3_s32(m1, static_cast<int32_t>(cur[e + 2]), static_cast<int32_t>(cur[e + 3]));
                    }
                    const int32_t m = max(m0, m1);
                    if (kDbg && dbg_mode == 3) { k1 = min(k1, m); return m; }       // diagnostic: group maxima only
                    const int32_t key = cgc - 2 * m;      // = min over the group of (||d_j||^2 - 2 q.d_j)
                    k2 = min(k2, max(k1, key));
                    if (key < k1) g1 = gid;
                    k1 = min(k1, key);
                    return m;
                };
                auto cg_of = [](const int4 cgv, const int c) { return (c == 0) ? cgv.x : (c == 1) ? cgv.y : (c == 2) ? cgv.z : cgv.w; };
                // ---- columns 0..127 of the tile: read; reduce group by group and refill every group's registers with
                //      the group 128 columns further right as soon as it has been reduced, so that the second read is in
                //      flight under the first reduction and the stage can be handed back right after it
#pragma unroll
                for (int c = 0; c < 4; ++c) ptx::tmem_ld_32x32(taddr0 + c * 32, v[c]);
                ptx::tmem_ld_wait();
                // (the seed of a group is the previous group's maximum with the sign bit set — still below every accumulator
                // value, which are non-negative — so that ptxas cannot interleave the four groups and finish them together)
                int32_t seed = static_cast<int32_t>(0x80000000);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    if (!(kDbg && dbg_mode == 2)) seed = reduce1(v[c], cg_of(cgv0, c), tile * 8 + c, seed) | static_cast<int32_t>(0x80000000);
                    ptx::tmem_ld_32x32(taddr0 + 128 + c * 32, v[c]);
                }
                ptx::tmem_ld_wait();
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive_cluster(t_empty_leader0 + acc * 8);
                if (kDbg) { const long long t = clock64(); ph_ld += t - ph_t; ph_t = t; }
                if (kDbg && leader && quarter == 0 && lane == 0 && ptx::cluster_id_x() == 0 && it >= 1024 && it < 1024 + 64) {
                    long long* d = dbg + 4 * 128 + (it - 1024) * 8;
                    d[3] = te0; d[4] = clock64();
                }
                if (kDbg && dbg_mode == 2) { k1 = min(k1, static_cast<int32_t>(v[0][0] ^ v[1][1] ^ v[2][2] ^ v[3][3])); continue; }
                // The second reduction must not start before the stage has been handed back: ptxas is free to sink the arrive
                // (nothing depends on it) under the ALU instructions and did so.  The chains are therefore seeded with a
                // neutral element (INT_MIN or INT_MIN + 1) derived from a clock read that stays behind the arrive.
                uint32_t clk;
                asm volatile("mov.u32 %0, %%clock;" : "=r"(clk) : : "memory");
                const int32_t z = static_cast<int32_t>(0x80000000u | (clk & 1u));
                if (kDbg) {      // latency of the seed load (shared memory is busy with tensor-core operand reads)
                    long long t;
                    asm volatile("mov.u64 %0, %%clock64;" : "=l"(t) : "r"(z) : "memory");
                    ph_z += t - ph_t;
                }
#pragma unroll
                for (int c = 0; c < 4; ++c) reduce1(v[c], cg_of(cgv1, c), tile * 8 + 4 + c, z);
                if (kDbg) {      // k1, k2 are inputs so that the stamp cannot be hoisted above the integer work
                    long long t;
                    asm volatile("mov.u64 %0, %%clock64;" : "=l"(t) : "r"(k1), "r"(k2) : "memory");
                    ph_work += t - ph_t;
                }
            }
            if (!active) continue;
            // ---- unit end: fold the results of the two warpgroups (even / odd tiles), write the row results
            int4* mbuf = smMerge + (cu & 1) * 128;
            if (wg == 1) mbuf[row] = make_int4(k1, k2, g1, 0);
            ptx::bar_sync(1, EPI_THREADS);
            if (wg == 0) {
                const int4 o = mbuf[row];
                k2 = __vimin3_s32(k2, o.y, max(k1, o.x));
                if (o.x < k1) g1 = o.z;
                k1 = min(k1, o.x);
                const size_t out = static_cast<size_t>(ctl.unit + static_cast<int>(rank)) * BM + row;
                const bool have1 = ni >= 0 && k1 < kDeadKey;
                const bool have2 = have1 && k2 < kDeadKey;
                res_g[out] = have1 ? g1 : -1;
                res_d1[out] = have1 ? (k1 + ni) : kIntInf;
                res_u[out] = have2 ? (k2 + ni) : kIntInf;
            }
        }
        if (kDbg && lane == 0 && ptx::cluster_id_x() == 0) {
            long long* d = dbg + 4 * 128 + 64 * 8 + (rank * 8 + (warp - EPI_WARP0)) * 4;
            d[0] = ph_ld; d[1] = ph_work; d[2] = ph_all; d[3] = it / 2; d[64] = ph_z;
        }
    }

    // ---- teardown: nobody leaves (and TMEM is not freed) while the peer may still signal or compute
    ptx::tc_fence_before();
    ptx::cluster_sync();
    if (warp == 2) ptx::tmem_dealloc_pair<512>(tmem_base);
}

}  // namespace k1

// host launcher (called from msfm_api.cu).  unit0 and num_units must be even (segments are an even number of units).
// item_scratch: device buffer of at least match_tile_item_bytes(num_units) bytes.
size_t match_tile_item_bytes(int num_units) { return static_cast<size_t>(num_units / 2 + 1) * sizeof(k1::Item); }

cudaError_t launch_match_tile_kernel(const ImgDev* imgs, const UnitDev* units, int unit0, int num_units, int32_t* res_g,
                                     int32_t* res_d1, int32_t* res_u, void* item_scratch, int num_sms, cudaStream_t stream) {
    static const int debug = [] {
        const char* e = std::getenv("MSFM_K1_DEBUG");       // 1: timeline + cycles; 11: no TMEM reads, 12: no integer work, 13: group maxima only (results garbage)
        return e ? atoi(e) : 0;
    }();
    auto kernel = debug ? k1::match_pair_kernel<true> : k1::match_pair_kernel<false>;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(k1::SMEM_BYTES));
    if (e != cudaSuccess) return e;
    if (num_units <= 0) return cudaSuccess;
    if ((unit0 | num_units) & 1) return cudaErrorInvalidValue;
    const int num_items = num_units / 2;
    const int clusters = num_items < num_sms / 2 ? num_items : num_sms / 2;
    if (clusters <= 0) return cudaErrorInvalidConfiguration;
    k1::Item* items = static_cast<k1::Item*>(item_scratch);
    { k1::build_items_kernel<<<(num_items + 255) / 256, 256, 0, stream>>>(imgs, units, unit0, num_items, items); MSFM_COUNT_LAUNCH(); }
    long long* dbg = nullptr;
    if (debug) {
        static long long* d_dbg = nullptr;
        if (!d_dbg) cudaMalloc(&d_dbg, (4 * 128 + 64 * 8 + 128) * sizeof(long long));
        dbg = d_dbg;
        cudaMemsetAsync(dbg, 0, (4 * 128 + 64 * 8 + 128) * sizeof(long long), stream);
    }
    { kernel<<<2 * clusters, k1::NUM_THREADS, k1::SMEM_BYTES, stream>>>(items, num_items, res_g, res_d1, res_u, dbg, debug > 10 ? debug - 10 : 0); MSFM_COUNT_LAUNCH(); }
    if (debug) {
        long long h[4 * 128 + 64 * 8 + 128];
        cudaMemcpyAsync(h, dbg, sizeof(h), cudaMemcpyDeviceToHost, stream);
        cudaStreamSynchronize(stream);
        double cyc = 0, ns = 0, tiles = 0;
        int n = 0;
        for (int i = 0; i < clusters; ++i)
            if (h[4 * i + 2] > 0) { cyc += double(h[4 * i]); ns += double(h[4 * i + 1]); tiles += double(h[4 * i + 2]); ++n; }
        {   // per-tile timeline of cluster 0, tiles 1024..1087 (clock64 of the leader SM)
            const long long* ts = h + 4 * 128;
            double w_te = 0, issue = 0, exec = 0, epi = 0, back = 0;
            int m = 0;
            for (int k = 2; k < 62; ++k) {
                const long long* a = ts + 8 * k;
                const long long* nx = ts + 8 * (k + 2);      // same accumulator stage, two tiles later
                if (!a[2] || !a[3] || !nx[1]) continue;
                w_te += double(a[1] - a[0]);     // issuer: b_full satisfied -> t_empty satisfied
                issue += double(a[2] - a[1]);    // issuer: 5 MMAs + commit issued
                exec += double(a[3] - a[2]);     // commit issued -> epilogue sees done
                epi += double(a[4] - a[3]);      // epilogue: TMEM -> registers, arrive sent
                back += double(nx[1] - a[4]);    // arrive sent -> issuer of tile+2 sees t_empty
                ++m;
            }
            if (m && std::getenv("MSFM_K1_DEBUG_RAW")) {      // absolute per-tile stamps relative to tile 1026's b_full
                const long long base = ts[8 * 2];
                for (int k = 2; k < 18; ++k) {
                    const long long* a = ts + 8 * k;
                    fprintf(stderr, "K1 raw tile %d (stage %d): top %lld b_full %lld t_empty %lld committed %lld | epi done-seen %lld released %lld\n",
                            k, k & 1, a[5] - base, a[0] - base, a[1] - base, a[2] - base, a[3] - base, a[4] - base);
                }
            }
            if (std::getenv("MSFM_K1_DEBUG_RAW")) {
                const long long* w = h + 4 * 128 + 64 * 8;
                for (int k = 0; k < 16; ++k)
                    if (w[4 * k + 3] > 0)
                        fprintf(stderr, "K1 epilogue warp cta%d w%d (subpartition %d, wg %d): per own tile  done -> release %.0f | seed %.0f | integer work after release %.0f | work + wait for next tile %.0f\n",
                                k / 8, 4 + k % 8, k % 4, (k % 8) / 4, double(w[4 * k]) / w[4 * k + 3], double(w[4 * k + 64]) / w[4 * k + 3], double(w[4 * k + 1]) / w[4 * k + 3], double(w[4 * k + 2]) / w[4 * k + 3]);
            }
            {
                double own = 0; int mo = 0;
                for (int k = 2; k < 62; ++k) { const long long* a = ts + 8 * k; if (a[6] && a[2]) { own += double(a[6] - a[2]); ++mo; } }
                if (mo) fprintf(stderr, "K1 issuer: commit issued -> own observation of the tile's completion: %.0f cycles (avg of %d)\n", own / mo, mo);
            }
            if (m) fprintf(stderr, "K1 timeline (avg of %d tiles): wait_t_empty %.0f | issue %.0f | commit->done %.0f | epi ld+arrive %.0f | arrive->issuer %.0f\n",
                           m, w_te / m, issue / m, exec / m, epi / m, back / m);
        }
        if (n) fprintf(stderr, "K1 debug: %d clusters, %.0f tiles/cluster, %.1f cycles/tile, %.3f ms, %.0f MHz\n", n, tiles / n,
                       cyc / tiles, ns / n * 1e-6, cyc / ns * 1e3);
    }
    return cudaGetLastError();
}

}  // namespace msfm
