// Micro-benchmarks of the CTA-pair (tcgen05 cta_group::2) issue / completion / release loop that bounds K1.
// One cluster of two CTAs per SM pair; the leader issues M=256 x N x K=32 i8 MMAs on garbage operands.
// Build: make microbench_pair ; run on the GPU box.  Prints cycles per "tile" (5 MMAs).
//
//   ISSUERS   1 or 2 issuing threads (warps 1, 2 of the leader); with 2, issuer w owns the tiles with (tile & 1) == w
//   STAGES    accumulator stages (2 x N=256 or 4 x N=128 columns)
//   EPI       0: no consumer at all (free-running MMAs, one commit at the end)
//             1: one commit per tile, nobody waits for it
//             2: + "epilogue" warps (8 per CTA) wait for the tile's commit and arrive (remotely for the peer) on the
//                leader's t_empty; issuer waits t_empty before reusing a stage
//             3: as 2, and the epilogue warps really read their 32 x (N/2) accumulator block from TMEM first
//             4: as 3, + ~the integer work of the K1 epilogue (one VIMNMX3 per two elements)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../monocularsfm_b200/csrc/ptx.cuh"
using namespace msfm;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 1; } } while (0)

constexpr int ND = 8;     // ring of done barriers

template <int N, int ISSUERS, int EPI, int DESC = 0, int UNI = 0, int TS = 0>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(384, 1)
k_pair(int iters, long long* cyc, long long* stamps, int* sink) {
    // TS: the A operand (128 rows x 160 K-bytes = 40 columns, padded to 48) lives in tensor memory behind the accumulators
    constexpr int STAGES = TS ? (512 - 48) / N : 512 / N;
    constexpr uint32_t A_TMEM_COL = STAGES * N;
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint32_t tptr;
    __shared__ uint64_t done[ND], t_empty[STAGES], fin;
    const uint32_t raw = ptx::smem_u32(smem_raw);
    uint8_t* smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = ptx::cluster_ctarank();
    for (int i = threadIdx.x; i < (8 * 16384) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = i * 2654435761u;
    if (threadIdx.x == 0) {
        for (int i = 0; i < ND; ++i) ptx::mbar_init(&done[i], 1);
        for (int i = 0; i < STAGES; ++i) ptx::mbar_init(&t_empty[i], 16);
        ptx::mbar_init(&fin, ISSUERS);
        ptx::fence_mbar_init();
    }
    ptx::fence_proxy_async();
    if (warp == 2) ptx::tmem_alloc_pair<512>(&tptr);
    ptx::tc_fence_before();
    ptx::cluster_sync();
    ptx::tc_fence_after();
    const uint32_t tmem_base = tptr;
    int acc_sink = 0;
    if (UNI && rank == 0 && warp >= 1 && warp <= ISSUERS) {
        // whole warp runs the loop (warp-uniform control flow and operands -> descriptors live in uniform registers);
        // only the tcgen05 instructions are predicated on one elected lane
        const uint32_t me = __shfl_sync(0xffffffffu, warp - 1, 0);
        const uint32_t sbase = __shfl_sync(0xffffffffu, ptx::smem_u32(smem), 0);
        const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
        constexpr uint32_t idesc = ptx::make_idesc_u8(256, N);
        const long long t0 = clock64();
        uint32_t st = 0, ph = 1, s6 = 0, dn = 0;
        for (int it = 0; it < iters; ++it) {
            if (ISSUERS == 1 || (it & 1) == static_cast<int>(me)) {
                if (EPI >= 2 && EPI <= 5) {
                    ptx::mbar_wait(&t_empty[st], ph);
                    ptx::tc_fence_after();
                }
                const uint64_t adx = ptx::make_smem_desc_sw128(sbase + (DESC ? ((it / 33) & 1) * 16384 : 0));
                const uint64_t bdx = ptx::make_smem_desc_sw128(sbase + (DESC ? 32768 + s6 * 16384 : 16384));
                if (ptx::elect_one()) {
#pragma unroll
                    for (int k = 0; k < 5; ++k) {
                        if (TS) ptx::mma_i8_ts_pair(tb + st * N, tb + A_TMEM_COL + 8 * k, bdx + 2 * (k & 3), idesc, k > 0);
                        else ptx::mma_i8_ss_pair(tb + st * N, adx + 2 * (k & 3), bdx + 2 * (k & 3), idesc, k > 0);
                    }
                    if (EPI >= 1) ptx::mma_commit_pair(&done[dn], 0b11);
                }
                __syncwarp();
            }
            if (++st == STAGES) { st = 0; ph ^= 1; }
            if (++s6 == 6) s6 = 0;
            if (++dn == ND) dn = 0;
        }
        if (ptx::elect_one()) ptx::mma_commit_pair(&fin, 0b11);
        __syncwarp();
        ptx::mbar_wait(&fin, 0);
        if (me == 0 && lane == 0) cyc[blockIdx.x >> 1] = clock64() - t0;
    } else if (rank == 0 && warp >= 1 && warp <= ISSUERS) {
        if (lane == 0) {
            const uint32_t me = warp - 1;
            const uint64_t ad = ptx::make_smem_desc_sw128(ptx::smem_u32(smem));
            const uint64_t bd = ptx::make_smem_desc_sw128(ptx::smem_u32(smem + 16384));
            constexpr uint32_t idesc = ptx::make_idesc_u8(256, N);
            const long long t0 = clock64();
            for (int it = 0; it < iters; ++it) {
                if (it % ISSUERS != static_cast<int>(me)) continue;
                const uint32_t st = it % STAGES;
                long long s0 = 0, s1 = 0;
                if (EPI >= 2) {
                    s0 = clock64();
                    ptx::mbar_wait(&t_empty[st], ((it / STAGES) & 1) ^ 1);
                    ptx::tc_fence_after();
                    s1 = clock64();
                }
                long long m[6];
                // DESC: per-tile descriptors as in the real kernel (6-stage B ring, A alternating per 33 tiles)
                const uint64_t adx = DESC ? ptx::make_smem_desc_sw128(ptx::smem_u32(smem + ((it / 33) & 1) * 16384)) : ad;
                const uint64_t bdx = DESC ? ptx::make_smem_desc_sw128(ptx::smem_u32(smem + 32768 + (it % 6) * 16384)) : bd;
#pragma unroll
                for (int k = 0; k < 5; ++k) {
                    if (stamps) m[k] = clock64();
                    ptx::mma_i8_ss_pair(tmem_base + st * N, adx + 2 * (k & 3), bdx + 2 * (k & 3), idesc, k > 0);
                }
                if (stamps) m[5] = clock64();
                if (EPI >= 1) ptx::mma_commit_pair(&done[it % ND], 0b11);
                if (stamps && blockIdx.x == 0 && it >= 512 && it < 512 + 32) {
                    long long* d = stamps + (it - 512) * 12;
                    d[0] = s0; d[1] = s1; d[2] = clock64();
#pragma unroll
                    for (int k = 0; k < 6; ++k) d[4 + k] = m[k];
                }
            }
            ptx::mma_commit_pair(&fin, 0b11);
            ptx::mbar_wait(&fin, 0);
            if (me == 0) cyc[blockIdx.x >> 1] = clock64() - t0;
        }
    } else if (warp >= 4 && EPI == 6) {
        // pure integer load on the epilogue warps (register-only VIMNMX3 chains), no TMEM reads, no barriers: does ALU
        // activity by itself slow the free-running tensor pipe?
        int a[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) a[k] = threadIdx.x * 7 + k;
        int x = threadIdx.x, y = lane;
        for (int i = 0; i < iters * 4; ++i) {
#pragma unroll
            for (int k = 0; k < 16; ++k) a[k] = __vimax3_s32(a[k], x + k, y - k);
            x ^= i; y += a[3] & 1;
        }
#pragma unroll
        for (int k = 0; k < 16; ++k) acc_sink ^= a[k];
    } else if (warp >= 4 && EPI >= 2) {
        const int quarter = warp & 3, wg = (warp - 4) >> 2;
        long long ph_ld = 0, ph_work = 0, ph_all = 0, ph_t = clock64();
        __shared__ int zneg;
        if (threadIdx.x == 128) *reinterpret_cast<volatile int*>(&zneg) = static_cast<int>(0x80000000);
        const uint32_t zaddr = ptx::smem_u32(&zneg);
        int k1 = 0x7fffffff, k2 = 0x7fffffff, g1 = 0;
        const uint32_t te[4] = {ptx::map_to_cta(ptx::smem_u32(&t_empty[0]), 0), ptx::map_to_cta(ptx::smem_u32(&t_empty[1 % STAGES]), 0),
                                ptx::map_to_cta(ptx::smem_u32(&t_empty[2 % STAGES]), 0), ptx::map_to_cta(ptx::smem_u32(&t_empty[3 % STAGES]), 0)};
        for (int it = 0; it < iters; ++it) {
            const uint32_t st = it % STAGES;
            ptx::mbar_wait(&done[it % ND], (it / ND) & 1);
            ptx::tc_fence_after();
            const long long e0 = clock64();
            ph_all += e0 - ph_t; ph_t = e0;
            if (EPI >= 3) {
                constexpr int NCH = N / 64;          // 32-column chunks of this warpgroup's half
                uint32_t v[NCH][32];
                const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + st * N + wg * (N / 2);
#pragma unroll
                for (int c = 0; c < NCH; ++c) ptx::tmem_ld_32x32(taddr + c * 32, v[c]);
                ptx::tmem_ld_wait();
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive_cluster(te[st]);
                { const long long t = clock64(); ph_ld += t - ph_t; ph_t = t; }
                if (stamps && rank == 0 && warp == 4 && lane == 0 && blockIdx.x == 0 && it >= 512 && it < 512 + 32) {
                    long long* d = stamps + (it - 512) * 12;
                    d[10] = e0; d[11] = clock64();
                }
                if (EPI == 5 || EPI == 7 || EPI == 8 || EPI == 9) {   // the arithmetic of the real K1 epilogue (N = 256 only)
                    // 5: seed from a volatile shared load + key/insert | 7: constant seed | 8: shared seed, group maxima only
                    // 9: seed from a clock read
                    int z = static_cast<int>(0x80000000);
                    if (EPI == 5 || EPI == 8) asm volatile("ld.volatile.shared.s32 %0, [%1];" : "=r"(z) : "r"(zaddr) : "memory");
                    if (EPI == 9) { uint32_t clk; asm volatile("mov.u32 %0, %%clock;" : "=r"(clk) : : "memory"); z = static_cast<int>(0x80000000u | (clk & 1u)); }
#pragma unroll
                    for (int c = 0; c < NCH; ++c) {
                        int m0 = __vimax3_s32(z, static_cast<int>(v[c][0]), static_cast<int>(v[c][1]));
                        int m1 = __vimax3_s32(z, static_cast<int>(v[c][2]), static_cast<int>(v[c][3]));
#pragma unroll
                        for (int e = 4; e < 32; e += 4) {
                            m0 = __vimax3_s32(m0, static_cast<int>(v[c][e]), static_cast<int>(v[c][e + 1]));
                            m1 = __vimax3_s32(m1, static_cast<int>(v[c][e + 2]), static_cast<int>(v[c][e + 3]));
                        }
                        const int m = max(m0, m1);
                        if (EPI == 8) { k1 = min(k1, m); continue; }
                        const int key = (it * 977 + c * 31) - 2 * m;
                        k2 = min(k2, max(k1, key));
                        if (key < k1) g1 = it * 8 + wg * 4 + c;
                        k1 = min(k1, key);
                    }
                    long long t;
                    asm volatile("mov.u64 %0, %%clock64;" : "=l"(t) : "r"(k1), "r"(k2) : "memory");
                    ph_work += t - ph_t;
                    acc_sink = k1 ^ k2 ^ g1;
                } else
                if (EPI >= 4) {
#pragma unroll
                    for (int c = 0; c < NCH; ++c) {
                        int m0 = v[c][0], m1 = v[c][1];
#pragma unroll
                        for (int e = 2; e < 32; e += 4) {
                            m0 = __vimax3_s32(m0, static_cast<int>(v[c][e]), static_cast<int>(v[c][e + 1]));
                            m1 = __vimax3_s32(m1, static_cast<int>(v[c][e + 2 < 32 ? e + 2 : 0]), static_cast<int>(v[c][e + 3 < 32 ? e + 3 : 1]));
                        }
                        acc_sink = max(acc_sink, max(m0, m1) - acc_sink / 3);
                    }
                } else {
#pragma unroll
                    for (int c = 0; c < NCH; ++c) acc_sink ^= v[c][c];
                }
            } else {
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive_cluster(te[st]);
                if (stamps && rank == 0 && warp == 4 && lane == 0 && blockIdx.x == 0 && it >= 512 && it < 512 + 32) {
                    long long* d = stamps + (it - 512) * 12;
                    d[10] = e0; d[11] = clock64();
                }
            }
        }
        if (stamps && blockIdx.x < 2 && lane == 0) {
            long long* d = stamps + 32 * 12 + (rank * 8 + (warp - 4)) * 4;
            d[0] = ph_ld; d[1] = ph_work; d[2] = ph_all; d[3] = iters;
        }
    }
    if (acc_sink == 0x7654321) sink[0] = acc_sink;
    ptx::tc_fence_before();
    ptx::cluster_sync();
    if (warp == 2) ptx::tmem_dealloc_pair<512>(tmem_base);
}

template <int N, int ISSUERS, int EPI, int DESC = 0, int UNI = 0, int TS = 0>
static int run(int G, long long* d_cyc, long long* d_stamps, int* d_sink, const char* name, bool print_stamps) {
    auto k = k_pair<N, ISSUERS, EPI, DESC, UNI, TS>;
    if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 140000) != cudaSuccess) return 1;
    const int iters = 4096;
    cudaMemset(d_stamps, 0, (32 * 12 + 64) * 8);
    for (int rep = 0; rep < 2; ++rep) k<<<G, 384, 140000>>>(iters, d_cyc, print_stamps ? d_stamps : nullptr, d_sink);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("%s failed: %s\n", name, cudaGetErrorString(cudaGetLastError())); return 1; }
    static long long h[512];
    cudaMemcpy(h, d_cyc, (G / 2) * 8, cudaMemcpyDeviceToHost);
    double s = 0; for (int i = 0; i < G / 2; ++i) s += h[i];
    printf("pair %-52s: %7.1f cyc per tile (5 MMAs 256x%dx32, floor %d)\n", name, s / (G / 2) / iters, N, 5 * N / 2);
    if (print_stamps) {
        static long long st[32 * 12 + 64];
        cudaMemcpy(st, d_stamps, sizeof(st), cudaMemcpyDeviceToHost);
        for (int k = 0; k < 16; k += 5)
            if (st[32 * 12 + 4 * k + 3] > 0)
                printf("   epilogue warp cta%d w%d: per tile  TMEM read+release %.0f | integer work %.0f | work + wait %.0f\n", k / 8, 4 + k % 8,
                       double(st[32 * 12 + 4 * k]) / iters, double(st[32 * 12 + 4 * k + 1]) / iters, double(st[32 * 12 + 4 * k + 2]) / iters);
        const long long base = st[4];
        for (int t = 0; t < 12; ++t) {
            const long long* d = st + t * 12;
            printf("   tile %2d: wait %6lld got %6lld | mma issue at %6lld %6lld %6lld %6lld %6lld end %6lld | commit %6lld | epi seen %6lld released %6lld\n", t,
                   d[0] ? d[0] - base : 0, d[1] ? d[1] - base : 0, d[4] - base, d[5] - base, d[6] - base, d[7] - base, d[8] - base, d[9] - base, d[2] - base,
                   d[10] ? d[10] - base : 0, d[11] ? d[11] - base : 0);
        }
    }
    return 0;
}

int main() {
    CK(cudaSetDevice(0));
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    printf("device %s SMs=%d\n", p.name, p.multiProcessorCount);
    const int G = p.multiProcessorCount / 2 * 2;
    long long *d_cyc, *d_stamps; int* d_sink;
    CK(cudaMalloc(&d_cyc, 512 * 8)); CK(cudaMalloc(&d_stamps, (32 * 12 + 64) * 8)); CK(cudaMalloc(&d_sink, 64));
    run<256, 1, 0>(G, d_cyc, d_stamps, d_sink, "N=256 1 issuer, free-running", true);
    run<256, 1, 0>(G, d_cyc, d_stamps, d_sink, "N=256 1 issuer, free-running (no stamps)", false);
    run<256, 2, 0>(G, d_cyc, d_stamps, d_sink, "N=256 2 issuers, free-running", false);
    run<256, 1, 1>(G, d_cyc, d_stamps, d_sink, "N=256 1 issuer, commit per tile", false);
    run<256, 1, 2>(G, d_cyc, d_stamps, d_sink, "N=256 1 issuer, 2 stages, empty epilogue", true);
    run<256, 2, 2>(G, d_cyc, d_stamps, d_sink, "N=256 2 issuers, 2 stages, empty epilogue", false);
    run<256, 1, 3>(G, d_cyc, d_stamps, d_sink, "N=256 1 issuer, 2 stages, epilogue reads TMEM", true);
    run<256, 2, 3>(G, d_cyc, d_stamps, d_sink, "N=256 2 issuers, 2 stages, epilogue reads TMEM", false);
    run<256, 1, 4>(G, d_cyc, d_stamps, d_sink, "N=256 1 issuer, 2 stages, epilogue reads + max", false);
    run<256, 2, 4>(G, d_cyc, d_stamps, d_sink, "N=256 2 issuers, 2 stages, epilogue reads + max", false);
    run<256, 1, 2, 1>(G, d_cyc, d_stamps, d_sink, "N=256 1 issuer, empty epilogue, per-tile descriptors", true);
    run<256, 2, 2, 1>(G, d_cyc, d_stamps, d_sink, "N=256 2 issuers, empty epilogue, per-tile descriptors", false);
    run<256, 2, 3, 1>(G, d_cyc, d_stamps, d_sink, "N=256 2 issuers, epilogue reads TMEM, per-tile descriptors", true);
    run<256, 2, 4, 1>(G, d_cyc, d_stamps, d_sink, "N=256 2 issuers, epilogue reads + max, per-tile descriptors", true);
    run<256, 1, 2, 1, 1>(G, d_cyc, d_stamps, d_sink, "UNIFORM N=256 1 issuer, empty epilogue, per-tile desc", false);
    run<256, 1, 3, 1, 1>(G, d_cyc, d_stamps, d_sink, "UNIFORM N=256 1 issuer, epilogue reads TMEM, per-tile desc", false);
    run<256, 1, 4, 1, 1>(G, d_cyc, d_stamps, d_sink, "UNIFORM N=256 1 issuer, epilogue reads + max, per-tile desc", false);
    run<256, 2, 4, 1, 1>(G, d_cyc, d_stamps, d_sink, "UNIFORM N=256 2 issuers, epilogue reads + max, per-tile desc", false);
    run<256, 1, 5, 1, 1>(G, d_cyc, d_stamps, d_sink, "UNIFORM N=256 1 issuer, REAL epilogue arithmetic", true);
    run<256, 1, 1, 1, 1>(G, d_cyc, d_stamps, d_sink, "UNIFORM N=256 1 issuer, free-running, idle epilogue warps", false);
    run<256, 1, 6, 1, 1>(G, d_cyc, d_stamps, d_sink, "UNIFORM N=256 1 issuer, free-running, ALU-busy epilogue warps", false);
    run<128, 1, 3, 1, 1>(G, d_cyc, d_stamps, d_sink, "UNIFORM N=128 1 issuer, 4 stages, epilogue reads TMEM", false);
    run<128, 1, 4, 1, 1>(G, d_cyc, d_stamps, d_sink, "UNIFORM N=128 1 issuer, 4 stages, epilogue reads + max", false);
    run<128, 2, 4, 1, 1>(G, d_cyc, d_stamps, d_sink, "UNIFORM N=128 2 issuers, 4 stages, epilogue reads + max", false);
    // round-2 candidates: query operand in tensor memory, which frees 20 KB of shared-memory reads per
    // tile and, with 192-column tiles, leaves room for the A operand next to two accumulator stages (or three of 128)
    run<256, 1, 1, 1, 1, 1>(G, d_cyc, d_stamps, d_sink, "UNIFORM TS N=256 1 issuer, free-running (1 stage + A)", false);
    run<192, 1, 1, 1, 1, 1>(G, d_cyc, d_stamps, d_sink, "UNIFORM TS N=192 1 issuer, free-running", false);
    run<192, 1, 5, 1, 1, 1>(G, d_cyc, d_stamps, d_sink, "UNIFORM TS N=192 1 issuer, 2 stages, REAL epilogue arithmetic", true);
    run<128, 1, 3, 1, 1, 1>(G, d_cyc, d_stamps, d_sink, "UNIFORM TS N=128 1 issuer, 3 stages, epilogue reads TMEM", false);
    run<128, 1, 0>(G, d_cyc, d_stamps, d_sink, "N=128 1 issuer, free-running", false);
    run<128, 1, 2>(G, d_cyc, d_stamps, d_sink, "N=128 1 issuer, 4 stages, empty epilogue", false);
    run<128, 1, 3>(G, d_cyc, d_stamps, d_sink, "N=128 1 issuer, 4 stages, epilogue reads TMEM", true);
    run<128, 2, 3>(G, d_cyc, d_stamps, d_sink, "N=128 2 issuers, 4 stages, epilogue reads TMEM", false);
    run<128, 1, 4>(G, d_cyc, d_stamps, d_sink, "N=128 1 issuer, 4 stages, epilogue reads + max", false);
    run<128, 2, 4>(G, d_cyc, d_stamps, d_sink, "N=128 2 issuers, 4 stages, epilogue reads + max", false);
    printf("MICROBENCH PAIR DONE\n");
    return 0;
}
