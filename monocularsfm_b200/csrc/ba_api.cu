// C-ABI entry points of the B-path (include/msfm_b200.h): problem residency, evaluate / linearize for parity,
// the Levenberg-Marquardt loop that replaces ceres::Solve (src/Optimizer/CeresBundleOptimizer.cpp:293), and the
// NCCL communicator (loaded with dlopen so the library has no link-time NCCL dependency).
#include <cusolverDn.h>
#include <dlfcn.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstring>
#include <new>
#include <vector>

#include "ba_tiles.hpp"
#include "ba_types.cuh"
#include "ctx.hpp"

namespace msfm {
using ba::CamPre;
using ba::Problem;
using ba::TailLayout;
using ba::Tile;
using ba::Item;
cudaError_t ba_launch_cam_prep(const double*, int, CamPre*, cudaStream_t);
cudaError_t ba_launch_evaluate(const Problem&, double*, float*, double*, int, cudaStream_t);
cudaError_t ba_launch_linearize(const Problem&, double, int, cudaStream_t);
cudaError_t ba_launch_expand_dense(const Problem&, double, double*, cudaStream_t);
cudaError_t ba_launch_track_errors(const Problem&, double*, int, cudaStream_t);
cudaError_t ba_launch_backsub(const Problem&, double, const double*, double*, double*, int, cudaStream_t);
cudaError_t ba_launch_update_cams(const double*, const int32_t*, int, const double*, double*, cudaStream_t);
cudaError_t ba_launch_copy(const double*, double*, int, cudaStream_t);
cudaError_t ba_launch_permute_obs(int, const int32_t*, const double*, const int32_t*, double*, int32_t*, cudaStream_t);
cudaError_t ba_launch_fill_obs_pt(int, const int32_t*, const int32_t*, int32_t*, cudaStream_t);
cudaError_t ba_launch_lm_record(const Problem&, const double*, const double*, const double*, const int*, int, int, double*, cudaStream_t);
cudaError_t ba_launch_filter_stats(const Problem&, double, uint8_t*, double*, int32_t*, double*, int, cudaStream_t);
size_t ba_fused_smem_bytes(bool);
namespace ba { struct TridiagSolver; }
ba::TridiagSolver* tridiag_create(int, const std::vector<int32_t>&, const std::vector<int32_t>&, cusolverDnHandle_t, cudaStream_t, cudaError_t*);
void tridiag_destroy(ba::TridiagSolver*);
void tridiag_info(const ba::TridiagSolver*, int32_t[4]);
int* tridiag_dev_info(ba::TridiagSolver*);
int tridiag_n_super(const ba::TridiagSolver*);
cudaError_t tridiag_factor_solve(ba::TridiagSolver*, const Problem&, double, cusolverDnHandle_t, double*, int, cudaStream_t);
namespace band { struct BandSolver; }
band::BandSolver* band_create(int, const std::vector<int32_t>&, const std::vector<int32_t>&, int, cudaError_t*);
void band_destroy(band::BandSolver*);
void band_info(const band::BandSolver*, int32_t[4]);
int32_t* band_dev_info(band::BandSolver*);
cudaError_t band_factor_solve(band::BandSolver*, const Problem&, double, double*, int, cudaStream_t);
}  // namespace msfm
using namespace msfm;

// ------------------------------------------------------------------------------------------------ NCCL via dlopen
struct Id128 { char b[128]; };   // ncclUniqueId is passed BY VALUE (128 bytes)
struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(void*) = nullptr;
    int (*CommInitRank)(void**, int, Id128, int) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
static NcclApi g_nccl;
static const char* nccl_load() {
    if (g_nccl.lib) return nullptr;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
        g_nccl.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl.lib) break;
    }
    if (!g_nccl.lib) return "libnccl.so.2 not found (dlopen)";
    g_nccl.GetUniqueId = reinterpret_cast<int (*)(void*)>(dlsym(g_nccl.lib, "ncclGetUniqueId"));
    g_nccl.CommInitRank = reinterpret_cast<int (*)(void**, int, Id128, int)>(dlsym(g_nccl.lib, "ncclCommInitRank"));
    g_nccl.AllReduce = reinterpret_cast<int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t)>(
        dlsym(g_nccl.lib, "ncclAllReduce"));
    g_nccl.CommDestroy = reinterpret_cast<int (*)(void*)>(dlsym(g_nccl.lib, "ncclCommDestroy"));
    g_nccl.GetErrorString = reinterpret_cast<const char* (*)(int)>(dlsym(g_nccl.lib, "ncclGetErrorString"));
    g_nccl.GroupStart = reinterpret_cast<int (*)()>(dlsym(g_nccl.lib, "ncclGroupStart"));
    g_nccl.GroupEnd = reinterpret_cast<int (*)()>(dlsym(g_nccl.lib, "ncclGroupEnd"));
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce || !g_nccl.CommDestroy || !g_nccl.GroupStart || !g_nccl.GroupEnd) {
        g_nccl.lib = nullptr;
        return "NCCL symbols missing";
    }
    return nullptr;
}
static constexpr int kNcclUint8 = 1, kNcclFloat32 = 7, kNcclFloat64 = 8;   // ncclDataType_t
static constexpr int kNcclSum = 0, kNcclMax = 2;

// ------------------------------------------------------------------------------------------------ problem object
struct msfm_ba {
    msfm_ctx* ctx = nullptr;
    int32_t n_cams = 0, n_pts = 0, n_obs = 0, n_free = 0;
    int32_t refine_focal = 0;
    double focal[2][2] = {{0, 0}, {0, 0}};      // (fx, fy) of the current / candidate parameter set
    double* pt_Wf = nullptr;
    // device arrays
    double *cams[2] = {nullptr, nullptr}, *pts[2] = {nullptr, nullptr};   // current / candidate
    CamPre* pre[2] = {nullptr, nullptr};
    double* obs_uv = nullptr;
    int32_t *obs_cam = nullptr, *obs_pt = nullptr, *obs_orig = nullptr, *pt_start = nullptr, *pt_order = nullptr, *cam_free = nullptr;
    uint8_t* obs_lcam = nullptr;
    // tiling + block structure (ba_types.cuh, ba_tiles.hpp)
    Tile* tiles = nullptr;
    Item* items = nullptr;
    uint32_t* runs = nullptr;
    int32_t *tile_cams = nullptr, *tile_free = nullptr, *tile_slots = nullptr, *blk_row = nullptr, *blk_col = nullptr;
    int32_t n_tiles = 0, w_max = 0, n_blocks = 0, first_long = 0, n_long = 0;
    uint8_t* obs_lpt = nullptr;
    double* long_V = nullptr;      // per long track: V^-1 | g_p (| Wf), written by the pre-pass of every linearisation
    std::vector<int32_t> h_blk_row, h_blk_col;
    // the system of one linearisation, ONE allocation: tail (fp64: scalars | per-rank max |g_p| | rhs | gc | diag U | focal
    // border) | tile counter | sblk (fp32 6x6 blocks).  tail and sblk are the two parts of the all-reduce message.
    unsigned char* sysbuf = nullptr;
    size_t sys_bytes = 0;
    TailLayout tl{};
    float* sblk = nullptr;
    int32_t* tile_counter = nullptr;
    double* dense = nullptr;      // [n6][n6] dense copy of S for the dense Cholesky (allocated by the first solve that needs it)
    band::BandSolver* bands = nullptr;   // own cooperative band Cholesky after renumbering (ba_band.cu), when the band is narrow
    ba::TridiagSolver* tri = nullptr;    // the same band as a block-tridiagonal chain of library calls (ba_solver.cu; MSFM_BA_SOLVER=chain)
    bool tri_tried = false;
    double* rec = nullptr;        // [16] per-iteration record (lm_record_kernel)
    cudaEvent_t ev[2] = {nullptr, nullptr};
    double* xsol = nullptr;       // solver right-hand side / solution [n6]
    double* small = nullptr;      // [8] scratch scalars (backsub out[3], new cost)
    double* work = nullptr;       // cusolver workspace
    int work_len = 0;
    int* dev_info = nullptr;
    int cur = 0;
    std::vector<double> h_cams;   // host mirror of the current cameras
    std::vector<int32_t> h_cam_free;
    // Everything above that is sized by the problem lives in ONE device allocation (the arena), carved up by build_problem;
    // msfm_ba_update re-uses it for the next problem when it is large enough (no cudaMalloc / cudaFree on the steady path).
    unsigned char* arena = nullptr;
    size_t arena_cap = 0;
    // host copy of the sparsity pattern the structure was analysed for (msfm_ba_update compares the next problem with it)
    int32_t flags = 0;
    std::vector<int32_t> h_obs_cam, h_obs_pt;
    std::vector<uint8_t> h_cam_const;
    int64_t last_h2d_bytes = 0;   // host -> device bytes of the last create / update
    int32_t last_reused = 0;      // 1: the last update kept the structure
    double* tail() const { return reinterpret_cast<double*>(sysbuf); }
    Problem view(int which) const {
        Problem P{};
        P.n_cams = n_cams; P.n_pts = n_pts; P.n_obs = n_obs; P.n_free = n_free;
        P.fx = focal[which][0]; P.fy = focal[which][1];
        P.refine_focal = refine_focal; P.pt_Wf = pt_Wf;
        P.pre = pre[which]; P.pts = pts[which]; P.obs_uv = obs_uv; P.obs_cam = obs_cam; P.obs_pt = obs_pt; P.obs_orig = obs_orig;
        P.obs_lcam = obs_lcam; P.pt_start = pt_start; P.pt_order = pt_order; P.cam_free = cam_free;
        P.tiles = tiles; P.items = items; P.runs = runs; P.n_tiles = n_tiles; P.tile_cams = tile_cams; P.tile_free = tile_free; P.tile_slots = tile_slots;
        P.obs_lpt = obs_lpt; P.first_long = first_long; P.n_long = n_long; P.long_V = long_V;
        P.n_blocks = n_blocks; P.blk_row = blk_row; P.blk_col = blk_col;
        P.sblk = sblk; P.tail = tail(); P.tl = tl; P.gpm_slot = ctx->comm ? ctx->comm_rank : 0; P.tile_counter = tile_counter;
        return P;
    }
};

#define BA_CUDA(call) MSFM_CUDA(c, call)

static int ensure_solver(msfm_ctx* c) {
    if (c->cusolver) return MSFM_OK;
    cusolverDnHandle_t h = nullptr;
    if (cusolverDnCreate(&h) != CUSOLVER_STATUS_SUCCESS) return c->fail(MSFM_E_CUDA, "cusolverDnCreate failed");
    if (cusolverDnSetStream(h, c->stream) != CUSOLVER_STATUS_SUCCESS) return c->fail(MSFM_E_CUDA, "cusolverDnSetStream failed");
    c->cusolver = h;
    return MSFM_OK;
}

namespace msfm {
void ctx_destroy_solver(msfm_ctx* c) {
    if (c && c->cusolver) { cusolverDnDestroy(static_cast<cusolverDnHandle_t>(c->cusolver)); c->cusolver = nullptr; }
}
}  // namespace msfm

// sum / max of a device buffer across the ranks of the ctx's communicator, in place, on the ctx stream
static int comm_allreduce(msfm_ctx* c, void* buf, size_t count, int dtype, int op) {
    if (count == 0) return MSFM_OK;
    const int r = g_nccl.AllReduce(buf, buf, count, dtype, op, c->comm, c->stream);
    if (r != 0) return c->fail(MSFM_E_CUDA, "ncclAllReduce: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "error");
    return MSFM_OK;
}

extern "C" {

void msfm_ba_default_options(msfm_ba_options* o, int32_t n_cams) {
    if (!o) return;
    o->max_num_iterations = 100;            // CeresBundleOptimizer.cpp:276
    o->verbose = 0;
    o->function_tolerance = 1e-6;           // Ceres Solver::Options defaults
    o->gradient_tolerance = 1e-10;
    o->parameter_tolerance = 1e-8;
    o->initial_trust_region_radius = 1e4;
    if (n_cams < 10) {                      // :282-291
        o->function_tolerance /= 10;
        o->gradient_tolerance /= 10;
        o->parameter_tolerance /= 10;
        o->max_num_iterations *= 2;
    }
}

// solver objects and buffers that depend on the block structure (allocated lazily by the first solve)
static void release_solver_state(msfm_ba* b) {
    if (b->tri) tridiag_destroy(b->tri);
    if (b->bands) band_destroy(b->bands);
    b->tri = nullptr; b->bands = nullptr; b->tri_tried = false;
    if (b->dense) cudaFree(b->dense);
    b->dense = nullptr;
}

void msfm_ba_destroy(msfm_ba* b) {
    if (!b) return;
    msfm_ctx* c = b->ctx;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    release_solver_state(b);
    void* ptrs[] = {b->arena, b->work, b->rec};
    for (void* p : ptrs)
        if (p) cudaFree(p);
    for (cudaEvent_t e : b->ev)
        if (e) cudaEventDestroy(e);
    delete b;
}

static int check_problem(msfm_ctx* c, const msfm_ba_problem* pr, const char* who) {
    if (!pr) return c->fail(MSFM_E_INVALID, "%s: null argument", who);
    if (pr->n_cams <= 0 || pr->n_pts < 0 || pr->n_obs < 0 || (pr->flags & ~MSFM_BA_REFINE_FOCAL) != 0 || !pr->cams || !pr->cam_const ||
        (pr->n_pts > 0 && !pr->pts) || (pr->n_obs > 0 && (!pr->obs_uv || !pr->obs_cam || !pr->obs_pt)))
        return c->fail(MSFM_E_INVALID, "%s: bad problem description", who);
    return MSFM_OK;
}

// Parameters and measurements of `pr` -> device (the structure of b must be the one of pr): cameras and points into the
// current parameter set, the measurements in the caller's order into a staging buffer and from there into device order with
// one kernel (obs_orig), and, with `with_cams`, the camera index of every observation the same way.
static int upload_values(msfm_ba* b, const msfm_ba_problem* pr, bool with_cams) {
    msfm_ctx* c = b->ctx;
    const size_t no = size_t(b->n_obs);
    BA_CUDA(cudaMemcpyAsync(b->cams[b->cur], pr->cams, size_t(b->n_cams) * 6 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    b->last_h2d_bytes += int64_t(b->n_cams) * 6 * sizeof(double);
    if (b->n_pts) {
        BA_CUDA(cudaMemcpyAsync(b->pts[b->cur], pr->pts, size_t(b->n_pts) * 3 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        b->last_h2d_bytes += int64_t(b->n_pts) * 3 * sizeof(double);
    }
    if (no) {
        BA_CUDA(c->d_ba_J.reserve(no * (2 * sizeof(double) + sizeof(int32_t))));
        double* st_uv = c->d_ba_J.as<double>();
        int32_t* st_cam = reinterpret_cast<int32_t*>(st_uv + 2 * no);
        BA_CUDA(cudaMemcpyAsync(st_uv, pr->obs_uv, no * 2 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        b->last_h2d_bytes += int64_t(no) * 2 * sizeof(double);
        if (with_cams) {
            BA_CUDA(cudaMemcpyAsync(st_cam, pr->obs_cam, no * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
            b->last_h2d_bytes += int64_t(no) * sizeof(int32_t);
        }
        BA_CUDA(ba_launch_permute_obs(b->n_obs, b->obs_orig, st_uv, with_cams ? st_cam : nullptr, b->obs_uv, b->obs_cam, c->stream));
    }
    b->focal[0][0] = b->focal[1][0] = pr->fx; b->focal[0][1] = b->focal[1][1] = pr->fy;
    b->h_cams.assign(pr->cams, pr->cams + size_t(pr->n_cams) * 6);
    return MSFM_OK;
}

// Structure analysis of `pr` on the host threads, carve-up of the arena, upload.  Used by msfm_ba_create and by
// msfm_ba_update when the sparsity pattern has changed.
static int build_problem(msfm_ba* b, const msfm_ba_problem* pr) {
    msfm_ctx* c = b->ctx;
    b->last_h2d_bytes = 0;
    b->last_reused = 0;
    // validate indices, free-camera map
    std::vector<int32_t> cam_free(pr->n_cams, -1);
    {
        std::atomic<int> bad{-1};
        ba::parallel_ranges(pr->n_obs, [&](int i0, int i1) {
            int prev = i0 > 0 ? pr->obs_pt[i0 - 1] : 0;
            for (int i = i0; i < i1; ++i) {
                const int p = pr->obs_pt[i], cam = pr->obs_cam[i];
                if (p < prev || p >= pr->n_pts || cam < 0 || cam >= pr->n_cams) { bad.store(i); return; }
                prev = p;
            }
        });
        if (bad.load() >= 0)
            return c->fail(MSFM_E_INVALID, "msfm_ba_create: observation %d has bad indices (obs_pt must be non-decreasing)", bad.load());
    }
    int nf = 0;
    for (int i = 0; i < pr->n_cams; ++i)
        if (!pr->cam_const[i]) cam_free[i] = nf++;
    // ---- structure analysis (host): device order, tiles, block structure of the reduced camera system
    ba::TilingParams tp;
    {
        // development overrides of the tile shape (ba_tiles.hpp clamps them to the kernel's limits)
        const char* e1 = getenv("MSFM_BA_TILE_OBS");
        const char* e2 = getenv("MSFM_BA_TILE_PTS");
        if (e1 && atoi(e1) > 0) tp.max_obs = atoi(e1);
        if (e2 && atoi(e2) > 0) tp.max_pts = atoi(e2);
        const char* e3 = getenv("MSFM_BA_MAX_RUN");
        if (e3 && atoi(e3) > 0) tp.max_run = atoi(e3);
    }
    ba::Tiling T;
    // a landmark holds at most one measurement per image (tracks in Map.cpp are keyed by image): the pair products rely on it
    if (!ba::build_tiling(pr->n_cams, pr->n_pts, pr->n_obs, pr->obs_cam, pr->obs_pt, cam_free.data(), tp, T))
        return c->fail(MSFM_E_INVALID, "msfm_ba_create: a point is observed twice by the same camera");
    std::vector<uint8_t> present;
    ba::mark_blocks(T, cam_free.data(), nf, present);
    if (c->comm && c->comm_ranks > 1 && !present.empty()) {
        // every rank owns other points: the block structure of the summed system is the union over the ranks
        uint8_t* d_present = nullptr;
        BA_CUDA(cudaMalloc(&d_present, present.size()));
        cudaError_t e = cudaMemcpyAsync(d_present, present.data(), present.size(), cudaMemcpyHostToDevice, c->stream);
        int rc = MSFM_OK;
        if (e == cudaSuccess) rc = comm_allreduce(c, d_present, present.size(), kNcclUint8, kNcclMax);
        if (e == cudaSuccess && rc == MSFM_OK) e = cudaMemcpyAsync(present.data(), d_present, present.size(), cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess && rc == MSFM_OK) e = cudaStreamSynchronize(c->stream);
        cudaFree(d_present);
        if (rc) return rc;
        if (e != cudaSuccess) return c->cuda_fail(e, "msfm_ba_create: merging the block structure across ranks");
    }
    ba::assign_slots(T, cam_free.data(), nf, present);
    present.clear(); present.shrink_to_fit();

    release_solver_state(b);                     // the band / dense solvers were set up for the previous block structure
    b->n_cams = pr->n_cams; b->n_pts = pr->n_pts; b->n_obs = pr->n_obs; b->n_free = nf;
    b->flags = pr->flags;
    b->refine_focal = (pr->flags & MSFM_BA_REFINE_FOCAL) ? 1 : 0;
    b->cur = 0;
    b->h_cam_free = cam_free;
    b->n_tiles = static_cast<int32_t>(T.tiles.size());
    b->w_max = T.w_max;
    b->first_long = T.first_long; b->n_long = T.n_long;
    b->n_blocks = static_cast<int32_t>(T.blk_col.size());
    b->h_blk_row = T.blk_row; b->h_blk_col = T.blk_col;
    b->h_obs_cam.assign(pr->obs_cam, pr->obs_cam + pr->n_obs);
    b->h_obs_pt.assign(pr->obs_pt, pr->obs_pt + pr->n_obs);
    b->h_cam_const.assign(pr->cam_const, pr->cam_const + pr->n_cams);
    const size_t n6 = size_t(nf) * 6;
    {
        const int R = c->comm ? c->comm_ranks : 1;
        TailLayout& tl = b->tl;
        tl.scal = 0; tl.gpm = 8; tl.rhs = 8 + R; tl.gc = tl.rhs + int(n6); tl.udiag = tl.gc + int(n6);
        tl.B0 = tl.udiag + int(n6); tl.B1 = tl.B0 + (b->refine_focal ? int(n6) : 0); tl.ff = tl.B1 + (b->refine_focal ? int(n6) : 0);
        tl.total = tl.ff + (b->refine_focal ? 16 : 0);
        tl.total = (tl.total + 1) & ~1;          // sblk stays 16-byte aligned behind the 16-byte counter slot
    }
    b->sys_bytes = size_t(b->tl.total) * sizeof(double) + 16 + size_t(b->n_blocks) * 36 * sizeof(float);
    // ---- carve-up of the arena (256-byte aligned parts)
    struct Part { void** ptr; size_t bytes; };
    const size_t np = size_t(pr->n_pts), no = size_t(pr->n_obs), nc = size_t(pr->n_cams);
    std::vector<Part> parts = {
        {reinterpret_cast<void**>(&b->cams[0]), nc * 6 * sizeof(double)}, {reinterpret_cast<void**>(&b->cams[1]), nc * 6 * sizeof(double)},
        {reinterpret_cast<void**>(&b->pts[0]), np * 3 * sizeof(double)},  {reinterpret_cast<void**>(&b->pts[1]), np * 3 * sizeof(double)},
        {reinterpret_cast<void**>(&b->pre[0]), nc * sizeof(CamPre)},      {reinterpret_cast<void**>(&b->pre[1]), nc * sizeof(CamPre)},
        {reinterpret_cast<void**>(&b->obs_uv), no * 2 * sizeof(double)},  {reinterpret_cast<void**>(&b->obs_cam), no * sizeof(int32_t)},
        {reinterpret_cast<void**>(&b->obs_pt), no * sizeof(int32_t)},     {reinterpret_cast<void**>(&b->obs_orig), no * sizeof(int32_t)},
        {reinterpret_cast<void**>(&b->obs_lcam), no},                     {reinterpret_cast<void**>(&b->obs_lpt), no},
        {reinterpret_cast<void**>(&b->long_V), size_t(std::max(1, T.n_long)) * 15 * sizeof(double)},
        {reinterpret_cast<void**>(&b->pt_start), (np + 1) * sizeof(int32_t)}, {reinterpret_cast<void**>(&b->pt_order), np * sizeof(int32_t)},
        {reinterpret_cast<void**>(&b->cam_free), nc * sizeof(int32_t)},
        {reinterpret_cast<void**>(&b->tiles), T.tiles.size() * sizeof(Tile)}, {reinterpret_cast<void**>(&b->items), T.items.size() * sizeof(Item)},
        {reinterpret_cast<void**>(&b->runs), T.runs.size() * sizeof(uint32_t)},
        {reinterpret_cast<void**>(&b->tile_cams), T.tile_cams.size() * sizeof(int32_t)},
        {reinterpret_cast<void**>(&b->tile_free), T.tile_free.size() * sizeof(int32_t)},
        {reinterpret_cast<void**>(&b->tile_slots), T.tile_slots.size() * sizeof(int32_t)},
        {reinterpret_cast<void**>(&b->blk_row), T.blk_row.size() * sizeof(int32_t)}, {reinterpret_cast<void**>(&b->blk_col), T.blk_col.size() * sizeof(int32_t)},
        {reinterpret_cast<void**>(&b->sysbuf), b->sys_bytes},
        {reinterpret_cast<void**>(&b->xsol), (3 * std::max<size_t>(1, n6) + 2) * sizeof(double)},     // up to 3 right-hand sides + (d fx, d fy)
        {reinterpret_cast<void**>(&b->pt_Wf), b->refine_focal ? std::max<size_t>(1, np) * 6 * sizeof(double) : 0},
        {reinterpret_cast<void**>(&b->small), 8 * sizeof(double)},        {reinterpret_cast<void**>(&b->dev_info), 4 * sizeof(int)},
    };
    size_t total = 0;
    for (const Part& q : parts) total += (std::max<size_t>(16, q.bytes) + 255) / 256 * 256;
    if (total > b->arena_cap) {
        BA_CUDA(cudaStreamSynchronize(c->stream));
        if (b->arena) cudaFree(b->arena);
        b->arena = nullptr;
        const size_t want = b->arena_cap ? total + total / 4 : total;      // a problem that grew will grow again
        b->arena_cap = 0;
        cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&b->arena), want);
        if (e != cudaSuccess) return c->cuda_fail(e, "cudaMalloc(problem arena)");
        b->arena_cap = want;
    }
    {
        size_t off = 0;
        for (const Part& q : parts) {
            *q.ptr = q.bytes || q.ptr != reinterpret_cast<void**>(&b->pt_Wf) ? b->arena + off : nullptr;
            off += (std::max<size_t>(16, q.bytes) + 255) / 256 * 256;
        }
    }
    b->tile_counter = reinterpret_cast<int32_t*>(b->sysbuf + size_t(b->tl.total) * sizeof(double));
    b->sblk = reinterpret_cast<float*>(b->sysbuf + size_t(b->tl.total) * sizeof(double) + 16);
    // ---- upload: structure tables, then the values (measurements and camera indices are permuted on the device)
    cudaError_t e = cudaSuccess;
    auto H2D = [&](void* dst, const void* src, size_t bytes) {
        if (e != cudaSuccess || bytes == 0) return;
        e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, c->stream);
        b->last_h2d_bytes += int64_t(bytes);
    };
    H2D(b->obs_orig, T.obs_perm.data(), T.obs_perm.size() * sizeof(int32_t));
    H2D(b->obs_lcam, T.obs_lcam.data(), T.obs_lcam.size());
    H2D(b->obs_lpt, T.obs_lpt.data(), T.obs_lpt.size());
    H2D(b->pt_start, T.pt_start.data(), T.pt_start.size() * sizeof(int32_t));
    H2D(b->pt_order, T.pt_order.data(), T.pt_order.size() * sizeof(int32_t));
    H2D(b->cam_free, cam_free.data(), cam_free.size() * sizeof(int32_t));
    H2D(b->tiles, T.tiles.data(), T.tiles.size() * sizeof(Tile));
    H2D(b->items, T.items.data(), T.items.size() * sizeof(Item));
    H2D(b->runs, T.runs.data(), T.runs.size() * sizeof(uint32_t));
    H2D(b->tile_cams, T.tile_cams.data(), T.tile_cams.size() * sizeof(int32_t));
    H2D(b->tile_free, T.tile_free.data(), T.tile_free.size() * sizeof(int32_t));
    H2D(b->tile_slots, T.tile_slots.data(), T.tile_slots.size() * sizeof(int32_t));
    H2D(b->blk_row, T.blk_row.data(), T.blk_row.size() * sizeof(int32_t));
    H2D(b->blk_col, T.blk_col.data(), T.blk_col.size() * sizeof(int32_t));
    if (e != cudaSuccess) return c->cuda_fail(e, "msfm_ba_create H2D");
    if (pr->n_pts > 0) {
        BA_CUDA(ba_launch_fill_obs_pt(pr->n_pts, b->pt_start, b->pt_order, b->obs_pt, c->stream));
    }
    int rc = upload_values(b, pr, true);
    if (rc) return rc;
    BA_CUDA(cudaStreamSynchronize(c->stream));       // the host tables of T go out of scope
    return MSFM_OK;
}

int msfm_ba_create(msfm_ctx* c, const msfm_ba_problem* pr, msfm_ba** out) {
    if (!c) return MSFM_E_INVALID;
    if (!out) return c->fail(MSFM_E_INVALID, "msfm_ba_create: null argument");
    *out = nullptr;
    int rc = check_problem(c, pr, "msfm_ba_create");
    if (rc) return rc;
    BA_CUDA(cudaSetDevice(c->device));
    msfm_ba* b = new (std::nothrow) msfm_ba();
    if (!b) return c->fail(MSFM_E_CUDA, "out of host memory");
    b->ctx = c;
    if ((rc = build_problem(b, pr))) { msfm_ba_destroy(b); return rc; }
    *out = b;
    return MSFM_OK;
}

int msfm_ba_update(msfm_ba* b, const msfm_ba_problem* pr, int32_t* reused) {
    if (!b) return MSFM_E_INVALID;
    msfm_ctx* c = b->ctx;
    int rc = check_problem(c, pr, "msfm_ba_update");
    if (rc) return rc;
    BA_CUDA(cudaSetDevice(c->device));
    // same sparsity pattern?  sizes, flags, constant cameras and the two index arrays, compared exactly
    int differs = pr->n_cams != b->n_cams || pr->n_pts != b->n_pts || pr->n_obs != b->n_obs || pr->flags != b->flags;
    if (!differs) differs = std::memcmp(pr->cam_const, b->h_cam_const.data(), size_t(pr->n_cams)) != 0;
    if (!differs && pr->n_obs > 0) {
        std::atomic<int> d{0};
        ba::parallel_ranges(pr->n_obs, [&](int i0, int i1) {
            if (std::memcmp(pr->obs_cam + i0, b->h_obs_cam.data() + i0, size_t(i1 - i0) * sizeof(int32_t)) != 0 ||
                std::memcmp(pr->obs_pt + i0, b->h_obs_pt.data() + i0, size_t(i1 - i0) * sizeof(int32_t)) != 0) d.store(1);
        });
        differs = d.load();
    }
    if (c->comm && c->comm_ranks > 1) {
        // every rank has to take the same path (a rebuild merges the block structure with a collective)
        double v = differs ? 1.0 : 0.0;
        BA_CUDA(cudaMemcpyAsync(b->small, &v, sizeof(double), cudaMemcpyHostToDevice, c->stream));
        if ((rc = msfm_comm_allreduce_f64(c, b->small, 1, 1))) return rc;
        BA_CUDA(cudaMemcpyAsync(&v, b->small, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        BA_CUDA(cudaStreamSynchronize(c->stream));
        differs = v != 0.0;
    }
    if (differs) {
        BA_CUDA(cudaStreamSynchronize(c->stream));
        rc = build_problem(b, pr);
    } else {
        b->last_h2d_bytes = 0;
        b->last_reused = 1;
        rc = upload_values(b, pr, false);
        if (rc == MSFM_OK) BA_CUDA(cudaStreamSynchronize(c->stream));       // the caller's arrays may go away
    }
    if (reused) *reused = b->last_reused;
    return rc;
}

int msfm_ba_sizes(msfm_ba* b, int64_t sizes[3]) {
    if (!b || !sizes) return MSFM_E_INVALID;
    sizes[0] = b->n_cams; sizes[1] = b->n_pts; sizes[2] = b->n_obs;
    return MSFM_OK;
}

int msfm_ba_last_upload(msfm_ba* b, int64_t info[2]) {
    if (!b || !info) return MSFM_E_INVALID;
    info[0] = b->last_h2d_bytes;
    info[1] = b->last_reused;
    return MSFM_OK;
}

int msfm_ba_structure(msfm_ba* b, int32_t info[8]) {
    if (!b || !info) return MSFM_E_INVALID;
    info[0] = b->n_free; info[1] = b->n_blocks; info[2] = b->n_tiles; info[3] = b->w_max;
    info[4] = static_cast<int32_t>(std::min<size_t>(b->sys_bytes, 0x7fffffff));
    info[5] = static_cast<int32_t>(ba_fused_smem_bytes(b->refine_focal != 0));
    info[6] = b->tl.total; info[7] = b->n_long;
    return MSFM_OK;
}

int msfm_ba_get_params(msfm_ba* b, double* cams, double* pts) {
    if (!b) return MSFM_E_INVALID;
    msfm_ctx* c = b->ctx;
    BA_CUDA(cudaSetDevice(c->device));
    BA_CUDA(cudaStreamSynchronize(c->stream));
    if (cams) BA_CUDA(cudaMemcpy(cams, b->cams[b->cur], size_t(b->n_cams) * 6 * sizeof(double), cudaMemcpyDeviceToHost));
    if (pts && b->n_pts) BA_CUDA(cudaMemcpy(pts, b->pts[b->cur], size_t(b->n_pts) * 3 * sizeof(double), cudaMemcpyDeviceToHost));
    return MSFM_OK;
}
int msfm_ba_set_params(msfm_ba* b, const double* cams, const double* pts) {
    if (!b) return MSFM_E_INVALID;
    msfm_ctx* c = b->ctx;
    BA_CUDA(cudaSetDevice(c->device));
    BA_CUDA(cudaStreamSynchronize(c->stream));
    if (cams) {
        BA_CUDA(cudaMemcpy(b->cams[b->cur], cams, size_t(b->n_cams) * 6 * sizeof(double), cudaMemcpyHostToDevice));
        b->h_cams.assign(cams, cams + size_t(b->n_cams) * 6);
    }
    if (pts && b->n_pts) BA_CUDA(cudaMemcpy(b->pts[b->cur], pts, size_t(b->n_pts) * 3 * sizeof(double), cudaMemcpyHostToDevice));
    return MSFM_OK;
}

static int prep(msfm_ba* b, int which) {
    msfm_ctx* c = b->ctx;
    c->prof_begin(MSFM_PROF_BA_OTHER);
    BA_CUDA(ba_launch_cam_prep(b->cams[which], b->n_cams, b->pre[which], c->stream));
    c->prof_end();
    return MSFM_OK;
}

int msfm_ba_evaluate(msfm_ba* b, double* r, float* J, double* cost) {
    if (!b) return MSFM_E_INVALID;
    msfm_ctx* c = b->ctx;
    BA_CUDA(cudaSetDevice(c->device));
    int rc = prep(b, b->cur);
    if (rc) return rc;
    // parity dumps go through two grow-only scratch buffers of the context (no allocation on the steady path)
    double* d_r = nullptr;
    float* d_J = nullptr;
    if (r && b->n_obs) {
        BA_CUDA(c->d_ba_r.reserve(size_t(b->n_obs) * 2 * sizeof(double)));
        d_r = c->d_ba_r.as<double>();
    }
    if (J && b->n_obs) {
        BA_CUDA(c->d_ba_J.reserve(size_t(b->n_obs) * 18 * sizeof(float)));
        d_J = c->d_ba_J.as<float>();
    }
    BA_CUDA(cudaMemsetAsync(b->small, 0, 8 * sizeof(double), c->stream));
    c->prof_begin(MSFM_PROF_BA_EVAL);
    BA_CUDA(ba_launch_evaluate(b->view(b->cur), d_r, d_J, b->small, c->num_sms, c->stream));
    c->prof_end();
    double h = 0;
    BA_CUDA(cudaMemcpyAsync(&h, b->small, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    if (d_r) BA_CUDA(cudaMemcpyAsync(r, d_r, size_t(b->n_obs) * 2 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    if (d_J) BA_CUDA(cudaMemcpyAsync(J, d_J, size_t(b->n_obs) * 18 * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    BA_CUDA(cudaStreamSynchronize(c->stream));
    if (cost) *cost = h;
    return MSFM_OK;
}

int msfm_ba_track_errors(msfm_ba* b, double* err) {
    if (!b) return MSFM_E_INVALID;
    msfm_ctx* c = b->ctx;
    if (!err && b->n_pts > 0) return c->fail(MSFM_E_INVALID, "msfm_ba_track_errors: null output");
    if (b->n_pts == 0) return MSFM_OK;
    BA_CUDA(cudaSetDevice(c->device));
    int rc = prep(b, b->cur);
    if (rc) return rc;
    BA_CUDA(c->d_ba_r.reserve(size_t(b->n_pts) * sizeof(double)));
    c->prof_begin(MSFM_PROF_BA_EVAL);
    BA_CUDA(ba_launch_track_errors(b->view(b->cur), c->d_ba_r.as<double>(), c->num_sms, c->stream));
    c->prof_end();
    BA_CUDA(cudaMemcpyAsync(err, c->d_ba_r.p, size_t(b->n_pts) * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    BA_CUDA(cudaStreamSynchronize(c->stream));
    return MSFM_OK;
}

int msfm_ba_filter_stats(msfm_ba* b, double max_reproj_error, uint8_t* obs_keep, double* pt_mean_error, int32_t* pt_kept, double* pt_max_parallax_deg) {
    if (!b) return MSFM_E_INVALID;
    msfm_ctx* c = b->ctx;
    if (b->n_pts == 0) return MSFM_OK;
    BA_CUDA(cudaSetDevice(c->device));
    int rc = prep(b, b->cur);
    if (rc) return rc;
    // scratch: err [n_pts] f64 | angle [n_pts] f64 | kept [n_pts] i32 | keep [n_obs] u8
    const size_t np = size_t(b->n_pts), no = size_t(b->n_obs);
    BA_CUDA(c->d_ba_r.reserve(np * 20 + no + 64));
    double* d_err = c->d_ba_r.as<double>();
    double* d_ang = d_err + np;
    int32_t* d_kept = reinterpret_cast<int32_t*>(d_ang + np);
    uint8_t* d_keep = reinterpret_cast<uint8_t*>(d_kept + np);
    c->prof_begin(MSFM_PROF_BA_EVAL);
    BA_CUDA(ba_launch_filter_stats(b->view(b->cur), max_reproj_error, d_keep, d_err, d_kept, d_ang, c->num_sms, c->stream));
    c->prof_end();
    if (obs_keep && no) BA_CUDA(cudaMemcpyAsync(obs_keep, d_keep, no, cudaMemcpyDeviceToHost, c->stream));
    if (pt_mean_error) BA_CUDA(cudaMemcpyAsync(pt_mean_error, d_err, np * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    if (pt_kept) BA_CUDA(cudaMemcpyAsync(pt_kept, d_kept, np * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
    if (pt_max_parallax_deg) BA_CUDA(cudaMemcpyAsync(pt_max_parallax_deg, d_ang, np * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    BA_CUDA(cudaStreamSynchronize(c->stream));
    return MSFM_OK;
}

// Zero the system, linearize + Schur at parameter set `which` (one kernel), all-reduce.  The message of the ONE
// collective (a single NCCL group launch) is the system buffer itself: the fp32 blocks of the upper block triangle that
// exist (blk_row / blk_col) and the fp64 tail, in which every rank's max |g_p| has its own slot so that the sum also
// delivers the maximum.  S stays undamped on the camera diagonal (diag U has to be summed first): the expansion for
// the solver adds the damping.
static int linearize(msfm_ba* b, int which, double inv_radius) {
    msfm_ctx* c = b->ctx;
    BA_CUDA(cudaMemsetAsync(b->sysbuf, 0, b->sys_bytes, c->stream));
    c->prof_begin(MSFM_PROF_BA_SCHUR);
    BA_CUDA(ba_launch_linearize(b->view(which), inv_radius, c->num_sms, c->stream));
    c->prof_end();
    if (c->comm && c->comm_ranks > 1) {
        c->prof_begin(MSFM_PROF_BA_COMM);
        if (g_nccl.GroupStart() != 0) return c->fail(MSFM_E_CUDA, "ncclGroupStart failed");
        int rc = comm_allreduce(c, b->sblk, size_t(b->n_blocks) * 36, kNcclFloat32, kNcclSum);
        if (rc == MSFM_OK) rc = comm_allreduce(c, b->tail(), size_t(b->tl.total), kNcclFloat64, kNcclSum);
        const int ge = g_nccl.GroupEnd();
        c->prof_end();
        if (rc) return rc;
        if (ge != 0) return c->fail(MSFM_E_CUDA, "ncclGroupEnd: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(ge) : "error");
    }
    return MSFM_OK;
}

int msfm_ba_linearize(msfm_ba* b, double inv_radius, double* S, double* rhs, double* gc, double* cost, int32_t* nf) {
    if (!b) return MSFM_E_INVALID;
    msfm_ctx* c = b->ctx;
    BA_CUDA(cudaSetDevice(c->device));
    int rc = prep(b, b->cur);
    if (rc) return rc;
    if ((rc = linearize(b, b->cur, inv_radius))) return rc;
    const size_t N = size_t(b->n_free) * 6;
    std::vector<double> tail(size_t(b->tl.total));
    std::vector<float> blk;
    BA_CUDA(cudaMemcpyAsync(tail.data(), b->tail(), tail.size() * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    if (S) {
        blk.resize(size_t(b->n_blocks) * 36);
        if (!blk.empty()) BA_CUDA(cudaMemcpyAsync(blk.data(), b->sblk, blk.size() * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    }
    BA_CUDA(cudaStreamSynchronize(c->stream));
    if (S) {
        std::memset(S, 0, N * N * sizeof(double));
        for (int k = 0; k < b->n_blocks; ++k) {
            const size_t fa = size_t(b->h_blk_row[k]), fb = size_t(b->h_blk_col[k]);
            for (int i = 0; i < 6; ++i)
                for (int j = 0; j < 6; ++j) {
                    // the device holds the upper block triangle; a diagonal block is taken from its own upper triangle
                    if (fa == fb && j < i) continue;
                    const double v = blk[size_t(k) * 36 + 6 * i + j];
                    S[(fa * 6 + i) * N + fb * 6 + j] = v;
                    S[(fb * 6 + j) * N + fa * 6 + i] = v;
                }
        }
        for (size_t i = 0; i < N; ++i) S[i * N + i] += std::max(tail[size_t(b->tl.udiag) + i], 1e-6) * inv_radius;
    }
    if (rhs) std::memcpy(rhs, tail.data() + b->tl.rhs, N * sizeof(double));
    if (gc) std::memcpy(gc, tail.data() + b->tl.gc, N * sizeof(double));
    if (cost) *cost = tail[size_t(b->tl.scal)];
    if (nf) *nf = b->n_free;
    return MSFM_OK;
}

int msfm_ba_get_focal(msfm_ba* b, double focal[2]) {
    if (!b || !focal) return MSFM_E_INVALID;
    focal[0] = b->focal[b->cur][0];
    focal[1] = b->focal[b->cur][1];
    return MSFM_OK;
}

int msfm_ba_linearize_focal(msfm_ba* b, double inv_radius, double* B, double* F, double* rhs_f, double* g_f) {
    if (!b) return MSFM_E_INVALID;
    msfm_ctx* c = b->ctx;
    if (!b->refine_focal) return c->fail(MSFM_E_INVALID, "msfm_ba_linearize_focal: the problem has no shared focal block");
    BA_CUDA(cudaSetDevice(c->device));
    int rc = prep(b, b->cur);
    if (rc) return rc;
    if ((rc = linearize(b, b->cur, inv_radius))) return rc;
    const size_t N = size_t(b->n_free) * 6;
    std::vector<double> tail(size_t(b->tl.total));
    BA_CUDA(cudaMemcpyAsync(tail.data(), b->tail(), tail.size() * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    BA_CUDA(cudaStreamSynchronize(c->stream));
    const double* ff = tail.data() + b->tl.ff;
    if (B)
        for (size_t i = 0; i < N; ++i) { B[2 * i] = tail[size_t(b->tl.B0) + i]; B[2 * i + 1] = tail[size_t(b->tl.B1) + i]; }
    if (F) {
        F[0] = ff[0] + std::max(ff[7], 1e-6) * inv_radius;
        F[1] = ff[1];
        F[2] = ff[2] + std::max(ff[8], 1e-6) * inv_radius;
    }
    if (rhs_f) { rhs_f[0] = ff[3]; rhs_f[1] = ff[4]; }
    if (g_f) { g_f[0] = ff[5]; g_f[1] = ff[6]; }
    return MSFM_OK;
}

// allocations and the choice of the linear solver, once per problem
static int solver_setup(msfm_ba* b) {
    msfm_ctx* c = b->ctx;
    int rc = ensure_solver(c);
    if (rc) return rc;
    cusolverDnHandle_t solver = static_cast<cusolverDnHandle_t>(c->cusolver);
    const int n6 = b->n_free * 6;
    const size_t N = size_t(n6);
    if (!b->rec) BA_CUDA(cudaMalloc(reinterpret_cast<void**>(&b->rec), 16 * sizeof(double)));
    for (cudaEvent_t& e : b->ev)
        if (!e) BA_CUDA(cudaEventCreate(&e));
    if (n6 > 0 && !b->tri_tried) {
        b->tri_tried = true;
        // MSFM_BA_SOLVER = band (default: own band Cholesky) | chain (library block-tridiagonal chain) | dense (cuSOLVER potrf of
        // the whole system); the band solvers fall back to dense when the band is wide.  MSFM_BA_DENSE_SOLVER=1 = dense.
        const char* sel = getenv("MSFM_BA_SOLVER");
        const bool dense = getenv("MSFM_BA_DENSE_SOLVER") || (sel && !strcmp(sel, "dense"));
        cudaError_t e = cudaSuccess;
        if (!dense && sel && !strcmp(sel, "chain")) b->tri = tridiag_create(b->n_free, b->h_blk_row, b->h_blk_col, solver, c->stream, &e);
        else if (!dense) b->bands = band_create(b->n_free, b->h_blk_row, b->h_blk_col, c->num_sms, &e);
        if (e != cudaSuccess) return c->cuda_fail(e, "msfm_ba_solve: band solver setup");
    }
    if (n6 > 0 && !b->tri && !b->bands) {
        if (!b->dense) BA_CUDA(cudaMalloc(reinterpret_cast<void**>(&b->dense), N * N * sizeof(double)));
        int lwork = 0;
        if (cusolverDnDpotrf_bufferSize(solver, CUBLAS_FILL_MODE_LOWER, n6, b->dense, n6, &lwork) != CUSOLVER_STATUS_SUCCESS)
            return c->fail(MSFM_E_CUDA, "cusolverDnDpotrf_bufferSize failed");
        if (lwork > b->work_len) {
            if (b->work) cudaFree(b->work);
            b->work = nullptr;
            BA_CUDA(cudaMalloc(&b->work, size_t(lwork) * sizeof(double)));
            b->work_len = lwork;
        }
    }
    return MSFM_OK;
}
// queue expansion + factorisation + solve of S x = xsol (nrhs columns) on the ctx stream; status words at *info_ptr
static int solver_queue(msfm_ba* b, double inv_radius, int nrhs, const int** info_ptr, int* n_info) {
    msfm_ctx* c = b->ctx;
    cusolverDnHandle_t solver = static_cast<cusolverDnHandle_t>(c->cusolver);
    const int n6 = b->n_free * 6;
    const size_t N = size_t(n6);
    c->prof_begin(MSFM_PROF_BA_SOLVE);
    if (b->bands) {
        BA_CUDA(band_factor_solve(b->bands, b->view(b->cur), inv_radius, b->xsol, nrhs, c->stream));
        *info_ptr = band_dev_info(b->bands);
        *n_info = 1;
    } else if (b->tri) {
        BA_CUDA(tridiag_factor_solve(b->tri, b->view(b->cur), inv_radius, solver, b->xsol, nrhs, c->stream));
        *info_ptr = tridiag_dev_info(b->tri);
        *n_info = tridiag_n_super(b->tri);
    } else {
        // dense Cholesky: the row-major upper block triangle is the column-major lower triangle cuSOLVER reads
        BA_CUDA(cudaMemsetAsync(b->dense, 0, N * N * sizeof(double), c->stream));
        BA_CUDA(ba_launch_expand_dense(b->view(b->cur), inv_radius, b->dense, c->stream));
        if (cusolverDnDpotrf(solver, CUBLAS_FILL_MODE_LOWER, n6, b->dense, n6, b->work, b->work_len, b->dev_info) != CUSOLVER_STATUS_SUCCESS)
            return c->fail(MSFM_E_CUDA, "cusolverDnDpotrf failed to launch");
        // a failed factorisation leaves garbage behind; the record carries the status and the step is then rejected
        if (cusolverDnDpotrs(solver, CUBLAS_FILL_MODE_LOWER, n6, nrhs, b->dense, n6, b->xsol, n6, b->dev_info + 1) != CUSOLVER_STATUS_SUCCESS)
            return c->fail(MSFM_E_CUDA, "cusolverDnDpotrs failed to launch");
        *info_ptr = b->dev_info;
        *n_info = 1;
    }
    c->prof_end();
    return MSFM_OK;
}

int msfm_ba_solver_info(msfm_ba* b, int32_t info[6]) {
    if (!b || !info) return MSFM_E_INVALID;
    info[0] = info[1] = info[2] = info[3] = 0;
    info[4] = b->n_free; info[5] = 0;
    if (b->bands) { band_info(b->bands, info); info[5] = 2; }
    else if (b->tri) { tridiag_info(b->tri, info); info[5] = 1; }
    return MSFM_OK;
}

int msfm_ba_solve_system(msfm_ba* b, double inv_radius, double* dc, int32_t* status) {
    if (!b) return MSFM_E_INVALID;
    msfm_ctx* c = b->ctx;
    if (!dc && b->n_free > 0) return c->fail(MSFM_E_INVALID, "msfm_ba_solve_system: null output");
    if (b->refine_focal) return c->fail(MSFM_E_INVALID, "msfm_ba_solve_system: not for problems with a shared focal block");
    BA_CUDA(cudaSetDevice(c->device));
    int rc = solver_setup(b);
    if (rc) return rc;
    if ((rc = prep(b, b->cur))) return rc;
    if ((rc = linearize(b, b->cur, inv_radius))) return rc;
    const int n6 = b->n_free * 6;
    if (status) *status = 0;
    if (n6 == 0) return MSFM_OK;
    BA_CUDA(ba_launch_copy(b->tail() + b->tl.rhs, b->xsol, n6, c->stream));
    const int* info_ptr = nullptr;
    int n_info = 0;
    if ((rc = solver_queue(b, inv_radius, 1, &info_ptr, &n_info))) return rc;
    std::vector<int> h_info(size_t(std::max(1, n_info)), 0);
    BA_CUDA(cudaMemcpyAsync(h_info.data(), info_ptr, size_t(n_info) * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    BA_CUDA(cudaMemcpyAsync(dc, b->xsol, size_t(n6) * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    BA_CUDA(cudaStreamSynchronize(c->stream));
    if (status)
        for (int v : h_info) if (v != 0) *status = v;
    return MSFM_OK;
}

// The Levenberg-Marquardt loop of Ceres' TrustRegionMinimizer + LevenbergMarquardtStrategy for this problem class.  Every
// iteration is queued as a whole — linearise, solve the reduced camera system, back-substitute, evaluate the candidate, one
// small kernel that condenses everything the step decision needs into a 72-byte record — and the host synchronises ONCE per
// iteration to read that record and accept or reject the step (with a shared focal block the 2 x 2 Schur complement of
// the border adds a second round trip).
int msfm_ba_solve(msfm_ba* b, const msfm_ba_options* uopt, msfm_ba_summary* sum) {
    if (!b) return MSFM_E_INVALID;
    msfm_ctx* c = b->ctx;
    if (!uopt || !sum) return c->fail(MSFM_E_INVALID, "msfm_ba_solve: null argument");
    BA_CUDA(cudaSetDevice(c->device));
    int rc;
    using clk = std::chrono::steady_clock;
    const auto t_begin = clk::now();
    double t_lin = 0;
    const int n6 = b->n_free * 6;
    const size_t N = size_t(n6);
    if ((rc = solver_setup(b))) return rc;
    std::memset(sum, 0, sizeof *sum);
    double radius = uopt->initial_trust_region_radius;
    double decrease = 2.0;
    double cost = 0.0;
    bool have_cost = false, converged = false, failed = false;
    long long nres_local = 2LL * b->n_obs;
    const bool focal = b->refine_focal != 0;
    const int nrhs = focal ? 3 : 1;
    const TailLayout tl = b->tl;
    const int n_ranks = c->comm ? c->comm_ranks : 1;
    const bool multi = c->comm && c->comm_ranks > 1;
    std::vector<double> h_tail(focal ? size_t(tl.total) : 0), h_dc(focal ? N + 2 : 0), h_x(focal ? 3 * N : 0);
    double h_rec[16];
    int it = 0, good = 0;
    if ((rc = prep(b, b->cur))) return rc;
    while (it < uopt->max_num_iterations) {
        ++it;
        const double inv_radius = 1.0 / radius;
        BA_CUDA(cudaEventRecord(b->ev[0], c->stream));
        if ((rc = linearize(b, b->cur, inv_radius))) return rc;
        BA_CUDA(cudaEventRecord(b->ev[1], c->stream));
        // ---- solve the reduced camera system
        const int* info_ptr = b->dev_info;
        int n_info = 0;
        if (n6 > 0) {
            BA_CUDA(ba_launch_copy(b->tail() + tl.rhs, b->xsol, n6, c->stream));
            if (focal)         // two more right-hand sides: the border columns (S^-1 B for the 2 x 2 Schur complement below)
                BA_CUDA(cudaMemcpyAsync(b->xsol + N, b->tail() + tl.B0, 2 * N * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
            if ((rc = solver_queue(b, inv_radius, nrhs, &info_ptr, &n_info))) return rc;
        }
        double df[2] = {0.0, 0.0};
        bool solved = true;
        if (focal) {
            // bordered system [S B; B^T F] [dc; df] = [rhs; rhs_f]:  (F - B^T S^-1 B) df = rhs_f - B^T S^-1 rhs,  dc = S^-1 rhs - S^-1 B df
            BA_CUDA(cudaMemcpyAsync(h_tail.data(), b->tail(), h_tail.size() * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
            if (n6 > 0) BA_CUDA(cudaMemcpyAsync(h_x.data(), b->xsol, 3 * N * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
            BA_CUDA(cudaStreamSynchronize(c->stream));
            const double* hB = h_tail.data() + tl.B0;              // B0 | B1
            const double* hff = h_tail.data() + tl.ff;             // F00 F01 F11 rhsf0 rhsf1 gf0 gf1 uf0 uf1
            double f00 = hff[0] + std::max(hff[7], 1e-6) * inv_radius, f01 = hff[1], f11 = hff[2] + std::max(hff[8], 1e-6) * inv_radius;
            double r0 = hff[3], r1 = hff[4];
            const double *y = h_x.data(), *x0 = h_x.data() + N, *x1 = h_x.data() + 2 * N;
            for (size_t i = 0; i < N; ++i) {
                f00 -= hB[i] * x0[i]; f01 -= hB[i] * x1[i]; f11 -= hB[N + i] * x1[i];
                r0 -= hB[i] * y[i];   r1 -= hB[N + i] * y[i];
            }
            const double det = f00 * f11 - f01 * f01;
            if (!(det > 0.0) || !(f00 > 0.0) || !std::isfinite(det)) {
                solved = false;
            } else {
                df[0] = (f11 * r0 - f01 * r1) / det;
                df[1] = (f00 * r1 - f01 * r0) / det;
                for (size_t i = 0; i < N; ++i) h_dc[i] = y[i] - x0[i] * df[0] - x1[i] * df[1];
                h_dc[N] = df[0]; h_dc[N + 1] = df[1];
                BA_CUDA(cudaMemcpyAsync(b->xsol, h_dc.data(), (N + 2) * sizeof(double), cudaMemcpyHostToDevice, c->stream));
            }
        }
        // ---- back-substitute the points, build and evaluate the candidate
        const int nxt = b->cur ^ 1;
        BA_CUDA(cudaMemsetAsync(b->small, 0, 4 * sizeof(double), c->stream));
        if (solved) {
            c->prof_begin(MSFM_PROF_BA_OTHER);
            BA_CUDA(ba_launch_backsub(b->view(b->cur), inv_radius, b->xsol, b->pts[nxt], b->small, c->num_sms, c->stream));
            BA_CUDA(ba_launch_update_cams(b->cams[b->cur], b->cam_free, b->n_cams, b->xsol, b->cams[nxt], c->stream));
            c->prof_end();
            b->focal[nxt][0] = b->focal[b->cur][0] + df[0];
            b->focal[nxt][1] = b->focal[b->cur][1] + df[1];
            if ((rc = prep(b, nxt))) return rc;
            c->prof_begin(MSFM_PROF_BA_EVAL);
            BA_CUDA(ba_launch_evaluate(b->view(nxt), nullptr, nullptr, b->small + 3, c->num_sms, c->stream));
            c->prof_end();
            if (multi && (rc = msfm_comm_allreduce_f64(c, b->small, 4, 0))) return rc;
        }
        BA_CUDA(ba_launch_lm_record(b->view(b->cur), b->cams[b->cur], b->xsol, b->small, info_ptr, n_info, n_ranks, b->rec, c->stream));
        BA_CUDA(cudaMemcpyAsync(h_rec, b->rec, 9 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        BA_CUDA(cudaStreamSynchronize(c->stream));                  // the iteration's one synchronisation
        {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, b->ev[0], b->ev[1]) == cudaSuccess) t_lin += ms * 1e-3;
        }
        const double lin_cost = h_rec[0], gmax = h_rec[1];
        if (!have_cost) { cost = lin_cost; sum->initial_cost = cost; have_cost = true; }
        if (gmax <= uopt->gradient_tolerance) { converged = true; break; }
        if (h_rec[8] != 0.0) solved = false;                          // not positive definite
        if (!solved) {
            radius /= decrease; decrease *= 2;
            if (radius < 1e-32) { failed = true; break; }
            continue;
        }
        const double model_decrease = h_rec[4], new_cost = h_rec[7];
        double dc2 = h_rec[2], xc2 = h_rec[3];
        if (focal) {
            xc2 += b->focal[b->cur][0] * b->focal[b->cur][0] + b->focal[b->cur][1] * b->focal[b->cur][1];
            dc2 += df[0] * df[0] + df[1] * df[1];
        }
        const double step_norm = std::sqrt(dc2 + h_rec[5]), x_norm = std::sqrt(xc2 + h_rec[6]);
        if (step_norm <= uopt->parameter_tolerance * (x_norm + uopt->parameter_tolerance)) { converged = true; break; }
        const double rho = model_decrease > 0 ? (cost - new_cost) / model_decrease : -1.0;
        if (uopt->verbose)
            printf("msfm_ba it %d cost %.9e -> %.9e rho %.3f radius %.3e |g|max %.3e\n", it, cost, new_cost, rho, radius, gmax);
        if (rho > 1e-3 && std::isfinite(new_cost)) {
            b->cur = nxt;                                             // accept: the candidate buffers become current
            const double dcost = cost - new_cost;
            cost = new_cost;
            ++good;
            radius = std::min(radius / std::max(1.0 / 3.0, 1.0 - std::pow(2.0 * rho - 1.0, 3)), 1e16);
            decrease = 2.0;
            if (dcost <= uopt->function_tolerance * cost) { converged = true; break; }
        } else {
            radius /= decrease; decrease *= 2;
            if (radius < 1e-32) { failed = true; break; }
        }
    }
    if (multi) {
        // residual count over all ranks
        double v = static_cast<double>(nres_local);
        BA_CUDA(cudaMemcpyAsync(b->small, &v, sizeof(double), cudaMemcpyHostToDevice, c->stream));
        if ((rc = msfm_comm_allreduce_f64(c, b->small, 1, 0))) return rc;
        BA_CUDA(cudaMemcpyAsync(&v, b->small, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        BA_CUDA(cudaStreamSynchronize(c->stream));
        nres_local = static_cast<long long>(v + 0.5);
    }
    sum->iterations = it;
    sum->successful_steps = good;
    sum->termination = failed ? MSFM_BA_FAILURE : (converged ? MSFM_BA_CONVERGENCE : MSFM_BA_NO_CONVERGENCE);
    sum->num_residuals = static_cast<int32_t>(nres_local);
    sum->final_cost = cost;
    sum->total_time_s = std::chrono::duration<double>(clk::now() - t_begin).count();
    sum->linearize_time_s = t_lin;
    sum->solve_time_s = sum->total_time_s - t_lin;
    return MSFM_OK;
}

// ------------------------------------------------------------------------------------------------ communicator
int msfm_comm_unique_id(void* id_out) {
    if (!id_out) return MSFM_E_INVALID;
    if (nccl_load()) return MSFM_E_CUDA;
    return g_nccl.GetUniqueId(id_out) == 0 ? MSFM_OK : MSFM_E_CUDA;
}
int msfm_comm_init(msfm_ctx* c, int32_t n_ranks, int32_t rank, const void* id) {
    if (!c) return MSFM_E_INVALID;
    if (!id || n_ranks < 1 || rank < 0 || rank >= n_ranks) return c->fail(MSFM_E_INVALID, "msfm_comm_init: bad arguments");
    if (const char* err = nccl_load()) return c->fail(MSFM_E_CUDA, "NCCL: %s", err);
    BA_CUDA(cudaSetDevice(c->device));
    if (c->comm) { g_nccl.CommDestroy(c->comm); c->comm = nullptr; }
    Id128 uid;
    std::memcpy(uid.b, id, 128);
    void* comm = nullptr;
    const int r = g_nccl.CommInitRank(&comm, n_ranks, uid, rank);
    if (r != 0) return c->fail(MSFM_E_CUDA, "ncclCommInitRank: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "error");
    c->comm = comm;
    c->comm_ranks = n_ranks;
    c->comm_rank = rank;
    return MSFM_OK;
}
int msfm_comm_destroy(msfm_ctx* c) {
    if (!c) return MSFM_E_INVALID;
    if (c->comm) {
        cudaSetDevice(c->device);
        cudaStreamSynchronize(c->stream);
        g_nccl.CommDestroy(c->comm);
        c->comm = nullptr;
    }
    return MSFM_OK;
}
int msfm_comm_allreduce_f64(msfm_ctx* c, double* buf, int64_t count, int32_t op) {
    if (!c) return MSFM_E_INVALID;
    if (!c->comm) return c->fail(MSFM_E_INVALID, "no communicator attached (msfm_comm_init)");
    if (!buf || count < 0 || (op != 0 && op != 1)) return c->fail(MSFM_E_INVALID, "msfm_comm_allreduce_f64: bad arguments");
    if (count == 0) return MSFM_OK;
    const int r = g_nccl.AllReduce(buf, buf, static_cast<size_t>(count), kNcclFloat64, op == 0 ? kNcclSum : kNcclMax, c->comm, c->stream);
    if (r != 0) return c->fail(MSFM_E_CUDA, "ncclAllReduce: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "error");
    return MSFM_OK;
}

}  // extern "C"
