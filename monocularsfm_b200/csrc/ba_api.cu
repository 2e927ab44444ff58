// C-ABI entry points of the B-path (include/msfm_b200.h): problem residency, evaluate / linearize for parity,
// the Levenberg-Marquardt loop that replaces ceres::Solve (src/Optimizer/CeresBundleOptimizer.cpp:293), and the
// NCCL communicator (loaded with dlopen so the library has no link-time NCCL dependency).
#include <cusolverDn.h>
#include <dlfcn.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>
#include <new>
#include <vector>

#include "ba_types.cuh"
#include "ctx.hpp"

namespace msfm {
using ba::CamPre;
using ba::Problem;
cudaError_t ba_launch_cam_prep(const double*, int, CamPre*, cudaStream_t);
cudaError_t ba_launch_evaluate(const Problem&, double*, float*, double*, int, cudaStream_t);
cudaError_t ba_launch_linearize(const Problem&, double, double*, int, cudaStream_t);
cudaError_t ba_launch_track_errors(const Problem&, double*, int, cudaStream_t);
cudaError_t ba_launch_compact_blocks(const int32_t*, long long, int32_t*, int32_t*, int, cudaStream_t);
cudaError_t ba_launch_count_tuples(const Problem&, int32_t*, int, cudaStream_t);
cudaError_t ba_launch_scan_tuples(const int32_t*, long long, int32_t*, int32_t*, cudaStream_t);
cudaError_t ba_launch_fill_tuples(const Problem&, int32_t*, int2*, int, cudaStream_t);
cudaError_t ba_launch_damp(double*, int, double, cudaStream_t);
cudaError_t ba_launch_backsub(const Problem&, double, const double*, double*, double*, int, cudaStream_t);
cudaError_t ba_launch_update_cams(const double*, const int32_t*, int, const double*, double*, cudaStream_t);
cudaError_t ba_launch_copy(const double*, double*, int, cudaStream_t);
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
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce || !g_nccl.CommDestroy) {
        g_nccl.lib = nullptr;
        return "NCCL symbols missing";
    }
    return nullptr;
}
static constexpr int kNcclFloat64 = 8;   // ncclDouble
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
    int32_t *obs_cam = nullptr, *obs_pt = nullptr, *pt_start = nullptr, *cam_free = nullptr;
    double* sys = nullptr;        // S | rhs | gc | udiag | scalars[8]
    double* xsol = nullptr;       // solver right-hand side / solution [n6]
    double* small = nullptr;      // [8] scratch scalars (backsub out[3], new cost, gpmax bits)
    // gather structures + per-linearisation intermediates (ba_types.cuh)
    int32_t *cam_obs_start = nullptr, *cam_obs_list = nullptr, *blk_start = nullptr;
    int2* blk_tuples = nullptr;
    int32_t* blk_list = nullptr;
    int32_t n_blk_list = 0;
    float* obs_J = nullptr;
    double *obs_r = nullptr, *pt_Vinv = nullptr, *pt_gp = nullptr;
    double* work = nullptr;       // cusolver workspace
    int work_len = 0;
    int* dev_info = nullptr;
    int cur = 0;
    std::vector<double> h_cams;   // host mirror of the current cameras
    std::vector<int32_t> h_cam_free;
    // S | rhs | gc | udiag | scalars[8] (| B0 | B1 | focal[16] with a shared focal block)
    size_t sys_len() const { const size_t n6 = size_t(n_free) * 6; return n6 * n6 + 3 * n6 + 8 + (refine_focal ? 2 * n6 + 16 : 0); }
    Problem view(int which) const {
        Problem P;
        P.n_cams = n_cams; P.n_pts = n_pts; P.n_obs = n_obs; P.n_free = n_free;
        P.fx = focal[which][0]; P.fy = focal[which][1];
        P.refine_focal = refine_focal; P.pt_Wf = pt_Wf;
        P.pre = pre[which]; P.pts = pts[which]; P.obs_uv = obs_uv; P.obs_cam = obs_cam; P.obs_pt = obs_pt;
        P.pt_start = pt_start; P.cam_free = cam_free;
        P.gpmax_bits = reinterpret_cast<unsigned long long*>(small + 4);
        P.cam_obs_start = cam_obs_start; P.cam_obs_list = cam_obs_list; P.blk_start = blk_start; P.blk_tuples = blk_tuples;
        P.blk_list = blk_list; P.n_blk_list = n_blk_list;
        P.obs_J = obs_J; P.obs_r = obs_r; P.pt_Vinv = pt_Vinv; P.pt_gp = pt_gp;
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

void msfm_ba_destroy(msfm_ba* b) {
    if (!b) return;
    msfm_ctx* c = b->ctx;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    void* ptrs[] = {b->cams[0], b->cams[1], b->pts[0], b->pts[1], b->pre[0], b->pre[1], b->obs_uv, b->obs_cam, b->obs_pt,
                    b->pt_start, b->cam_free, b->sys, b->xsol, b->small, b->work, b->dev_info, b->cam_obs_start,
                    b->cam_obs_list, b->blk_start, b->blk_tuples, b->blk_list, b->obs_J, b->obs_r, b->pt_Vinv, b->pt_gp, b->pt_Wf};
    for (void* p : ptrs)
        if (p) cudaFree(p);
    delete b;
}

int msfm_ba_create(msfm_ctx* c, const msfm_ba_problem* pr, msfm_ba** out) {
    if (!c) return MSFM_E_INVALID;
    if (!pr || !out) return c->fail(MSFM_E_INVALID, "msfm_ba_create: null argument");
    *out = nullptr;
    if (pr->n_cams <= 0 || pr->n_pts < 0 || pr->n_obs < 0 || (pr->flags & ~MSFM_BA_REFINE_FOCAL) != 0 || !pr->cams || !pr->cam_const ||
        (pr->n_pts > 0 && !pr->pts) || (pr->n_obs > 0 && (!pr->obs_uv || !pr->obs_cam || !pr->obs_pt)))
        return c->fail(MSFM_E_INVALID, "msfm_ba_create: bad problem description");
    BA_CUDA(cudaSetDevice(c->device));
    // validate indices, build CSR over points and the free-camera map on the host
    std::vector<int32_t> pt_start(size_t(pr->n_pts) + 1, 0), cam_free(pr->n_cams, -1);
    int prev = 0;
    for (int i = 0; i < pr->n_obs; ++i) {
        const int p = pr->obs_pt[i], cam = pr->obs_cam[i];
        if (p < prev || p >= pr->n_pts || cam < 0 || cam >= pr->n_cams)
            return c->fail(MSFM_E_INVALID, "msfm_ba_create: observation %d has bad indices (obs_pt must be non-decreasing)", i);
        prev = p;
        pt_start[size_t(p) + 1] += 1;
    }
    for (int p = 0; p < pr->n_pts; ++p) pt_start[size_t(p) + 1] += pt_start[p];
    int nf = 0;
    for (int i = 0; i < pr->n_cams; ++i)
        if (!pr->cam_const[i]) cam_free[i] = nf++;
    // a landmark holds at most one measurement per image (tracks in Map.cpp are keyed by image): the gather lists rely on it
    for (int p = 0; p < pr->n_pts; ++p)
        for (int a = pt_start[p]; a < pt_start[size_t(p) + 1]; ++a)
            for (int q = a + 1; q < pt_start[size_t(p) + 1]; ++q)
                if (pr->obs_cam[a] == pr->obs_cam[q])
                    return c->fail(MSFM_E_INVALID, "msfm_ba_create: point %d is observed twice by camera %d", p, pr->obs_cam[a]);
    // CSR of the observations of every free camera (counting sort)
    std::vector<int32_t> cam_obs_start(size_t(nf) + 1, 0), cam_obs_list;
    for (int i = 0; i < pr->n_obs; ++i) {
        const int f = cam_free[pr->obs_cam[i]];
        if (f >= 0) cam_obs_start[size_t(f) + 1] += 1;
    }
    for (int f = 0; f < nf; ++f) cam_obs_start[size_t(f) + 1] += cam_obs_start[f];
    cam_obs_list.resize(size_t(cam_obs_start[nf]));
    {
        std::vector<int32_t> cur(cam_obs_start.begin(), cam_obs_start.end() - 1);
        for (int i = 0; i < pr->n_obs; ++i) {
            const int f = cam_free[pr->obs_cam[i]];
            if (f >= 0) cam_obs_list[size_t(cur[f]++)] = i;
        }
    }
    msfm_ba* b = new (std::nothrow) msfm_ba();
    if (!b) return c->fail(MSFM_E_CUDA, "out of host memory");
    b->ctx = c;
    b->n_cams = pr->n_cams; b->n_pts = pr->n_pts; b->n_obs = pr->n_obs; b->n_free = nf;
    b->refine_focal = (pr->flags & MSFM_BA_REFINE_FOCAL) ? 1 : 0;
    b->focal[0][0] = b->focal[1][0] = pr->fx; b->focal[0][1] = b->focal[1][1] = pr->fy;
    b->h_cams.assign(pr->cams, pr->cams + size_t(pr->n_cams) * 6);
    b->h_cam_free = cam_free;
    const size_t n6 = size_t(nf) * 6;
    auto fail_free = [&](int rc) { msfm_ba_destroy(b); return rc; };
#define BA_ALLOC(ptr, bytes)                                                         \
    do {                                                                             \
        cudaError_t e__ = cudaMalloc(reinterpret_cast<void**>(&(ptr)), std::max<size_t>(16, (bytes))); \
        if (e__ != cudaSuccess) return fail_free(c->cuda_fail(e__, "cudaMalloc(" #ptr ")")); \
    } while (0)
    for (int k = 0; k < 2; ++k) {
        BA_ALLOC(b->cams[k], size_t(pr->n_cams) * 6 * sizeof(double));
        BA_ALLOC(b->pts[k], size_t(pr->n_pts) * 3 * sizeof(double));
        BA_ALLOC(b->pre[k], size_t(pr->n_cams) * sizeof(CamPre));
    }
    BA_ALLOC(b->obs_uv, size_t(pr->n_obs) * 2 * sizeof(double));
    BA_ALLOC(b->obs_cam, size_t(pr->n_obs) * sizeof(int32_t));
    BA_ALLOC(b->obs_pt, size_t(pr->n_obs) * sizeof(int32_t));
    BA_ALLOC(b->pt_start, (size_t(pr->n_pts) + 1) * sizeof(int32_t));
    BA_ALLOC(b->cam_free, size_t(pr->n_cams) * sizeof(int32_t));
    BA_ALLOC(b->sys, b->sys_len() * sizeof(double));
    BA_ALLOC(b->xsol, (3 * std::max<size_t>(1, n6) + 2) * sizeof(double));      // up to 3 right-hand sides + (d fx, d fy)
    if (b->refine_focal) BA_ALLOC(b->pt_Wf, std::max<size_t>(1, size_t(pr->n_pts)) * 6 * sizeof(double));
    BA_ALLOC(b->small, 8 * sizeof(double));
    BA_ALLOC(b->dev_info, sizeof(int));
    BA_ALLOC(b->cam_obs_start, cam_obs_start.size() * sizeof(int32_t));
    BA_ALLOC(b->cam_obs_list, cam_obs_list.size() * sizeof(int32_t));
    BA_ALLOC(b->blk_start, (size_t(nf) * nf + 1) * sizeof(int32_t));
    BA_ALLOC(b->obs_J, size_t(pr->n_obs) * 18 * sizeof(float));
    BA_ALLOC(b->obs_r, size_t(pr->n_obs) * 2 * sizeof(double));
    BA_ALLOC(b->pt_Vinv, size_t(pr->n_pts) * 6 * sizeof(double));
    BA_ALLOC(b->pt_gp, size_t(pr->n_pts) * 3 * sizeof(double));
#undef BA_ALLOC
    auto H2D = [&](void* dst, const void* src, size_t bytes) {
        return bytes ? cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice) : cudaSuccess;
    };
    cudaError_t e = cudaSuccess;
    if (e == cudaSuccess) e = H2D(b->cams[0], pr->cams, size_t(pr->n_cams) * 6 * sizeof(double));
    if (e == cudaSuccess) e = H2D(b->pts[0], pr->pts, size_t(pr->n_pts) * 3 * sizeof(double));
    if (e == cudaSuccess) e = H2D(b->obs_uv, pr->obs_uv, size_t(pr->n_obs) * 2 * sizeof(double));
    if (e == cudaSuccess) e = H2D(b->obs_cam, pr->obs_cam, size_t(pr->n_obs) * sizeof(int32_t));
    if (e == cudaSuccess) e = H2D(b->obs_pt, pr->obs_pt, size_t(pr->n_obs) * sizeof(int32_t));
    if (e == cudaSuccess) e = H2D(b->pt_start, pt_start.data(), pt_start.size() * sizeof(int32_t));
    if (e == cudaSuccess) e = H2D(b->cam_free, cam_free.data(), cam_free.size() * sizeof(int32_t));
    if (e == cudaSuccess) e = H2D(b->cam_obs_start, cam_obs_start.data(), cam_obs_start.size() * sizeof(int32_t));
    if (e == cudaSuccess) e = H2D(b->cam_obs_list, cam_obs_list.data(), cam_obs_list.size() * sizeof(int32_t));
    if (e != cudaSuccess) return fail_free(c->cuda_fail(e, "msfm_ba_create H2D"));
    // co-observation tuples of every camera pair, built on the device: count -> scan -> fill
    {
        const long long nblk = static_cast<long long>(nf) * nf;
        int32_t *counts = nullptr, *cursor = nullptr;
        const size_t cbytes = std::max<size_t>(16, size_t(nblk) * sizeof(int32_t));
        if ((e = cudaMalloc(&counts, cbytes)) == cudaSuccess) e = cudaMalloc(&cursor, cbytes);
        if (e == cudaSuccess) e = cudaMemsetAsync(counts, 0, cbytes, c->stream);
        const Problem P0 = b->view(0);
        if (e == cudaSuccess) e = ba_launch_count_tuples(P0, counts, c->num_sms, c->stream);
        if (e == cudaSuccess) e = ba_launch_scan_tuples(counts, nblk, b->blk_start, cursor, c->stream);
        int32_t total = 0;
        if (e == cudaSuccess) e = cudaMemcpyAsync(&total, b->blk_start + nblk, sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&b->blk_tuples), std::max<size_t>(16, size_t(total) * sizeof(int2)));
        if (e == cudaSuccess) e = ba_launch_fill_tuples(b->view(0), cursor, b->blk_tuples, c->num_sms, c->stream);
        // compact list of the non-empty blocks (at most one per tuple); `counts` is free again and serves as the counter
        const size_t max_list = static_cast<size_t>(std::min<long long>(nblk, total));
        if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&b->blk_list), std::max<size_t>(16, max_list * sizeof(int32_t)));
        if (e == cudaSuccess) e = cudaMemsetAsync(counts, 0, sizeof(int32_t), c->stream);
        if (e == cudaSuccess) e = ba_launch_compact_blocks(b->blk_start, nblk, b->blk_list, counts, c->num_sms, c->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(&b->n_blk_list, counts, sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (counts) cudaFree(counts);
        if (cursor) cudaFree(cursor);
        c->launches += 4;
        if (e != cudaSuccess) return fail_free(c->cuda_fail(e, "msfm_ba_create: building the gather lists"));
    }
    *out = b;
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
    c->launches += 1;
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
    c->launches += 1;
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
    c->launches += 1;
    BA_CUDA(cudaMemcpyAsync(err, c->d_ba_r.p, size_t(b->n_pts) * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    BA_CUDA(cudaStreamSynchronize(c->stream));
    return MSFM_OK;
}

// zero sys, linearize + Schur at parameter set `which`, all-reduce.  Leaves S undamped on the camera diagonal
// (udiag is summed across ranks first); the caller applies damp afterwards.
static int linearize(msfm_ba* b, int which, double inv_radius) {
    msfm_ctx* c = b->ctx;
    BA_CUDA(cudaMemsetAsync(b->sys, 0, b->sys_len() * sizeof(double), c->stream));
    BA_CUDA(cudaMemsetAsync(b->small, 0, 8 * sizeof(double), c->stream));
    c->prof_begin(MSFM_PROF_BA_SCHUR);
    BA_CUDA(ba_launch_linearize(b->view(which), inv_radius, b->sys, c->num_sms, c->stream));
    c->prof_end();
    c->launches += 3;
    if (c->comm) {
        int rc = msfm_comm_allreduce_f64(c, b->sys, static_cast<int64_t>(b->sys_len()), 0);
        if (rc) return rc;
        rc = msfm_comm_allreduce_f64(c, b->small + 4, 1, 1);     // max |g_p| (non-negative: bits order = value order)
        if (rc) return rc;
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
    const int n6 = b->n_free * 6;
    BA_CUDA(ba_launch_damp(b->sys, n6, inv_radius, c->stream));
    c->launches += 1;
    const size_t N = size_t(n6);
    std::vector<double> hS, tail(3 * N + 8);
    if (S) {
        hS.resize(N * N);
        BA_CUDA(cudaMemcpyAsync(hS.data(), b->sys, N * N * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    }
    BA_CUDA(cudaMemcpyAsync(tail.data(), b->sys + N * N, tail.size() * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    BA_CUDA(cudaStreamSynchronize(c->stream));
    if (S) {
        for (size_t i = 0; i < N; ++i)
            for (size_t j = 0; j < N; ++j) {
                const size_t bi = i / 6, bj = j / 6;
                S[i * N + j] = (bi <= bj) ? hS[i * N + j] : hS[j * N + i];     // device holds the upper block triangle
            }
    }
    if (rhs) std::memcpy(rhs, tail.data(), N * sizeof(double));
    if (gc) std::memcpy(gc, tail.data() + N, N * sizeof(double));
    if (cost) *cost = tail[3 * N];
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
    std::vector<double> tail(2 * N + 16);
    BA_CUDA(cudaMemcpyAsync(tail.data(), b->sys + N * N + 3 * N + 8, tail.size() * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    BA_CUDA(cudaStreamSynchronize(c->stream));
    const double* ff = tail.data() + 2 * N;
    if (B)
        for (size_t i = 0; i < N; ++i) { B[2 * i] = tail[i]; B[2 * i + 1] = tail[N + i]; }
    if (F) {
        F[0] = ff[0] + std::max(ff[7], 1e-6) * inv_radius;
        F[1] = ff[1];
        F[2] = ff[2] + std::max(ff[8], 1e-6) * inv_radius;
    }
    if (rhs_f) { rhs_f[0] = ff[3]; rhs_f[1] = ff[4]; }
    if (g_f) { g_f[0] = ff[5]; g_f[1] = ff[6]; }
    return MSFM_OK;
}

int msfm_ba_solve(msfm_ba* b, const msfm_ba_options* uopt, msfm_ba_summary* sum) {
    if (!b) return MSFM_E_INVALID;
    msfm_ctx* c = b->ctx;
    if (!uopt || !sum) return c->fail(MSFM_E_INVALID, "msfm_ba_solve: null argument");
    BA_CUDA(cudaSetDevice(c->device));
    int rc = ensure_solver(c);
    if (rc) return rc;
    cusolverDnHandle_t solver = static_cast<cusolverDnHandle_t>(c->cusolver);
    using clk = std::chrono::steady_clock;
    const auto t_begin = clk::now();
    double t_lin = 0, t_sol = 0;
    const int n6 = b->n_free * 6;
    const size_t N = size_t(n6);
    if (n6 > 0) {
        int lwork = 0;
        if (cusolverDnDpotrf_bufferSize(solver, CUBLAS_FILL_MODE_LOWER, n6, b->sys, n6, &lwork) != CUSOLVER_STATUS_SUCCESS)
            return c->fail(MSFM_E_CUDA, "cusolverDnDpotrf_bufferSize failed");
        if (lwork > b->work_len) {
            if (b->work) cudaFree(b->work);
            b->work = nullptr;
            BA_CUDA(cudaMalloc(&b->work, size_t(lwork) * sizeof(double)));
            b->work_len = lwork;
        }
    }
    std::memset(sum, 0, sizeof *sum);
    double radius = uopt->initial_trust_region_radius;
    double decrease = 2.0;
    double cost = 0.0;
    bool have_cost = false, converged = false, failed = false;
    long long nres_local = 2LL * b->n_obs;
    const bool focal = b->refine_focal != 0;
    const int nrhs = focal ? 3 : 1;
    std::vector<double> h_tail(3 * N + 8 + (focal ? 2 * N + 16 : 0)), h_dc(N + 2), h_small(8), h_x(focal ? 3 * N : 0);
    int it = 0, good = 0;
    if ((rc = prep(b, b->cur))) return rc;
    while (it < uopt->max_num_iterations) {
        ++it;
        const double inv_radius = 1.0 / radius;
        auto t0 = clk::now();
        if ((rc = linearize(b, b->cur, inv_radius))) return rc;
        // tail of sys: rhs | gc | udiag | scalars ; plus max |g_p|
        BA_CUDA(cudaMemcpyAsync(h_tail.data(), b->sys + N * N, h_tail.size() * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        BA_CUDA(cudaMemcpyAsync(h_small.data(), b->small, 8 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        BA_CUDA(cudaStreamSynchronize(c->stream));
        t_lin += std::chrono::duration<double>(clk::now() - t0).count();
        const double lin_cost = h_tail[3 * N];
        if (!have_cost) { cost = lin_cost; sum->initial_cost = cost; have_cost = true; }
        double gmax = h_small[4];
        for (size_t i = 0; i < N; ++i) gmax = std::max(gmax, std::fabs(h_tail[N + i]));
        const double* hB = h_tail.data() + 3 * N + 8;          // B0 | B1 (focal only)
        const double* hff = hB + 2 * N;                        // F00 F01 F11 rhsf0 rhsf1 gf0 gf1 uf0 uf1
        if (focal) gmax = std::max(gmax, std::max(std::fabs(hff[5]), std::fabs(hff[6])));
        if (gmax <= uopt->gradient_tolerance) { converged = true; break; }
        // ---- solve the reduced camera system (dense Cholesky; the row-major upper block triangle written by the
        //      kernel is the column-major lower triangle cuSOLVER reads)
        t0 = clk::now();
        bool solved = true;
        if (n6 > 0) {
            BA_CUDA(ba_launch_damp(b->sys, n6, inv_radius, c->stream));
            BA_CUDA(ba_launch_copy(b->sys + N * N, b->xsol, n6, c->stream));
            if (focal) {       // two more right-hand sides: the border columns (S^-1 B for the 2 x 2 Schur complement below)
                BA_CUDA(cudaMemcpyAsync(b->xsol + N, b->sys + N * N + 3 * N + 8, 2 * N * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
            }
            c->launches += 2;
            if (cusolverDnDpotrf(solver, CUBLAS_FILL_MODE_LOWER, n6, b->sys, n6, b->work, b->work_len, b->dev_info) != CUSOLVER_STATUS_SUCCESS)
                return c->fail(MSFM_E_CUDA, "cusolverDnDpotrf failed to launch");
            int info = 0;
            BA_CUDA(cudaMemcpyAsync(&info, b->dev_info, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
            BA_CUDA(cudaStreamSynchronize(c->stream));
            if (info != 0) {
                solved = false;
            } else {
                if (cusolverDnDpotrs(solver, CUBLAS_FILL_MODE_LOWER, n6, nrhs, b->sys, n6, b->xsol, n6, b->dev_info) != CUSOLVER_STATUS_SUCCESS)
                    return c->fail(MSFM_E_CUDA, "cusolverDnDpotrs failed to launch");
            }
        }
        double df[2] = {0.0, 0.0};
        if (focal && solved) {
            // bordered system [S B; B^T F] [dc; df] = [rhs; rhs_f]:  (F - B^T S^-1 B) df = rhs_f - B^T S^-1 rhs,  dc = S^-1 rhs - S^-1 B df
            if (n6 > 0) {
                BA_CUDA(cudaMemcpyAsync(h_x.data(), b->xsol, 3 * N * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
                BA_CUDA(cudaStreamSynchronize(c->stream));
            }
            double f00 = hff[0] + std::max(hff[7], 1e-6) * inv_radius, f01 = hff[1], f11 = hff[2] + std::max(hff[8], 1e-6) * inv_radius;
            double r0 = hff[3], r1 = hff[4];
            const double *y = h_x.data(), *x0 = h_x.data() + N, *x1 = h_x.data() + 2 * N;
            for (size_t i = 0; i < N; ++i) {
                f00 -= hB[i] * x0[i]; f01 -= hB[i] * x1[i]; f11 -= hB[N + i] * x1[i];
                r0 -= hB[i] * y[i];   r1 -= hB[N + i] * y[i];
            }
            const double det = f00 * f11 - f01 * f01;
            if (!(det > 0.0) || !(f00 > 0.0)) {
                solved = false;
            } else {
                df[0] = (f11 * r0 - f01 * r1) / det;
                df[1] = (f00 * r1 - f01 * r0) / det;
                for (size_t i = 0; i < N; ++i) h_dc[i] = y[i] - x0[i] * df[0] - x1[i] * df[1];
                h_dc[N] = df[0]; h_dc[N + 1] = df[1];
                BA_CUDA(cudaMemcpyAsync(b->xsol, h_dc.data(), (N + 2) * sizeof(double), cudaMemcpyHostToDevice, c->stream));
            }
        }
        if (!solved) {
            t_sol += std::chrono::duration<double>(clk::now() - t0).count();
            radius /= decrease; decrease *= 2;
            if (radius < 1e-32) { failed = true; break; }
            continue;
        }
        // ---- back-substitute the points, build the candidate
        const int nxt = b->cur ^ 1;
        BA_CUDA(cudaMemsetAsync(b->small, 0, 4 * sizeof(double), c->stream));
        c->prof_begin(MSFM_PROF_BA_OTHER);
        BA_CUDA(ba_launch_backsub(b->view(b->cur), inv_radius, b->xsol, b->pts[nxt], b->small, c->num_sms, c->stream));
        BA_CUDA(ba_launch_update_cams(b->cams[b->cur], b->cam_free, b->n_cams, b->xsol, b->cams[nxt], c->stream));
        c->prof_end();
        c->launches += 2;
        b->focal[nxt][0] = b->focal[b->cur][0] + df[0];
        b->focal[nxt][1] = b->focal[b->cur][1] + df[1];
        if ((rc = prep(b, nxt))) return rc;
        c->prof_begin(MSFM_PROF_BA_EVAL);
        BA_CUDA(ba_launch_evaluate(b->view(nxt), nullptr, nullptr, b->small + 3, c->num_sms, c->stream));
        c->prof_end();
        c->launches += 1;
        if (c->comm && (rc = msfm_comm_allreduce_f64(c, b->small, 4, 0))) return rc;
        BA_CUDA(cudaMemcpyAsync(h_small.data(), b->small, 4 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        if (n6 > 0 && !focal) BA_CUDA(cudaMemcpyAsync(h_dc.data(), b->xsol, N * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        BA_CUDA(cudaStreamSynchronize(c->stream));
        t_sol += std::chrono::duration<double>(clk::now() - t0).count();
        const double model_decrease = h_small[0], new_cost = h_small[3];
        double dc2 = 0, xc2 = 0;
        for (size_t i = 0; i < N; ++i) dc2 += h_dc[i] * h_dc[i];
        for (int cam = 0; cam < b->n_cams; ++cam)
            if (b->h_cam_free[cam] >= 0)
                for (int k = 0; k < 6; ++k) xc2 += b->h_cams[size_t(cam) * 6 + k] * b->h_cams[size_t(cam) * 6 + k];
        if (focal) {
            dc2 += df[0] * df[0] + df[1] * df[1];
            xc2 += b->focal[b->cur][0] * b->focal[b->cur][0] + b->focal[b->cur][1] * b->focal[b->cur][1];
        }
        const double step_norm = std::sqrt(dc2 + h_small[1]), x_norm = std::sqrt(xc2 + h_small[2]);
        if (step_norm <= uopt->parameter_tolerance * (x_norm + uopt->parameter_tolerance)) { converged = true; break; }
        const double rho = model_decrease > 0 ? (cost - new_cost) / model_decrease : -1.0;
        if (uopt->verbose)
            printf("msfm_ba it %d cost %.9e -> %.9e rho %.3f radius %.3e |g|max %.3e\n", it, cost, new_cost, rho, radius, gmax);
        if (rho > 1e-3 && std::isfinite(new_cost)) {
            // accept: the candidate buffers become current
            for (int cam = 0; cam < b->n_cams; ++cam) {
                const int f = b->h_cam_free[cam];
                if (f >= 0)
                    for (int k = 0; k < 6; ++k) b->h_cams[size_t(cam) * 6 + k] += h_dc[size_t(f) * 6 + k];
            }
            b->cur = nxt;
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
    if (c->comm) {
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
    sum->linearize_time_s = t_lin;
    sum->solve_time_s = t_sol;
    sum->total_time_s = std::chrono::duration<double>(clk::now() - t_begin).count();
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
