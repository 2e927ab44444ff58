// Solver of the reduced camera system S dc = rhs inside the LM loop (what Ceres' DENSE_SCHUR / SPARSE_SCHUR linear solvers do
// behind ceres::Solve, src/Optimizer/CeresBundleOptimizer.cpp:264-273,293).
//
// S is block-sparse (ba_tiles.hpp: blk_row / blk_col).  The cameras are renumbered by reverse Cuthill-McKee on the camera
// graph, which turns S into a band of half-width `bw` cameras; grouping bw consecutive cameras into a super-block makes S
// BLOCK-TRIDIAGONAL with dense M x M blocks (M = 6 bw).  The factorisation walks the chain
//     L_ss = chol(D_s),   E_s <- E_s L_ss^-T,   D_{s+1} <- D_{s+1} - E_s E_s^T
// with the dense library routines (cuSOLVER potrf, cuBLAS trsm / syrk / gemm) on the M x M blocks: O(n M^2) flops instead of
// the O(n^3 / 3) of a dense Cholesky of the whole system (for the 1329-camera ring of BASELINE configs[4], 7968 unknowns:
// ~10 GFLOP instead of 170).  When the band is not narrow (bw >= n / 3) the dense path of ba_api.cu is used instead.
// fp64 throughout, like the dense path it replaces.
#include <cublas_v2.h>
#include <cusolverDn.h>

#include <algorithm>
#include <cstdint>
#include <numeric>
#include <queue>
#include <vector>

#include "ba_types.cuh"
#include "launch_count.hpp"

namespace msfm {
namespace ba {

// Reverse Cuthill-McKee ordering of the camera graph (off-diagonal blocks = edges).  Returns pos[f] = new index of free
// camera f, and the half band width (in cameras) of the renumbered block structure.
void rcm_order(int n, const std::vector<int32_t>& blk_row, const std::vector<int32_t>& blk_col, std::vector<int32_t>& pos, int& bw) {
    std::vector<std::vector<int32_t>> adj(static_cast<size_t>(n));
    for (size_t k = 0; k < blk_row.size(); ++k)
        if (blk_row[k] != blk_col[k]) { adj[blk_row[k]].push_back(blk_col[k]); adj[blk_col[k]].push_back(blk_row[k]); }
    std::vector<int32_t> order;
    order.reserve(static_cast<size_t>(n));
    std::vector<char> seen(static_cast<size_t>(n), 0);
    auto bfs = [&](int start, std::vector<int32_t>& out, bool commit) {
        std::vector<char> mark(seen);
        std::queue<int32_t> q;
        q.push(start); mark[start] = 1;
        int last = start;
        while (!q.empty()) {
            const int u = q.front(); q.pop();
            out.push_back(u);
            last = u;
            std::vector<int32_t> nb;
            for (int v : adj[u]) if (!mark[v]) { mark[v] = 1; nb.push_back(v); }
            std::sort(nb.begin(), nb.end(), [&](int a, int b) { return adj[a].size() != adj[b].size() ? adj[a].size() < adj[b].size() : a < b; });
            for (int v : nb) q.push(v);
        }
        if (commit) seen.swap(mark);
        return last;
    };
    for (int s = 0; s < n; ++s) {
        if (seen[s]) continue;
        // pseudo-peripheral start: the last node of a BFS from the component's minimum-degree node, twice
        std::vector<int32_t> comp;
        bfs(s, comp, false);
        int start = comp[0];
        for (int v : comp) if (adj[v].size() < adj[start].size()) start = v;
        for (int rep = 0; rep < 2; ++rep) { std::vector<int32_t> tmp; start = bfs(start, tmp, false); }
        std::vector<int32_t> out;
        bfs(start, out, true);
        order.insert(order.end(), out.begin(), out.end());
    }
    std::reverse(order.begin(), order.end());
    pos.assign(static_cast<size_t>(n), 0);
    for (int i = 0; i < n; ++i) pos[order[i]] = i;
    auto width = [&](const std::vector<int32_t>& ps) {
        int w = 0;
        for (size_t k = 0; k < blk_row.size(); ++k) w = std::max(w, std::abs(ps[blk_row[k]] - ps[blk_col[k]]));
        return w;
    };
    bw = width(pos);
    // Two cheap alternatives that beat BFS level orderings on the graphs sequential capture produces: the cameras' own order
    // (an open chain: neighbours in time see the same points) and the same order FOLDED (0, n-1, 1, n-2, ...: a closed loop,
    // where the first and the last cameras meet again — a ring's band is then twice the co-visibility window instead of a
    // BFS level pair).  The narrowest band wins.
    std::vector<int32_t> cand(static_cast<size_t>(n));
    std::iota(cand.begin(), cand.end(), 0);
    int w = width(cand);
    if (w < bw) { bw = w; pos = cand; }
    for (int i = 0; i < n; ++i) cand[i] = i <= n - 1 - i ? 2 * i : 2 * (n - 1 - i) + 1;
    w = width(cand);
    if (w < bw) { bw = w; pos = cand; }
}

// block slot -> (super-block, local offsets) scatter of the fp32 blocks into the fp64 block-tridiagonal storage
__global__ void expand_tridiag_kernel(Problem P, const int32_t* __restrict__ pos, int m, int M, double inv_radius,
                                      double* __restrict__ D, double* __restrict__ E) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= P.n_blocks * 36) return;
    const int b = idx / 36, e = idx - 36 * b, i = e / 6, j = e - 6 * i;
    const int fa = __ldg(P.blk_row + b), fb = __ldg(P.blk_col + b);
    const int pa = __ldg(pos + fa), pb = __ldg(pos + fb);
    double v = static_cast<double>(P.sblk[idx]);                    // S[fa*6+i][fb*6+j]
    if (fa == fb) {
        if (j > i) return;                                         // the lower triangle of a diagonal block (it is symmetric)
        // take the value from the block's upper triangle, the part every accumulation path fills consistently
        v = static_cast<double>(P.sblk[b * 36 + 6 * j + i]);
        if (i == j) v += fmax(P.tail[P.tl.udiag + fa * 6 + i], 1e-6) * inv_radius;
    }
    // lower triangle of the renumbered matrix: row = the later camera
    int r, c, ri, ci;
    if (pa >= pb) { r = pa; c = pb; ri = i; ci = j; } else { r = pb; c = pa; ri = j; ci = i; }
    const int sr = r / m, sc = c / m;
    const size_t lr = static_cast<size_t>(r - sr * m) * 6 + ri, lc = static_cast<size_t>(c - sc * m) * 6 + ci;
    const size_t MM = static_cast<size_t>(M) * M;
    if (sr == sc) D[sr * MM + lr + lc * M] = v;
    else E[sc * MM + lr + lc * M] = v;                              // sr == sc + 1 by construction of the band
}
__global__ void permute_rhs_kernel(const double* __restrict__ src, const int32_t* __restrict__ pos, int nf, int ld, double* __restrict__ dst) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nf * 6) return;
    const int f = i / 6, k = i - 6 * f;
    dst[static_cast<size_t>(pos[f]) * 6 + k] = src[i];
    (void)ld;
}
__global__ void unpermute_kernel(const double* __restrict__ src, const int32_t* __restrict__ pos, int nf, double* __restrict__ dst) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nf * 6) return;
    const int f = i / 6, k = i - 6 * f;
    dst[i] = src[static_cast<size_t>(pos[f]) * 6 + k];
}

struct TridiagSolver {
    int nf = 0, m = 0, M = 0, n_super = 0, N = 0;
    int32_t* d_pos = nullptr;
    double *D = nullptr, *E = nullptr, *X = nullptr, *work = nullptr;
    int* d_info = nullptr;
    int lwork = 0;
    cublasHandle_t blas = nullptr;
    std::vector<int> h_info;
    int size_of(int s) const { return std::min(M, N - s * M); }
};

}  // namespace ba

using ba::TridiagSolver;

// Returns nullptr when the band is too wide for the chain to pay (the caller keeps the dense path), or on allocation failure
// (*err set).
TridiagSolver* tridiag_create(int nf, const std::vector<int32_t>& blk_row, const std::vector<int32_t>& blk_col, cusolverDnHandle_t solver,
                              cudaStream_t st, cudaError_t* err) {
    *err = cudaSuccess;
    if (nf < 48) return nullptr;
    std::vector<int32_t> pos;
    int bw = 0;
    ba::rcm_order(nf, blk_row, blk_col, pos, bw);
    const int m = std::max(1, bw);
    if (3 * m >= nf) return nullptr;
    TridiagSolver* T = new TridiagSolver();
    T->nf = nf; T->m = m; T->M = 6 * m; T->N = 6 * nf; T->n_super = (nf + m - 1) / m;
    const size_t MM = static_cast<size_t>(T->M) * T->M;
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&T->d_pos), static_cast<size_t>(nf) * sizeof(int32_t));
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&T->D), 2 * static_cast<size_t>(T->n_super) * MM * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&T->X), 3 * static_cast<size_t>(T->N) * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&T->d_info), static_cast<size_t>(T->n_super) * sizeof(int));
    if (e == cudaSuccess) e = cudaMemcpy(T->d_pos, pos.data(), static_cast<size_t>(nf) * sizeof(int32_t), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        T->E = T->D + static_cast<size_t>(T->n_super) * MM;
        if (cusolverDnDpotrf_bufferSize(solver, CUBLAS_FILL_MODE_LOWER, T->M, T->D, T->M, &T->lwork) != CUSOLVER_STATUS_SUCCESS) e = cudaErrorUnknown;
    }
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&T->work), static_cast<size_t>(std::max(1, T->lwork)) * sizeof(double));
    if (e == cudaSuccess && (cublasCreate(&T->blas) != CUBLAS_STATUS_SUCCESS || cublasSetStream(T->blas, st) != CUBLAS_STATUS_SUCCESS)) e = cudaErrorUnknown;
    if (e != cudaSuccess) {
        *err = e;
        if (T->blas) cublasDestroy(T->blas);
        cudaFree(T->d_pos); cudaFree(T->D); cudaFree(T->X); cudaFree(T->d_info); cudaFree(T->work);
        delete T;
        return nullptr;
    }
    T->h_info.assign(static_cast<size_t>(T->n_super), 0);
    return T;
}
void tridiag_destroy(TridiagSolver* T) {
    if (!T) return;
    if (T->blas) cublasDestroy(T->blas);
    cudaFree(T->d_pos); cudaFree(T->D); cudaFree(T->X); cudaFree(T->d_info); cudaFree(T->work);
    delete T;
}
void tridiag_info(const TridiagSolver* T, int32_t out[4]) { out[0] = T->m; out[1] = T->M; out[2] = T->n_super; out[3] = T->nf; }
int* tridiag_dev_info(TridiagSolver* T) { return T->d_info; }
int tridiag_n_super(const TridiagSolver* T) { return T->n_super; }

// Expand S (+ damping) into the chain, factor it, solve for nrhs right-hand sides.  rhs: nrhs columns of N doubles each,
// column stride N (caller's camera order); the solution overwrites it.  All asynchronous on `st`; potrf status words land in
// tridiag_dev_info (non-zero = the system was not positive definite).
cudaError_t tridiag_factor_solve(TridiagSolver* T, const ba::Problem& P, double inv_radius, cusolverDnHandle_t solver, double* rhs, int nrhs,
                                 cudaStream_t st) {
    const int M = T->M, n = T->n_super;
    const size_t MM = static_cast<size_t>(M) * M;
    cudaError_t e = cudaMemsetAsync(T->D, 0, 2 * static_cast<size_t>(n) * MM * sizeof(double), st);
    if (e != cudaSuccess) return e;
    const int tot = P.n_blocks * 36;
    { ba::expand_tridiag_kernel<<<(tot + 255) / 256, 256, 0, st>>>(P, T->d_pos, T->m, M, inv_radius, T->D, T->E); MSFM_COUNT_LAUNCH(); }
    for (int c = 0; c < nrhs; ++c)
        { ba::permute_rhs_kernel<<<(T->N + 255) / 256, 256, 0, st>>>(rhs + static_cast<size_t>(c) * T->N, T->d_pos, T->nf, T->N, T->X + static_cast<size_t>(c) * T->N); MSFM_COUNT_LAUNCH(); }
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    const double one = 1.0, minus = -1.0;
    for (int s = 0; s < n; ++s) {
        const int ms = T->size_of(s);
        double* Ds = T->D + s * MM;
        if (cusolverDnDpotrf(solver, CUBLAS_FILL_MODE_LOWER, ms, Ds, M, T->work, T->lwork, T->d_info + s) != CUSOLVER_STATUS_SUCCESS) return cudaErrorUnknown;
        if (s + 1 < n) {
            const int mn = T->size_of(s + 1);
            double* Es = T->E + s * MM;
            if (cublasDtrsm(T->blas, CUBLAS_SIDE_RIGHT, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_T, CUBLAS_DIAG_NON_UNIT, mn, ms, &one, Ds, M, Es, M) != CUBLAS_STATUS_SUCCESS)
                return cudaErrorUnknown;
            if (cublasDsyrk(T->blas, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, mn, ms, &minus, Es, M, &one, T->D + (s + 1) * MM, M) != CUBLAS_STATUS_SUCCESS)
                return cudaErrorUnknown;
        }
    }
    // forward: y_s = L_ss^-1 (b_s - E_{s-1} y_{s-1});  backward: x_s = L_ss^-T (y_s - E_s^T x_{s+1})
    for (int s = 0; s < n; ++s) {
        const int ms = T->size_of(s);
        double* Xs = T->X + static_cast<size_t>(s) * M;
        if (cublasDtrsm(T->blas, CUBLAS_SIDE_LEFT, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, CUBLAS_DIAG_NON_UNIT, ms, nrhs, &one, T->D + s * MM, M, Xs, T->N) != CUBLAS_STATUS_SUCCESS)
            return cudaErrorUnknown;
        if (s + 1 < n) {
            const int mn = T->size_of(s + 1);
            if (cublasDgemm(T->blas, CUBLAS_OP_N, CUBLAS_OP_N, mn, nrhs, ms, &minus, T->E + s * MM, M, Xs, T->N, &one, Xs + M, T->N) != CUBLAS_STATUS_SUCCESS)
                return cudaErrorUnknown;
        }
    }
    for (int s = n - 1; s >= 0; --s) {
        const int ms = T->size_of(s);
        double* Xs = T->X + static_cast<size_t>(s) * M;
        if (s + 1 < n) {
            const int mn = T->size_of(s + 1);
            if (cublasDgemm(T->blas, CUBLAS_OP_T, CUBLAS_OP_N, ms, nrhs, mn, &minus, T->E + s * MM, M, Xs + M, T->N, &one, Xs, T->N) != CUBLAS_STATUS_SUCCESS)
                return cudaErrorUnknown;
        }
        if (cublasDtrsm(T->blas, CUBLAS_SIDE_LEFT, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_T, CUBLAS_DIAG_NON_UNIT, ms, nrhs, &one, T->D + s * MM, M, Xs, T->N) != CUBLAS_STATUS_SUCCESS)
            return cudaErrorUnknown;
    }
    for (int c = 0; c < nrhs; ++c)
        { ba::unpermute_kernel<<<(T->N + 255) / 256, 256, 0, st>>>(T->X + static_cast<size_t>(c) * T->N, T->d_pos, T->nf, rhs + static_cast<size_t>(c) * T->N); MSFM_COUNT_LAUNCH(); }
    return cudaGetLastError();
}

}  // namespace msfm
