/* B-path oracle in plain C (TEST INFRASTRUCTURE — only tests/ and bench.py's cpu_baseline leg may use it).
 *
 * Restates, in float64 and on one thread (the reference never sets Ceres' num_threads,
 * CeresBundleOptimizer.cpp:262-291), what one Ceres iteration does for this problem class:
 *   - evaluate every residual block with 9-wide Jets: the functor of CeresBundleOptimizer.cpp:29-53 with Ceres'
 *     AngleAxisRotatePoint (rotation.h) including its Taylor branch;
 *   - eliminate the points (e-blocks, 3) onto the free cameras (f-blocks, 6): S, rhs of the reduced camera system with
 *     Marquardt damping diag(J^T J)/radius (what Ceres' LM strategy + Jacobi scaling amounts to).
 * Same conventions as oracle/ba_oracle.py (which pins it in tests/test_oracle_ba.py::test_c_oracle_matches_python).
 * Parity with Ceres itself is UNPINNED: Ceres is not available in this image (see ba_oracle.py docstring).
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ND 9
typedef struct { double a; double v[ND]; } jet;

static jet jc(double a) { jet r; r.a = a; memset(r.v, 0, sizeof r.v); return r; }
static jet jvar(double a, int k) { jet r = jc(a); r.v[k] = 1.0; return r; }
static jet jadd(jet x, jet y) { jet r; r.a = x.a + y.a; for (int i = 0; i < ND; ++i) r.v[i] = x.v[i] + y.v[i]; return r; }
static jet jsub(jet x, jet y) { jet r; r.a = x.a - y.a; for (int i = 0; i < ND; ++i) r.v[i] = x.v[i] - y.v[i]; return r; }
static jet jmul(jet x, jet y) { jet r; r.a = x.a * y.a; for (int i = 0; i < ND; ++i) r.v[i] = x.a * y.v[i] + y.a * x.v[i]; return r; }
static jet jdiv(jet x, jet y) { jet r; double inv = 1.0 / y.a; r.a = x.a * inv; for (int i = 0; i < ND; ++i) r.v[i] = (x.v[i] - r.a * y.v[i]) * inv; return r; }
static jet jsqrt(jet x) { jet r; r.a = sqrt(x.a); double s = 1.0 / (2.0 * r.a); for (int i = 0; i < ND; ++i) r.v[i] = x.v[i] * s; return r; }
static jet jsin(jet x) { jet r; r.a = sin(x.a); double c = cos(x.a); for (int i = 0; i < ND; ++i) r.v[i] = c * x.v[i]; return r; }
static jet jcos(jet x) { jet r; r.a = cos(x.a); double s = -sin(x.a); for (int i = 0; i < ND; ++i) r.v[i] = s * x.v[i]; return r; }

/* ceres::AngleAxisRotatePoint */
static void rotate(const jet w[3], const jet pt[3], jet out[3]) {
    jet th2 = jadd(jadd(jmul(w[0], w[0]), jmul(w[1], w[1])), jmul(w[2], w[2]));
    if (th2.a > 2.220446049250313e-16) {
        jet th = jsqrt(th2), c = jcos(th), s = jsin(th), ti = jdiv(jc(1.0), th);
        jet n[3] = {jmul(w[0], ti), jmul(w[1], ti), jmul(w[2], ti)};
        jet x[3] = {jsub(jmul(n[1], pt[2]), jmul(n[2], pt[1])), jsub(jmul(n[2], pt[0]), jmul(n[0], pt[2])),
                    jsub(jmul(n[0], pt[1]), jmul(n[1], pt[0]))};
        jet tmp = jmul(jadd(jadd(jmul(n[0], pt[0]), jmul(n[1], pt[1])), jmul(n[2], pt[2])), jsub(jc(1.0), c));
        for (int k = 0; k < 3; ++k) out[k] = jadd(jadd(jmul(pt[k], c), jmul(x[k], s)), jmul(n[k], tmp));
    } else {
        jet x[3] = {jsub(jmul(w[1], pt[2]), jmul(w[2], pt[1])), jsub(jmul(w[2], pt[0]), jmul(w[0], pt[2])),
                    jsub(jmul(w[0], pt[1]), jmul(w[1], pt[0]))};
        for (int k = 0; k < 3; ++k) out[k] = jadd(pt[k], x[k]);
    }
}

/* one residual block: r[2], J[2][9] (rvec | tvec | point) */
static void residual(const double* cam, const double* X, double u, double v, double fx, double fy, double r[2], double J[18]) {
    jet w[3], t[3], p[3], q[3];
    for (int k = 0; k < 3; ++k) { w[k] = jvar(cam[k], k); t[k] = jvar(cam[3 + k], 3 + k); p[k] = jvar(X[k], 6 + k); }
    rotate(w, p, q);
    for (int k = 0; k < 3; ++k) q[k] = jadd(q[k], t[k]);
    jet xp = jdiv(q[0], q[2]), yp = jdiv(q[1], q[2]);
    jet rx = jsub(jmul(jc(fx), xp), jc(u)), ry = jsub(jmul(jc(fy), yp), jc(v));
    r[0] = rx.a; r[1] = ry.a;
    for (int i = 0; i < ND; ++i) { J[i] = rx.v[i]; J[9 + i] = ry.v[i]; }
}

/* Evaluate all residual blocks.  r [n_obs][2], J [n_obs][2][9] (either may be NULL).  Returns the cost. */
double ba_oracle_evaluate(int n_obs, const double* cams, const double* pts, const double* obs_uv, const int32_t* obs_cam,
                          const int32_t* obs_pt, double fx, double fy, double* r_out, double* J_out) {
    double cost = 0.0;
    for (int i = 0; i < n_obs; ++i) {
        double r[2], J[18];
        residual(cams + 6 * obs_cam[i], pts + 3 * obs_pt[i], obs_uv[2 * i], obs_uv[2 * i + 1], fx, fy, r, J);
        cost += 0.5 * (r[0] * r[0] + r[1] * r[1]);
        if (r_out) { r_out[2 * i] = r[0]; r_out[2 * i + 1] = r[1]; }
        if (J_out) memcpy(J_out + (size_t)i * 18, J, sizeof J);
    }
    return cost;
}

static void inv3(const double m[9], double o[9]) {
    double c00 = m[4] * m[8] - m[5] * m[7], c01 = m[5] * m[6] - m[3] * m[8], c02 = m[3] * m[7] - m[4] * m[6];
    double det = m[0] * c00 + m[1] * c01 + m[2] * c02, id = 1.0 / det;
    o[0] = c00 * id; o[1] = (m[2] * m[7] - m[1] * m[8]) * id; o[2] = (m[1] * m[5] - m[2] * m[4]) * id;
    o[3] = c01 * id; o[4] = (m[0] * m[8] - m[2] * m[6]) * id; o[5] = (m[2] * m[3] - m[0] * m[5]) * id;
    o[6] = c02 * id; o[7] = (m[1] * m[6] - m[0] * m[7]) * id; o[8] = (m[0] * m[4] - m[1] * m[3]) * id;
}

/* dst += val; in the multi-threaded pass as an atomic update (compare-and-swap on the bit pattern) */
static inline void atomic_add_double(double* p, double v) {
    uint64_t* q = (uint64_t*)p;
    uint64_t old = __atomic_load_n(q, __ATOMIC_RELAXED), neu;
    do {
        double d;
        memcpy(&d, &old, 8);
        d += v;
        memcpy(&neu, &d, 8);
    } while (!__atomic_compare_exchange_n(q, &old, neu, 1, __ATOMIC_RELAXED, __ATOMIC_RELAXED));
}
#define ACC(dst, val) do { if (atomic) atomic_add_double(&(dst), (val)); else (dst) += (val); } while (0)

/* Eliminates the point whose observations are [i, e): its contribution to S, rhs, udiag.  Jb / rb: scratch for e - i
 * residual blocks.  Returns the point's share of the cost. */
static double eliminate_point(int i, int e, const double* cams, const double* pts, const double* obs_uv, const int32_t* obs_cam,
                              const int32_t* obs_pt, const int32_t* cam_free, size_t N, double fx, double fy, double inv_radius,
                              double* S, double* rhs, double* udiag, double* Jb, double* rb, int atomic) {
    const int p = obs_pt[i], k = e - i;
    double cost = 0.0;
    double V[9] = {0}, gp[3] = {0};
    for (int a = 0; a < k; ++a) {
        double* J = Jb + (size_t)a * 18; double* r = rb + 2 * a;
        residual(cams + 6 * obs_cam[i + a], pts + 3 * p, obs_uv[2 * (i + a)], obs_uv[2 * (i + a) + 1], fx, fy, r, J);
        cost += 0.5 * (r[0] * r[0] + r[1] * r[1]);
        for (int x = 0; x < 3; ++x) {
            gp[x] += J[6 + x] * r[0] + J[15 + x] * r[1];
            for (int y = 0; y < 3; ++y) V[3 * x + y] += J[6 + x] * J[6 + y] + J[15 + x] * J[15 + y];
        }
    }
    for (int x = 0; x < 3; ++x) V[4 * x] += fmax(V[4 * x], 1e-6) * inv_radius;
    double Vi[9];
    inv3(V, Vi);
    for (int a = 0; a < k; ++a) {
        int fa = cam_free[obs_cam[i + a]];
        if (fa < 0) continue;
        const double* Ja = Jb + (size_t)a * 18; const double* ra = rb + 2 * a;
        double W[18], Y[18];
        for (int x = 0; x < 6; ++x)
            for (int y = 0; y < 3; ++y) W[3 * x + y] = Ja[x] * Ja[6 + y] + Ja[9 + x] * Ja[15 + y];
        for (int x = 0; x < 6; ++x)
            for (int y = 0; y < 3; ++y) Y[3 * x + y] = W[3 * x] * Vi[y] + W[3 * x + 1] * Vi[3 + y] + W[3 * x + 2] * Vi[6 + y];
        for (int x = 0; x < 6; ++x) {
            double jr = Ja[x] * ra[0] + Ja[9 + x] * ra[1];
            ACC(rhs[6 * fa + x], Y[3 * x] * gp[0] + Y[3 * x + 1] * gp[1] + Y[3 * x + 2] * gp[2] - jr);
            ACC(udiag[6 * fa + x], Ja[x] * Ja[x] + Ja[9 + x] * Ja[9 + x]);
            for (int y = 0; y < 6; ++y) ACC(S[(6 * (size_t)fa + x) * N + 6 * fa + y], Ja[x] * Ja[y] + Ja[9 + x] * Ja[9 + y]);
        }
        for (int b = 0; b < k; ++b) {
            int fb = cam_free[obs_cam[i + b]];
            if (fb < 0) continue;
            const double* Jq = Jb + (size_t)b * 18;
            for (int x = 0; x < 6; ++x)
                for (int y = 0; y < 6; ++y) {
                    double wb0 = Jq[y] * Jq[6] + Jq[9 + y] * Jq[15], wb1 = Jq[y] * Jq[7] + Jq[9 + y] * Jq[16], wb2 = Jq[y] * Jq[8] + Jq[9 + y] * Jq[17];
                    ACC(S[(6 * (size_t)fa + x) * N + 6 * fb + y], -(Y[3 * x] * wb0 + Y[3 * x + 1] * wb1 + Y[3 * x + 2] * wb2));
                }
        }
    }
    return cost;
}

/* One "iteration" = evaluate + Schur-eliminate.  obs grouped by point (obs_pt non-decreasing).
 * S [6F][6F] dense, rhs [6F]; cam_free[n_cams] = reduced index or -1.  Returns the cost. */
double ba_oracle_linearize(int n_cams, int n_pts, int n_obs, const double* cams, const double* pts, const double* obs_uv,
                           const int32_t* obs_cam, const int32_t* obs_pt, const int32_t* cam_free, int n_free, double fx,
                           double fy, double inv_radius, double* S, double* rhs) {
    const size_t N = (size_t)n_free * 6;
    memset(S, 0, N * N * sizeof(double));
    memset(rhs, 0, N * sizeof(double));
    double* udiag = (double*)calloc(N ? N : 1, sizeof(double));
    int cap = 64;
    double* Jb = (double*)malloc((size_t)cap * 18 * sizeof(double));
    double* rb = (double*)malloc((size_t)cap * 2 * sizeof(double));
    double cost = 0.0;
    int i = 0;
    (void)n_cams; (void)n_pts;
    while (i < n_obs) {
        int p = obs_pt[i], e = i;
        while (e < n_obs && obs_pt[e] == p) ++e;
        int k = e - i;
        if (k > cap) { cap = 2 * k; Jb = (double*)realloc(Jb, (size_t)cap * 18 * sizeof(double)); rb = (double*)realloc(rb, (size_t)cap * 2 * sizeof(double)); }
        cost += eliminate_point(i, e, cams, pts, obs_uv, obs_cam, obs_pt, cam_free, N, fx, fy, inv_radius, S, rhs, udiag, Jb, rb, 0);
        i = e;
    }
    for (size_t d = 0; d < N; ++d) S[d * N + d] += fmax(udiag[d], 1e-6) * inv_radius;
    free(udiag); free(Jb); free(rb);
    return cost;
}

/* The same pass on n_threads host threads (pthreads): points are handed out dynamically in chunks of 64, every thread
 * evaluates its points' residual blocks and adds its products into the shared S / rhs with atomic updates.  This is NOT what
 * the reference does (it never sets Ceres' num_threads): bench.py reports it beside the one-thread figure as the fairer
 * all-cores bound SURVEY 8(d) asks for.  Sums are accumulated in a thread-dependent order: equal to the one-thread result up to
 * rounding. */
typedef struct {
    const double *cams, *pts, *obs_uv;
    const int32_t *obs_cam, *obs_pt, *cam_free;
    const int* first;
    int np, kmax;
    size_t N;
    double fx, fy, inv_radius;
    double *S, *rhs, *udiag;   /* shared (atomic updates) or private to the thread (summed after the join) */
    int atomic;
    int* next;          /* shared chunk counter */
    double cost;        /* per thread */
} mt_job;

static void* mt_worker(void* arg) {
    mt_job* j = (mt_job*)arg;
    double* Jb = (double*)malloc((size_t)j->kmax * 18 * sizeof(double));
    double* rb = (double*)malloc((size_t)j->kmax * 2 * sizeof(double));
    double cost = 0.0;
    for (;;) {
        const int q0 = __atomic_fetch_add(j->next, 64, __ATOMIC_RELAXED);
        if (q0 >= j->np) break;
        const int q1 = q0 + 64 < j->np ? q0 + 64 : j->np;
        for (int q = q0; q < q1; ++q)
            cost += eliminate_point(j->first[q], j->first[q + 1], j->cams, j->pts, j->obs_uv, j->obs_cam, j->obs_pt, j->cam_free, j->N, j->fx,
                                    j->fy, j->inv_radius, j->S, j->rhs, j->udiag, Jb, rb, j->atomic);
    }
    free(Jb); free(rb);
    j->cost = cost;
    return NULL;
}

double ba_oracle_linearize_mt(int n_threads, int n_cams, int n_pts, int n_obs, const double* cams, const double* pts,
                              const double* obs_uv, const int32_t* obs_cam, const int32_t* obs_pt, const int32_t* cam_free,
                              int n_free, double fx, double fy, double inv_radius, double* S, double* rhs) {
    const size_t N = (size_t)n_free * 6;
    memset(S, 0, N * N * sizeof(double));
    memset(rhs, 0, N * sizeof(double));
    double* udiag = (double*)calloc(N ? N : 1, sizeof(double));
    (void)n_cams;
    /* first observation of every non-empty point */
    int* first = (int*)malloc(((size_t)n_pts + 1) * sizeof(int));
    int np = 0, kmax = 1;
    for (int i = 0; i < n_obs;) {
        int e = i;
        while (e < n_obs && obs_pt[e] == obs_pt[i]) ++e;
        first[np++] = i;
        if (e - i > kmax) kmax = e - i;
        i = e;
    }
    first[np] = n_obs;
    int force_shared = 0;                      /* n_threads < 0: |n_threads| threads on ONE shared S (tests of the atomic path) */
    if (n_threads < 0) { n_threads = -n_threads; force_shared = 1; }
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 256) n_threads = 256;
    /* small systems: a private copy of S / rhs / udiag per thread (no contention on the few hot camera blocks), summed after the
     * join; large systems (configs[4]: 508 MB per copy) share one S with atomic updates — their blocks are rarely hit together */
    const int priv = !force_shared && n_threads > 1 && (double)N * (double)N * 8.0 * n_threads <= 1.5e9;
    int next = 0;
    mt_job jobs[256];
    pthread_t th[256];
    for (int t = 0; t < n_threads; ++t) {
        double *St = S, *rt = rhs, *ut = udiag;
        if (priv && t > 0) {
            St = (double*)calloc(N * N ? N * N : 1, sizeof(double));
            rt = (double*)calloc(N ? N : 1, sizeof(double));
            ut = (double*)calloc(N ? N : 1, sizeof(double));
        }
        mt_job j = {cams, pts, obs_uv, obs_cam, obs_pt, cam_free, first, np, kmax, N, fx, fy, inv_radius, St, rt, ut, !priv, &next, 0.0};
        jobs[t] = j;
        pthread_create(&th[t], NULL, mt_worker, &jobs[t]);
    }
    double cost = 0.0;
    for (int t = 0; t < n_threads; ++t) { pthread_join(th[t], NULL); cost += jobs[t].cost; }
    if (priv)
        for (int t = 1; t < n_threads; ++t) {
            for (size_t d = 0; d < N * N; ++d) S[d] += jobs[t].S[d];
            for (size_t d = 0; d < N; ++d) { rhs[d] += jobs[t].rhs[d]; udiag[d] += jobs[t].udiag[d]; }
            free(jobs[t].S); free(jobs[t].rhs); free(jobs[t].udiag);
        }
    for (size_t d = 0; d < N; ++d) S[d * N + d] += fmax(udiag[d], 1e-6) * inv_radius;
    free(udiag); free(first);
    return cost;
}
