/*
 * msfm_b200.h — C-ABI of the B200-native hot path of MonocularSfM.
 *
 * This is the drop-in boundary: plain pointers and sizes, int status codes, no C++ or
 * torch types.  The reference (nebula-beta/MonocularSfM) has no FFI layer of its own — its boundary is
 * the C++ class API — so every entry point below cites the reference symbol whose arithmetic it
 * replaces.  The C++ classes that keep the reference's names (FeatureUtils, FeatureMatcher,
 * CeresBundelOptimizer, BundleData) live in monocularsfm_b200/host/ and call only these functions.
 *
 * Conventions
 *   - return 0 (MSFM_OK) on success, a negative MSFM_E_* code otherwise; msfm_last_error() gives text.
 *   - one msfm_ctx per (process, device).  A ctx is not thread-safe (the reference is single-threaded,
 *     Database.cpp:297 opens SQLite NOMUTEX); different ctx objects are independent.
 *   - "host" pointers are ordinary (pageable or pinned) CPU memory; "_dev" entry points take CUDA
 *     device pointers valid on the ctx's device.  Nothing allocated by the library crosses the boundary
 *     except the opaque ctx.
 *   - there is NO CPU fallback: without a CUDA device msfm_init fails with MSFM_E_NO_DEVICE.
 */
#ifndef MSFM_B200_H_
#define MSFM_B200_H_

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MSFM_OK             0
#define MSFM_E_INVALID     -1   /* bad argument */
#define MSFM_E_NO_DEVICE   -2   /* no CUDA device / wrong architecture (needs sm_100) */
#define MSFM_E_CUDA        -3   /* a CUDA call or kernel failed; see msfm_last_error */
#define MSFM_E_CAPACITY    -4   /* caller's output buffer too small; needed size reported */
#define MSFM_E_NOT_FOUND   -5   /* unknown image id / handle */
#define MSFM_E_NUMERIC     -6   /* BA: linear solve failed (non-SPD reduced system) */

#define MSFM_DESC_DIM 128       /* SIFT descriptor length (reference: cv::SIFT, FeatureUtils.cpp:27) */

typedef struct msfm_ctx msfm_ctx;

/* ---------------------------------------------------------------------------------------------
 * Context
 * ------------------------------------------------------------------------------------------- */
int  msfm_init(msfm_ctx** ctx, int device_id);
void msfm_destroy(msfm_ctx* ctx);
/* ctx may be NULL: returns the last error of a failed msfm_init in this thread. */
const char* msfm_last_error(const msfm_ctx* ctx);
/* "major.minor.patch" of this library */
const char* msfm_version(void);
/* Blocks until all work queued on the ctx's stream has finished. */
int  msfm_sync(msfm_ctx* ctx);
/* The cudaStream_t (as void*) the ctx launches on — for callers that time with CUDA events. */
void* msfm_stream(msfm_ctx* ctx);
/* Kernels the library has launched since the ctx was created (bench.py's gpu_launches): a counter incremented AT every launch
 * site of the library's own kernels (csrc/launch_count.hpp), process-wide — one ctx per process is the intended use.  Library
 * routines of the fallback solvers (cuSOLVER / cuBLAS) are not counted. */
int64_t msfm_launch_count(const msfm_ctx* ctx);

/* Optional per-kernel-class device timing (CUDA events on the ctx stream, recorded around every launch
 * while enabled).  msfm_prof_read synchronises the stream, then returns the accumulated milliseconds and
 * launch counts per class since the last msfm_prof_reset.  Replaces nothing in the reference (its only
 * instrumentation is Common/Timer.cpp wall clocks, FeatureMatching.cpp:16-17,65). */
#define MSFM_PROF_DESC_FORMAT  0   /* descriptor upload formatting */
#define MSFM_PROF_BUILD_UNITS  1
#define MSFM_PROF_MATCH_TILE   2   /* K1: tcgen05 distance tiles + fused top-2 */
#define MSFM_PROF_RESOLVE      3   /* ratio test + in-group rescan */
#define MSFM_PROF_EXACT        4   /* exact slow-path rows */
#define MSFM_PROF_COMPACT      5   /* cross-check + compaction (3 launches per batch, timed together) */
#define MSFM_PROF_BA_EVAL      6   /* K2: residual + Jacobian + normal-equation blocks */
#define MSFM_PROF_BA_SCHUR     7   /* K2: Schur reduction onto the camera system */
#define MSFM_PROF_BA_OTHER     8
#define MSFM_PROF_BA_COMM      9   /* the all-reduce of the reduced camera system (multi-GPU) */
#define MSFM_PROF_VERIFY      10   /* batched F-matrix RANSAC */
#define MSFM_PROF_BA_SOLVE    11   /* reduced camera system: expansion + Cholesky + triangular solves */
#define MSFM_PROF_NCAT        12
int msfm_prof_enable(msfm_ctx* ctx, int on);
int msfm_prof_reset(msfm_ctx* ctx);
int msfm_prof_read(msfm_ctx* ctx, double ms[MSFM_PROF_NCAT], int64_t launches[MSFM_PROF_NCAT]);

/* ---------------------------------------------------------------------------------------------
 * M-path: brute-force 2-NN descriptor matching
 *
 * Replaces the arithmetic behind
 *   FeatureUtils::ComputeMatches        src/Feature/FeatureUtils.cpp:141-157  (OpenCV knnMatch k=2 + ratio test)
 *   FeatureUtils::ComputeCrossMatches   src/Feature/FeatureUtils.cpp:160-174
 *   FeatureUtils::CrossCheck            src/Feature/FeatureUtils.cpp:281-310
 *   FeatureUtils::FilterMatchesByDistance src/Feature/FeatureUtils.cpp:208-218
 * as driven by FeatureMatcher::MatchImagePairs, src/Feature/FeatureMatching.cpp:10-73.
 *
 * Descriptors are n x 128 uint8, row-major (the BASELINE.json contract; the reference's CV_32F
 * blobs are bridged in the C++ shim).  Results are bit-exact with OpenCV's
 * BFMatcher(NORM_L2) on the same uint8 inputs: distances are sqrtf of the exact integer squared
 * distance, neighbours are ordered by (float distance, index), the ratio test is
 * `d0 < (float)ratio * d1` in float.
 * ------------------------------------------------------------------------------------------- */

/* Upload (or replace) the descriptor set of one image and keep it resident on the device.
 * Replaces the per-pair Database::ReadDescriptors of FeatureMatching.cpp:32-33 (TODO "cache" at :31).
 * n may be 0.  image_id >= 0 (image_t, Common/Types.h:9).  The copy is queued on the ctx stream: a PINNED host buffer
 * must stay valid until the next msfm_sync / matching call returns (pageable memory is staged immediately). */
int msfm_desc_upload_u8(msfm_ctx* ctx, int32_t image_id, const uint8_t* desc_host, int32_t n);
/* Same for the float32 descriptors the reference's database holds (Database::ReadDescriptors, src/Database/Database.cpp:
 * 510-523; blob = rows x 128 little-endian float32, :174-199): the set is copied as float and converted to uint8 on the
 * device.  mode 0: a set whose values are all integers in [0,255] (un-normalised SIFT) converts exactly, any other set
 * is quantised as clamp(rint(512 v), 0, 255), round-half-to-even — the bridge INTEGRATION.md documents for the
 * L1-root / L2 normalised descriptors of FeatureExtraction.cpp:260-281; mode 1: always quantise. */
int msfm_desc_upload_f32(msfm_ctx* ctx, int32_t image_id, const float* desc_host, int32_t n, int32_t mode);
/* RAW SIFT rows (what cv::SIFT::compute returns, FeatureUtils.cpp:27-66) of one image: the extraction-time normalisation
 * of the reference (FeatureExtraction.cpp:143-160: L1RootNormalized :260-271 = row /= |row|_1 then sqrt, or L2Normalized
 * :272-281 = row /= |row|_2, with OpenCV's arithmetic) runs on the device, followed by the x512 quantisation of the
 * bridge above; the set becomes resident.  normalized_out (optional, HOST, [n][128] float32) receives the normalised rows —
 * bit for bit what FeatureExtraction hands to Database::WriteDescriptors (:157) — so that extraction can store them and
 * matching needs no second upload. */
#define MSFM_NORM_L1_ROOT 1
#define MSFM_NORM_L2      2
int msfm_desc_upload_raw_f32(msfm_ctx* ctx, int32_t image_id, const float* desc_host, int32_t n, int32_t normalization,
                             float* normalized_out);
/* 1 if the resident set of image_id came from a float32 upload that had to be quantised (distances are then on the
 * x512 scale: FeatureUtils::FilterMatchesByDistance thresholds must be scaled), 0 otherwise
 * (uint8 upload or exactly converted floats); negative error code.  Synchronises the stream. */
int msfm_desc_quantised(msfm_ctx* ctx, int32_t image_id);
/* Same, source already in device memory (synthetic-data benchmarks; inputs resident in HBM). */
int msfm_desc_upload_u8_dev(msfm_ctx* ctx, int32_t image_id, const uint8_t* desc_dev, int32_t n);
/* Number of descriptors of a resident image, or MSFM_E_NOT_FOUND. */
int msfm_desc_count(msfm_ctx* ctx, int32_t image_id);
int msfm_desc_release(msfm_ctx* ctx, int32_t image_id);
int msfm_desc_release_all(msfm_ctx* ctx);

typedef struct {
    double  max_distance;    /* FilterMatchesByDistance threshold (a double in the reference, FeatureMatching.h:30) on the
                                u8 scale; < 0 disables (FeatureMatching.cpp:49) */
    float   distance_ratio;  /* FeatureMatcher::distance_ratio_, narrowed to float as FeatureUtils.h:95 does (default 0.8f) */
    int32_t cross_check;     /* FeatureMatcher::cross_check_ (default 1) */
    int32_t opencv_quirks;   /* 1: reproduce CrossCheck's unordered_map default-0 behaviour (FeatureUtils.cpp:302) */
    int32_t reserved;        /* must be 0 */
} msfm_match_options;

/* Match P image pairs.  pairs[p] = (image_id1, image_id2): image 1 is the query side (queryIdx),
 * image 2 the train side (trainIdx), exactly as ComputeCrossMatches(desc1, desc2) at
 * FeatureMatching.cpp:39.  Output is CSR: pair p's matches are rows
 * out_offsets[p] .. out_offsets[p+1]-1 of out_matches, each (queryIdx, trainIdx), ascending queryIdx
 * (the order of FeatureUtils.cpp:150-156).  out_dist (optional, may be NULL) receives the float
 * DMatch::distance of each match.  capacity = number of rows out_matches/out_dist can hold; when too
 * small the call returns MSFM_E_CAPACITY and *total_out holds the needed row count (offsets are
 * still filled).  All output pointers are HOST memory. */
int msfm_match_pairs(msfm_ctx* ctx, const int32_t* pairs /*[P][2]*/, int32_t P,
                     const msfm_match_options* opt,
                     int64_t* out_offsets /*[P+1]*/, int32_t* out_matches /*[capacity][2]*/,
                     float* out_dist /*[capacity] or NULL*/, int64_t capacity, int64_t* total_out);

/* Same computation, outputs left in DEVICE memory (no D2H; kernel-only timing, multi-GPU sharding).
 * pairs_dev may be NULL when pairs_host is given. */
int msfm_match_pairs_dev(msfm_ctx* ctx, const int32_t* pairs_host /*[P][2]*/, int32_t P,
                         const msfm_match_options* opt,
                         int64_t* out_offsets_dev /*[P+1]*/, int32_t* out_matches_dev /*[capacity][2]*/,
                         float* out_dist_dev /*[capacity] or NULL*/, int64_t capacity, int64_t* total_out);

/* Raw 2-NN of every row of A in B (host pointers) — the knnMatch(desc1, desc2, k=2) of
 * FeatureUtils.cpp:149, exposed for parity tests.
 *   mode 0: tensor-core path (tcgen05 distance tiles + fused top-2).  idx[i][1] is -1 unless the row
 *           took the exact slow path: the production path only needs the runner-up's DISTANCE.
 *   mode 1: exact CUDA-core scan of every row (also yields idx[i][1]).
 * idx [nA][2] int32 (-1 = none), dist [nA][2] float (inf = none), d2 [nA][2] int32 exact squared
 * distances (optional, may be NULL; -1 = none). */
int msfm_match_knn2_u8(msfm_ctx* ctx, const uint8_t* A, int32_t nA, const uint8_t* B, int32_t nB,
                       int32_t mode, int32_t* idx, float* dist, int32_t* d2);

/* Counters of the last msfm_match_pairs / msfm_match_knn2_u8 call: [0] rows processed, [1] rows that needed the
 * in-group rescan, [2] rows that took the exact slow path, [3] tensor-kernel work units.  */
int msfm_match_stats(msfm_ctx* ctx, int64_t stats[4]);

/* ---------------------------------------------------------------------------------------------
 * Geometric verification of the matches (the step MatchImagePairs applies between matching and WriteMatches)
 *
 * Replaces FeatureUtils::FilterMatches (src/Feature/FeatureUtils.cpp:176-206) =
 * cv::findFundamentalMat(aligned_pts1, aligned_pts2, cv::FM_RANSAC, 3.0, 0.99, inlier_mask), called per pair at
 * src/Feature/FeatureMatching.cpp:60, by ONE batched kernel over all pairs of a call (one CTA per pair): RANSAC over
 * minimal 8-point samples (Hartley normalisation, rank-2 constraint), OpenCV's error measure and adaptive iteration count,
 * refits on the consensus set.  Deterministic (counter-based sampler).  NOT bit-compatible with OpenCV's RANSAC (different
 * sampler and minimal solver): the parity criterion is inlier-set agreement, tests/test_verify_gpu.py.
 * ------------------------------------------------------------------------------------------- */
/* Keep the keypoint positions (KeyPoint::pt, x then y, pixels) of an image resident; uploaded once per image instead of
 * Database::ReadKeyPoints per pair (FeatureMatching.cpp:51-52). */
int  msfm_keypoints_upload(msfm_ctx* ctx, int32_t image_id, const float* xy /*[n][2]*/, int32_t n);
int  msfm_keypoints_count(msfm_ctx* ctx, int32_t image_id);
int  msfm_keypoints_release_all(msfm_ctx* ctx);
typedef struct {
    double  threshold;    /* 3.0 px (FeatureUtils.cpp:196) */
    double  confidence;   /* 0.99 */
    int32_t max_iters;    /* 1000 (OpenCV default) */
    int32_t reserved;     /* must be 0 */
} msfm_verify_options;
void msfm_verify_default_options(msfm_verify_options* opt);
/* Inlier masks for the matches of P pairs.  pairs / offsets / matches exactly as msfm_match_pairs takes and returns them
 * (matches[k] = (queryIdx into image 1's keypoints, trainIdx into image 2's)).  inlier_mask[k] = 1 for the matches
 * FilterMatches would keep (all 0 for a pair with fewer than 8 matches or without a model); inlier_counts [P] optional.
 * HOST pointers. */
int  msfm_verify_pairs(msfm_ctx* ctx, const int32_t* pairs /*[P][2]*/, int32_t P, const int64_t* offsets /*[P+1]*/,
                       const int32_t* matches /*[offsets[P]][2]*/, const msfm_verify_options* opt,
                       uint8_t* inlier_mask /*[offsets[P]]*/, int32_t* inlier_counts /*[P] or NULL*/);
/* Same with DEVICE match lists / outputs (chained behind msfm_match_pairs_dev without a host round trip); pairs on the host. */
int  msfm_verify_pairs_dev(msfm_ctx* ctx, const int32_t* pairs_host /*[P][2]*/, int32_t P, const int64_t* offsets_dev,
                           const int32_t* matches_dev, const msfm_verify_options* opt, uint8_t* inlier_mask_dev,
                           int32_t* inlier_counts_dev /*or NULL*/);

/* ---------------------------------------------------------------------------------------------
 * B-path: bundle-adjustment inner loop
 *
 * Replaces what CeresBundelOptimizer::Optimize (src/Optimizer/CeresBundleOptimizer.cpp:188-328) delegates to
 * Ceres: residual + Jacobian of BundleAutoDiffConstantFocalCostFunction (:29-53, autodiff <2,3,3,3> at :65) for
 * every Measurement of every Landmark (:213-247), the Schur elimination of the points onto the non-constant
 * cameras (solver choice :264-273) and the Levenberg-Marquardt loop with the option block of :262-291.
 * The problem is the flattened BundleData (include/Optimizer/BundleData.h:19-65); the C++ shim
 * monocularsfm_b200/host flattens/unflattens it.
 * ------------------------------------------------------------------------------------------- */
typedef struct msfm_ba msfm_ba;

#define MSFM_BA_REFINE_FOCAL 1      /* msfm_ba_problem.flags: CeresBundelOptimizer::Parameters::refine_focal_length (:20) */

typedef struct {
    int32_t n_cams, n_pts, n_obs;
    int32_t flags;                /* 0, or MSFM_BA_REFINE_FOCAL: (fx, fy) become ONE shared 2-parameter block of every residual
                                     (BundleAutoDiffCostFunction :76-121, AddResidualBlock(..., focal) :227-233) */
    double  fx, fy;               /* K(0,0), K(1,1) (:197-198); the initial focal block with MSFM_BA_REFINE_FOCAL */
    const double*  cams;          /* [n_cams][6]  rvec | tvec  (BundleData::CameraPose, cv::Mat 3x1 f64 each) */
    const double*  pts;           /* [n_pts][3]   Landmark::point3D */
    const double*  obs_uv;        /* [n_obs][2]   Measurement::point2D already centred: (x - cx, y - cy) as :221-222 */
    const int32_t* obs_cam;       /* [n_obs]      index into cams */
    const int32_t* obs_pt;        /* [n_obs]      index into pts, NON-DECREASING (observations grouped by landmark) */
    const uint8_t* cam_const;     /* [n_cams]     1 = BundleData::constant_camera_pose (:256-260) */
} msfm_ba_problem;

typedef struct {
    int32_t max_num_iterations;           /* 100 (:276), doubled below 10 cameras (:289) */
    int32_t verbose;                      /* 1: one line per iteration on stdout */
    double  function_tolerance;           /* Ceres default 1e-6, /10 below 10 cameras (:285) */
    double  gradient_tolerance;           /* Ceres default 1e-10, /10 below 10 cameras (:286) */
    double  parameter_tolerance;          /* Ceres default 1e-8, /10 below 10 cameras (:287) */
    double  initial_trust_region_radius;  /* Ceres default 1e4 */
} msfm_ba_options;

#define MSFM_BA_CONVERGENCE     0   /* ceres::CONVERGENCE: Optimize returns true (:296) */
#define MSFM_BA_NO_CONVERGENCE  1   /* iteration limit: the reference prints "Bundle Adjustment failed." and returns false */
#define MSFM_BA_FAILURE         2

typedef struct {
    int32_t iterations;          /* LM iterations (linearisations) */
    int32_t successful_steps;
    int32_t termination;         /* MSFM_BA_* */
    int32_t num_residuals;       /* 2 * n_obs summed over all ranks (summary.num_residuals, :305) */
    double  initial_cost, final_cost;     /* 1/2 sum r^2, as Ceres (:306-307 print sqrt(2 cost / #res)) */
    double  total_time_s, linearize_time_s, solve_time_s;
} msfm_ba_summary;

/* The option block the reference sets for a problem with n_cams camera poses (:262-291). */
void msfm_ba_default_options(msfm_ba_options* opt, int32_t n_cams);

/* Copy a problem to the device (all pointers HOST).  With a communicator attached to the ctx
 * (msfm_comm_init) every rank passes ALL cameras and its own share of the points/observations. */
int  msfm_ba_create(msfm_ctx* ctx, const msfm_ba_problem* prob, msfm_ba** out);
void msfm_ba_destroy(msfm_ba* ba);
/* The next problem on the same object.  MapBuilder calls Optimize again and again on a map that changes a little between
 * the calls (LocalBA after almost every registered image, GlobalBA in rounds: src/Reconstruction/MapBuilder.cpp:189-215,
 * 576-625), and the reference rebuilds the whole ceres::Problem every time (CeresBundleOptimizer.cpp:206-260).  Here the
 * device problem persists:
 *   - same sparsity pattern (n_cams, n_pts, n_obs, flags, cam_const, obs_cam, obs_pt all equal — compared exactly): only
 *     the VALUES travel (cams, pts, obs_uv, fx, fy); device order, tiles, block structure, solver set-up and every
 *     allocation are kept.  *reused = 1.
 *   - anything else: the structure is analysed again INTO the same object; its device arena is re-used when it is large
 *     enough (grown by 25 % otherwise), so a map that grows does not pay cudaMalloc / cudaFree per call.  *reused = 0.
 * With a communicator attached every rank must call it; the ranks agree on the path with one 8-byte all-reduce.
 * After an error other than MSFM_E_INVALID the object can only be destroyed.  reused may be NULL. */
int  msfm_ba_update(msfm_ba* ba, const msfm_ba_problem* prob, int32_t* reused);
/* sizes[0..2] = n_cams, n_pts, n_obs of the resident problem (the array lengths msfm_ba_get_params / _filter_stats expect). */
int  msfm_ba_sizes(msfm_ba* ba, int64_t sizes[3]);
/* Host -> device traffic of the last msfm_ba_create / msfm_ba_update: info[0] bytes, info[1] = 1 if the structure was kept. */
int  msfm_ba_last_upload(msfm_ba* ba, int64_t info[2]);
/* Structure of the problem as the device sees it: info[0] free cameras F, [1] non-empty 6x6 blocks of the upper block
 * triangle of the reduced camera system (the fp32 part of the all-reduce message), [2] point tiles, [3] max cameras per
 * tile, [4] bytes of the system buffer (= the all-reduce message + 16), [5] shared memory per CTA of the linearisation
 * kernel, [6] fp64 words of the message tail (scalars | rhs | g_c | diag U | focal border), [7] tracks longer than 32 views
 * (handled by a pre-pass + item tiles). */
int  msfm_ba_structure(msfm_ba* ba, int32_t info[8]);
/* Current parameters -> host ([n_cams][6], [n_pts][3]; either may be NULL). */
int  msfm_ba_get_params(msfm_ba* ba, double* cams, double* pts);
int  msfm_ba_set_params(msfm_ba* ba, const double* cams, const double* pts);

/* Residuals (fp64) and Jacobians (fp32, [n_obs][2][9], columns rvec|tvec|point — the layout of the
 * autodiff cost function's three parameter blocks) of the local observations; cost = 1/2 sum r^2 (local).
 * r, J may be NULL. */
int  msfm_ba_evaluate(msfm_ba* ba, double* r /*[n_obs][2]*/, float* J /*[n_obs][2][9]*/, double* cost);
/* Mean reprojection error (pixels) of every local point over its observations at the current parameters:
 * err[p] = mean_o sqrt(dx^2 + dy^2) — the per-track error Map::UpdateFromBAData recomputes on the host after
 * every BA (src/Reconstruction/Map.cpp:1201 -> ComputeTrackError :1834-1846 ->
 * Projection::CalculateReprojectionError, src/Reconstruction/Projection.cpp:114-133).  0 for a point without
 * observations. */
int  msfm_ba_track_errors(msfm_ba* ba, double* err /*[n_pts]*/);

/* The statistics Map::FilterAllPoints3D recomputes on the host after every global BA (src/Reconstruction/Map.cpp:793-917),
 * served from the resident problem at the current parameters: obs_keep[o] = HasPositiveDepth && reprojection error <=
 * max_reproj_error (the observations FilterPoints3DWithLargeReprojectionError keeps, :804-870; caller's observation order),
 * pt_mean_error[p] = mean error over the kept observations (what SetError stores, :860), pt_kept[p] their number, and
 * pt_max_parallax_deg[p] = the largest Projection::CalculateParallaxAngle over the point's camera pairs
 * (Projection.cpp:149-194) — FilterPoints3DWithSmallTriangulationAngle keeps a point iff it is >= min_tri_angle (:871-917).
 * Any output may be NULL.  The removals themselves stay with the caller's Map. */
int  msfm_ba_filter_stats(msfm_ba* ba, double max_reproj_error, uint8_t* obs_keep /*[n_obs]*/, double* pt_mean_error /*[n_pts]*/,
                          int32_t* pt_kept /*[n_pts]*/, double* pt_max_parallax_deg /*[n_pts]*/);

/* One linearisation at the current parameters: the reduced camera system of the free cameras with Marquardt
 * damping diag(J^T J) / radius (inv_radius = 1/radius; 0 = undamped), summed over all ranks.
 * S [6F][6F] (symmetric, fully filled), rhs [6F] (S dc = rhs), gc [6F] (gradient of the camera blocks); any may
 * be NULL.  *n_free_out = F. */
int  msfm_ba_linearize(msfm_ba* ba, double inv_radius, double* S, double* rhs, double* gc, double* cost,
                       int32_t* n_free_out);

/* MSFM_BA_REFINE_FOCAL problems: the part of the same linearisation that involves the shared focal block, which borders
 * the camera system:  [S B; B^T F] [dc; df] = [rhs; rhs_f].  B [6F][2], F [3] = (F00, F01, F11) damped like S,
 * rhs_f [2], g_f [2] (gradient of the focal block); any may be NULL.  MSFM_E_INVALID for other problems. */
int  msfm_ba_linearize_focal(msfm_ba* ba, double inv_radius, double* B, double* F, double* rhs_f, double* g_f);
/* Current (fx, fy): what Optimize writes back into K after the solve (:313-317). */
int  msfm_ba_get_focal(msfm_ba* ba, double focal[2]);

/* The whole Levenberg-Marquardt solve; parameters stay on the device (msfm_ba_get_params to read).  One host
 * synchronisation per iteration (a 72-byte record).  The reduced camera system is solved in fp64: after renumbering the cameras
 * for a narrow band, by the library's own cooperative band Cholesky kernel (csrc/ba_band.cu); by a dense Cholesky (cuSOLVER)
 * when the band is wide.  MSFM_BA_SOLVER=band|chain|dense selects explicitly (development / A-B measurements). */
int  msfm_ba_solve(msfm_ba* ba, const msfm_ba_options* opt, msfm_ba_summary* summary);
/* One linearisation at the current parameters and the solution dc [6F] of the damped reduced camera system S dc = rhs by the
 * solver msfm_ba_solve uses (parity hook: tests compare it with a host solve of msfm_ba_linearize's S, rhs).  *status != 0:
 * the system was not positive definite. */
int  msfm_ba_solve_system(msfm_ba* ba, double inv_radius, double* dc /*[6F]*/, int32_t* status);
/* After a solve: info[5] = which linear solver ran: 2 own band Cholesky (info[0] band half-width in cameras, [1] tile order,
 * [2] tile rows, [3] band half-width in tiles), 1 library block-tridiagonal chain (info[0] cameras per super-block, [1] its
 * order, [2] super-blocks), 0 dense Cholesky; info[4] free cameras. */
int  msfm_ba_solver_info(msfm_ba* ba, int32_t info[6]);

/* ---------------------------------------------------------------------------------------------
 * Multi-GPU: one process per GPU; the reduced camera system is summed with ONE ncclAllReduce per
 * linearisation over NVLink.  The unique id is created on rank 0 and distributed by the host program
 * (torch.distributed / MPI / a file) — see INTEGRATION.md.
 * ------------------------------------------------------------------------------------------- */
#define MSFM_COMM_ID_BYTES 128
int  msfm_comm_unique_id(void* id_out /*[128]*/);
int  msfm_comm_init(msfm_ctx* ctx, int32_t n_ranks, int32_t rank, const void* id /*[128]*/);
int  msfm_comm_destroy(msfm_ctx* ctx);
/* sum (op 0) or max (op 1) of a DEVICE fp64 buffer across ranks, in place, on the ctx stream */
int  msfm_comm_allreduce_f64(msfm_ctx* ctx, double* buf_dev, int64_t count, int32_t op);

#ifdef __cplusplus
}
#endif
#endif /* MSFM_B200_H_ */
