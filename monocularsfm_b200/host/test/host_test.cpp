// Test driver of the drop-in C++ classes (run by tests/test_host_cpp.py).
//   host_test cpu <tmp.db>             Database round trip, pair ids, CrossCheck quirk, distance filter  (no GPU)
//   host_test match <db> <preempt 0|1> [noverify] [quiet] [max_pairs_size]  BruteFeatureMatcher(db).RunMatching() on a database made by the
//                                      Python test; prints the time of RunMatching (bench.py: dropin_database)
//   host_test seq <db> [noverify]                  SequentialFeatureMatcher(db).RunMatching()
//   (noverify = the explicit opt-out of the geometric verification, for tests of the exact match lists)
//   host_test two <a.u8> <na> <b.u8> <nb> <out.txt>   FeatureUtils::ComputeMatches / ComputeCrossMatches on raw files
//   host_test ba <in.bin> <out.bin>    CeresBundelOptimizer::Optimize on a BundleData read from a flat binary file
//   host_test ba_persist <in.bin> <out.bin>   three Optimize calls on ONE optimizer (same map, perturbed values, changed map) vs fresh ones
//   host_test ransac <in.bin> <out.bin>  FeatureUtils::FilterMatches on float32 point pairs (no GPU)
//   host_test scenegraph <script.txt> <out.txt>  SceneGraph driven by a script of I / M / F / Q lines (no GPU)
//   host_test scenegraph_db <db> <min_matches> <out.txt>  SceneGraph::Load on a database + the same dump (no GPU)
#include <cmath>
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <chrono>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include "Database/Database.h"
#include "Feature/FeatureMatching.h"
#include "Feature/FeatureUtils.h"
#include "Optimizer/CeresBundleOptimizer.h"
#include "Reconstruction/SceneGraph.h"

using namespace MonocularSfM;

#define REQUIRE(cond)                                                                    \
    do {                                                                                 \
        if (!(cond)) { std::fprintf(stderr, "REQUIRE failed: %s (line %d)\n", #cond, __LINE__); return 1; } \
    } while (0)

static int test_cpu(const std::string& path) {
    std::remove(path.c_str());
    {
        Database db;
        db.Open(path);
        db.BeginTransaction();
        Database::Image im;
        im.name = "a.jpg";
        const image_t id0 = db.WriteImage(im);
        im.name = "b.jpg";
        const image_t id1 = db.WriteImage(im);
        REQUIRE(id0 == 1 && id1 == 2);                       // AUTOINCREMENT starts at 1, like the reference's DBs
        cv::Mat d(3, 128, CV_32F);
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 128; ++j) d.at<float>(i, j) = static_cast<float>((i * 7 + j) % 200);
        db.WriteDescriptors(id0, d);
        std::vector<cv::KeyPoint> kp(3);
        kp[1].pt = cv::Point2f(3.5f, 4.5f);
        kp[1].size = 9.f;
        db.WriteKeyPoints(id0, kp);
        std::vector<cv::DMatch> m;
        m.push_back(cv::DMatch(5, 7, 0, 1.f));
        m.push_back(cv::DMatch(6, 1, 0, 2.f));
        db.WriteMatches(id1, id0, m);                        // id1 > id0: stored swapped
        db.EndTransaction();
        db.Close();
    }
    {
        Database db;
        db.Open(path);
        REQUIRE(db.NumImages() == 2);
        // the rest of the reference's public surface (include/Database/Database.h:37-69)
        REQUIRE(db.ExistImageByName("a.jpg") && !db.ExistImageByName("zzz.jpg"));
        REQUIRE(db.ReadImageById(2).name == "b.jpg" && db.ReadImageByName("a.jpg").id == 1 && db.ReadImageById(77).id == INVALID);
        REQUIRE(db.NumKeyPoints(1) == 3 && db.NumKeyPoints(2) == 0);
        REQUIRE(!db.ExistKeyPointsColor(1) && db.NumKeyPointsColor(1) == 0);
        {
            std::vector<cv::Vec3b> col(3);
            col[1] = cv::Vec3b(10, 20, 30);
            db.WriteKeyPointsColor(1, col);
            REQUIRE(db.ExistKeyPointsColor(1) && db.NumKeyPointsColor(1) == 3);
            const std::vector<cv::Vec3b> back = db.ReadKeyPointsColor(1);
            REQUIRE(back.size() == 3 && back[1][0] == 10 && back[1][1] == 20 && back[1][2] == 30 && back[2][0] == 0);
        }
        REQUIRE(db.NumMatches(Database::ImagePairToPairId(2, 1)) == 2);
        {
            const std::vector<cv::DMatch> pm = db.ReadMatches(Database::ImagePairToPairId(1, 2));   // (image 1, image 2) orientation
            REQUIRE(pm.size() == 2 && pm[0].queryIdx == 7 && pm[0].trainIdx == 5);
        }
        REQUIRE(db.ExistDescriptors(1) && !db.ExistDescriptors(2));
        cv::Mat d = db.ReadDescriptors(1);
        REQUIRE(d.rows == 3 && d.cols == 128 && d.type() == CV_32F && d.at<float>(2, 5) == 19.f);
        REQUIRE(db.ReadKeyPoints(1)[1].size == 9.f && db.ReadKeyPoints(1)[1].pt.y == 4.5f);
        REQUIRE(db.ExistMatches(2, 1) && db.ExistMatches(1, 2) && !db.ExistMatches(1, 3));
        std::vector<cv::DMatch> a = db.ReadMatches(2, 1), b = db.ReadMatches(1, 2);
        REQUIRE(a.size() == 2 && a[0].queryIdx == 5 && a[0].trainIdx == 7);
        REQUIRE(b.size() == 2 && b[0].queryIdx == 7 && b[0].trainIdx == 5);     // read back in the other orientation
        REQUIRE(db.ReadAllMatches().size() == 1 && db.ReadAllMatches()[0].first == 10002);
        db.Close();
    }
    REQUIRE(Database::ImagePairToPairId(3, 12) == 30012 && Database::ImagePairToPairId(12, 3) == 30012);
    image_t a, b;
    Database::PairIdToImagePair(30012, &a, &b);
    REQUIRE(a == 3 && b == 12);
    // CrossCheck: queryIdx 0 survives when its trainIdx has no reverse match (unordered_map default 0)
    std::vector<cv::DMatch> m12, m21, out;
    m12.push_back(cv::DMatch(0, 4, 0, 1.f));
    m12.push_back(cv::DMatch(1, 5, 0, 1.f));
    m12.push_back(cv::DMatch(2, 6, 0, 1.f));
    m21.push_back(cv::DMatch(6, 2, 0, 1.f));
    FeatureUtils::CrossCheck(m12, m21, out);
    REQUIRE(out.size() == 2 && out[0].queryIdx == 0 && out[1].queryIdx == 2);
    std::vector<cv::DMatch> f;
    m12[1].distance = 0.70001f;     // dropped
    m12[0].distance = 0.7f;         // float(0.7) < 0.7 (double): kept, as `distance > max_distance` in the reference
    m12[2].distance = 0.5f;
    FeatureUtils::FilterMatchesByDistance(m12, f, 0.7);
    REQUIRE(f.size() == 2);
    // bridge: integral floats are cast, normalised floats are quantised x512
    cv::Mat fl(1, 128, CV_32F);
    fl.at<float>(0, 0) = 0.25f;
    cv::Mat q = FeatureUtils::ToUint8Descriptors(fl);
    REQUIRE(q.at<unsigned char>(0, 0) == 128);
    std::printf("host cpu tests ok\n");
    return 0;
}

static int run_two(char** argv) {
    const int na = std::atoi(argv[3]), nb = std::atoi(argv[5]);
    cv::Mat a(na, 128, CV_8U), b(nb, 128, CV_8U);
    std::ifstream fa(argv[2], std::ios::binary), fb(argv[4], std::ios::binary);
    fa.read(reinterpret_cast<char*>(a.data), static_cast<std::streamsize>(na) * 128);
    fb.read(reinterpret_cast<char*>(b.data), static_cast<std::streamsize>(nb) * 128);
    std::vector<cv::DMatch> m, mc;
    FeatureUtils::ComputeMatches(a, b, m);                   // defaults: ratio 0.8
    FeatureUtils::ComputeCrossMatches(a, b, mc, 0.8f);
    std::ofstream out(argv[6]);
    out.precision(9);
    out << m.size() << "\n";
    for (const cv::DMatch& d : m) out << d.queryIdx << " " << d.trainIdx << " " << d.distance << "\n";
    out << mc.size() << "\n";
    for (const cv::DMatch& d : mc) out << d.queryIdx << " " << d.trainIdx << " " << d.distance << "\n";
    return 0;
}

static int run_ba(char** argv) {
    // flat file: int32 n_cams n_pts n_obs n_const | f64 fx fy cx cy | cams[n_cams*6] | pts[n_pts*3] | obs_xy[n_obs*2] (pixels,
    // NOT centred) | int32 obs_cam[n_obs] obs_pt[n_obs] const_ids[n_const]
    std::ifstream in(argv[2], std::ios::binary);
    int32_t hdr[4];
    in.read(reinterpret_cast<char*>(hdr), sizeof hdr);
    double k[4];
    in.read(reinterpret_cast<char*>(k), sizeof k);
    std::vector<double> cams(hdr[0] * 6), pts(hdr[1] * 3), xy(hdr[2] * 2);
    std::vector<int32_t> oc(hdr[2]), op(hdr[2]), cst(hdr[3]);
    in.read(reinterpret_cast<char*>(cams.data()), cams.size() * 8);
    in.read(reinterpret_cast<char*>(pts.data()), pts.size() * 8);
    in.read(reinterpret_cast<char*>(xy.data()), xy.size() * 8);
    in.read(reinterpret_cast<char*>(oc.data()), oc.size() * 4);
    in.read(reinterpret_cast<char*>(op.data()), op.size() * 4);
    in.read(reinterpret_cast<char*>(cst.data()), cst.size() * 4);
    BundleData bd;
    bd.K = cv::Mat(3, 3, CV_64F);
    bd.K.at<double>(0, 0) = k[0]; bd.K.at<double>(1, 1) = k[1]; bd.K.at<double>(0, 2) = k[2]; bd.K.at<double>(1, 2) = k[3];
    bd.K.at<double>(2, 2) = 1.0;
    const int id_off = 100;                                  // image ids need not be 0-based
    for (int c = 0; c < hdr[0]; ++c) {
        // every other pose as a 1x3 row: the reference hands `rvec.data` to Ceres as double[3], either shape works there
        const bool row = (c & 1) != 0;
        cv::Mat r(row ? 1 : 3, row ? 3 : 1, CV_64F), t(row ? 1 : 3, row ? 3 : 1, CV_64F);
        for (int j = 0; j < 3; ++j) { reinterpret_cast<double*>(r.data)[j] = cams[6 * c + j]; reinterpret_cast<double*>(t.data)[j] = cams[6 * c + 3 + j]; }
        bd.camera_poses[id_off + c] = BundleData::CameraPose(r, t);
    }
    for (int p = 0; p < hdr[1]; ++p) bd.landmarks[7 * p + 3].point3D = cv::Vec3d(pts[3 * p], pts[3 * p + 1], pts[3 * p + 2]);
    for (int i = 0; i < hdr[2]; ++i)
        bd.landmarks[7 * op[i] + 3].measurements.push_back(BundleData::Measurement(id_off + oc[i], cv::Vec2d(xy[2 * i], xy[2 * i + 1])));
    for (int32_t c : cst) bd.constant_camera_pose.insert(id_off + c);
    const double before = bd.Debug();
    CeresBundelOptimizer::Parameters params;
    CeresBundelOptimizer opt(params);
    const bool ok = opt.Optimize(bd);
    const double after = bd.Debug();
    for (int c = 0; c < hdr[0]; ++c)
        for (int j = 0; j < 3; ++j) {
            cams[6 * c + j] = reinterpret_cast<const double*>(bd.camera_poses[id_off + c].rvec.data)[j];
            cams[6 * c + 3 + j] = reinterpret_cast<const double*>(bd.camera_poses[id_off + c].tvec.data)[j];
        }
    for (int p = 0; p < hdr[1]; ++p)
        for (int j = 0; j < 3; ++j) pts[3 * p + j] = bd.landmarks[7 * p + 3].point3D(j);
    std::ofstream out(argv[3], std::ios::binary);
    const double head[6] = {ok ? 1.0 : 0.0, before, after, opt.last_initial_cost(), opt.last_final_cost(), static_cast<double>(opt.last_iterations())};
    out.write(reinterpret_cast<const char*>(head), sizeof head);
    out.write(reinterpret_cast<const char*>(cams.data()), cams.size() * 8);
    out.write(reinterpret_cast<const char*>(pts.data()), pts.size() * 8);
    // shared-focal variant (refine_focal_length, CeresBundleOptimizer.cpp:225-233, 313-317): start from a wrong focal
    // length on the already optimised scene; the solve has to pull K(0,0), K(1,1) back and lower the error
    CeresBundelOptimizer::Parameters p2;
    p2.refine_focal_length = true;
    CeresBundelOptimizer opt2(p2);
    const double fx_true = bd.K.at<double>(0, 0), fy_true = bd.K.at<double>(1, 1);
    bd.K.at<double>(0, 0) = fx_true * 1.02;
    bd.K.at<double>(1, 1) = fy_true * 0.98;
    const double before2 = bd.Debug();
    if (!opt2.Optimize(bd)) return 3;
    const double after2 = bd.Debug();
    const double tail[4] = {before2, after2, bd.K.at<double>(0, 0) / fx_true, bd.K.at<double>(1, 1) / fy_true};
    out.write(reinterpret_cast<const char*>(tail), sizeof tail);
    return 0;
}

// The optimizer object persists across Optimize calls like MapBuilder's bundle_optimizer_ (MapBuilder.cpp:92,582,618):
//   call 1  the scene as read;  call 2  the same map with perturbed values (structure kept: only values travel);
//   call 3  a changed map (every 7th landmark removed, one camera released from the constant set): analysed again into the same
//   device object.  Every call is repeated by a FRESH optimizer on a copy of the same input; the costs must agree.
// out: f64 [3][6] = reused, h2d bytes, final cost, fresh final cost, fresh h2d bytes, iterations
static int run_ba_persist(char** argv) {
    std::ifstream in(argv[2], std::ios::binary);
    int32_t hdr[4];
    in.read(reinterpret_cast<char*>(hdr), sizeof hdr);
    double k[4];
    in.read(reinterpret_cast<char*>(k), sizeof k);
    std::vector<double> cams(hdr[0] * 6), pts(hdr[1] * 3), xy(hdr[2] * 2);
    std::vector<int32_t> oc(hdr[2]), op(hdr[2]), cst(hdr[3]);
    in.read(reinterpret_cast<char*>(cams.data()), cams.size() * 8);
    in.read(reinterpret_cast<char*>(pts.data()), pts.size() * 8);
    in.read(reinterpret_cast<char*>(xy.data()), xy.size() * 8);
    in.read(reinterpret_cast<char*>(oc.data()), oc.size() * 4);
    in.read(reinterpret_cast<char*>(op.data()), op.size() * 4);
    in.read(reinterpret_cast<char*>(cst.data()), cst.size() * 4);
    auto make = [&](double shift, bool changed) {
        BundleData bd;
        bd.K = cv::Mat(3, 3, CV_64F);
        bd.K.at<double>(0, 0) = k[0]; bd.K.at<double>(1, 1) = k[1]; bd.K.at<double>(0, 2) = k[2]; bd.K.at<double>(1, 2) = k[3];
        bd.K.at<double>(2, 2) = 1.0;
        for (int c = 0; c < hdr[0]; ++c) {
            cv::Mat r(3, 1, CV_64F), t(3, 1, CV_64F);
            for (int j = 0; j < 3; ++j) { r.at<double>(j, 0) = cams[6 * c + j]; t.at<double>(j, 0) = cams[6 * c + 3 + j] + shift * ((c + j) % 3 - 1); }
            bd.camera_poses[c] = BundleData::CameraPose(r, t);
        }
        for (int p = 0; p < hdr[1]; ++p) {
            if (changed && p % 7 == 0) continue;
            bd.landmarks[p].point3D = cv::Vec3d(pts[3 * p] + shift, pts[3 * p + 1] - shift, pts[3 * p + 2] + 0.5 * shift);
        }
        for (int i = 0; i < hdr[2]; ++i) {
            if (changed && op[i] % 7 == 0) continue;
            bd.landmarks[op[i]].measurements.push_back(BundleData::Measurement(oc[i], cv::Vec2d(xy[2 * i], xy[2 * i + 1])));
        }
        for (size_t i = 0; i < cst.size(); ++i)
            if (!changed || i + 1 < cst.size() || cst.size() == 1) bd.constant_camera_pose.insert(cst[i]);
        return bd;
    };
    CeresBundelOptimizer::Parameters params;
    CeresBundelOptimizer persistent(params);
    std::ofstream out(argv[3], std::ios::binary);
    const double shifts[3] = {0.0, 0.01, 0.005};
    for (int call = 0; call < 3; ++call) {
        BundleData a = make(shifts[call], call == 2), b = make(shifts[call], call == 2);
        if (!persistent.Optimize(a)) return 4;
        CeresBundelOptimizer fresh(params);
        if (!fresh.Optimize(b)) return 5;
        const double rec[6] = {persistent.last_structure_reused() ? 1.0 : 0.0, static_cast<double>(persistent.last_h2d_bytes()), persistent.last_final_cost(),
                               fresh.last_final_cost(), static_cast<double>(fresh.last_h2d_bytes()), static_cast<double>(persistent.last_iterations())};
        out.write(reinterpret_cast<const char*>(rec), sizeof rec);
        // both optimizers must have written the same parameters back
        double worst = 0;
        for (auto& el : a.landmarks)
            for (int j = 0; j < 3; ++j) worst = std::max(worst, std::fabs(el.second.point3D(j) - b.landmarks[el.first].point3D(j)));
        out.write(reinterpret_cast<const char*>(&worst), sizeof worst);
        if (call == 2) {
            // Map::FilterAllPoints3D's statistics from the resident problem
            std::vector<unsigned char> keep;
            std::vector<double> err, ang;
            std::vector<int> kept;
            if (!persistent.FilterStatistics(4.0, &keep, &err, &kept, &ang)) return 6;
            double s[4] = {static_cast<double>(keep.size()), static_cast<double>(err.size()), 0, 0};
            for (unsigned char v : keep) s[2] += v;
            for (double v : err) s[3] += v;
            out.write(reinterpret_cast<const char*>(s), sizeof s);
        }
    }
    return 0;
}

// in: int32 n | float32 [n][2] pts1 | float32 [n][2] pts2 ; matches are (i, i).  out: uint8 [n] kept-mask
static int run_ransac(char** argv) {
    std::ifstream in(argv[2], std::ios::binary);
    int32_t n = 0;
    in.read(reinterpret_cast<char*>(&n), 4);
    std::vector<float> a(2 * n), b(2 * n);
    in.read(reinterpret_cast<char*>(a.data()), a.size() * 4);
    in.read(reinterpret_cast<char*>(b.data()), b.size() * 4);
    std::vector<cv::Point2f> p1(n), p2(n);
    std::vector<cv::DMatch> matches;
    for (int i = 0; i < n; ++i) {
        p1[i] = cv::Point2f(a[2 * i], a[2 * i + 1]);
        p2[n - 1 - i] = cv::Point2f(b[2 * i], b[2 * i + 1]);      // train points stored reversed: exercises the index alignment
        matches.push_back(cv::DMatch(i, n - 1 - i, 0, 1.f));
    }
    std::vector<cv::DMatch> kept;
    FeatureUtils::FilterMatches(p1, p2, matches, kept);
    std::vector<unsigned char> mask(n, 0);
    for (const cv::DMatch& m : kept) {
        if (m.trainIdx != n - 1 - m.queryIdx) return 4;
        mask[m.queryIdx] = 1;
    }
    std::ofstream out(argv[3], std::ios::binary);
    out.write(reinterpret_cast<const char*>(mask.data()), mask.size());
    return 0;
}

// full, order-independent dump of a scene graph: per image (ascending id) its counters and every point's list
static void dump_scene_graph(SceneGraph& g, std::ostream& out) {
    std::vector<image_t> ids = g.GetAllImageIds();
    std::sort(ids.begin(), ids.end());
    out << "images " << g.NumImages() << "\n";
    for (image_t id : ids) {
        out << "image " << id << " obs " << g.NumObservationsForImage(id) << " corrs " << g.NumCorrespondencesForImage(id) << "\n";
    }
    std::vector<std::pair<image_pair_t, point2D_t>> pairs;
    for (const auto& kv : g.ImagePairs()) pairs.push_back(kv);
    std::sort(pairs.begin(), pairs.end());
    for (const auto& kv : pairs) out << "pair " << kv.first << " " << kv.second << "\n";
}

static int run_scenegraph(char** argv) {
    std::ifstream in(argv[2]);
    std::ofstream out(argv[3]);
    SceneGraph g;
    std::string op;
    while (in >> op) {
        if (op == "I") { int id, n; in >> id >> n; g.AddImage(id, n); }
        else if (op == "M") {
            int a, b, k; in >> a >> b >> k;
            std::vector<cv::DMatch> m;
            for (int i = 0; i < k; ++i) { int q, t; in >> q >> t; m.push_back(cv::DMatch(q, t, 0, 0.f)); }
            g.AddCorrespondences(a, b, m);
        }
        else if (op == "F") { g.Finalize(); }
        else if (op == "D") { dump_scene_graph(g, out); }
        else if (op == "Q") {          // Q image point: correspondences, has, two-view
            int id, p; in >> id >> p;
            out << "q " << id << " " << p << " has " << (g.HasCorrespondences(id, p) ? 1 : 0) << " two " << (g.IsTwoViewObservation(id, p) ? 1 : 0) << " :";
            for (const SceneGraph::Correspondence& c : g.FindCorrespondences(id, p)) out << " " << c.image_id << "," << c.point2D_idx;
            out << "\n";
        }
        else if (op == "B") {          // B id1 id2: matches between two images + the pair counter
            int a, b; in >> a >> b;
            out << "b " << a << " " << b << " n " << g.NumCorrespondencesBetweenImages(a, b) << " :";
            for (const cv::DMatch& m : g.FindCorrespondencesBetweenImages(a, b)) out << " " << m.queryIdx << "," << m.trainIdx;
            out << "\n";
        }
        else return 5;
    }
    return 0;
}

static int run_scenegraph_db(char** argv) {
    cv::Ptr<Database> db(new Database());
    db->Open(argv[2]);
    SceneGraph g;
    g.Load(db, static_cast<size_t>(std::atoi(argv[3])));
    std::ofstream out(argv[4]);
    dump_scene_graph(g, out);
    db->Close();
    return 0;
}

int main(int argc, char** argv) {
    if (argc < 2) return 2;
    const std::string mode = argv[1];
    if (mode == "cpu" && argc >= 3) return test_cpu(argv[2]);
    if (mode == "match" && argc >= 4) {
        // optional arguments in any order: "noverify", "quiet", a number = max_pairs_size (default 100 like the reference)
        int max_pairs = 100;
        bool noverify = false, quiet = false;
        for (int a = 4; a < argc; ++a) {
            const std::string arg = argv[a];
            if (arg == "noverify") noverify = true;
            else if (arg == "quiet") quiet = true;
            else if (std::atoi(argv[a]) > 0) max_pairs = std::atoi(argv[a]);
        }
        BruteFeatureMatcher matcher(argv[2], max_pairs, std::atoi(argv[3]) != 0);
        if (noverify) matcher.SetGeometricFilter(FeatureMatcher::GeometricFilter());
        if (quiet) matcher.SetVerbose(false);
        const auto t0 = std::chrono::steady_clock::now();
        matcher.RunMatching();
        std::printf("RunMatching: %.6f s\n", std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
        return 0;
    }
    if (mode == "seq" && argc >= 3) {
        SequentialFeatureMatcher matcher(argv[2]);
        if (argc >= 4 && std::string(argv[3]) == "noverify") matcher.SetGeometricFilter(FeatureMatcher::GeometricFilter());
        matcher.RunMatching();
        return 0;
    }
    if (mode == "two" && argc >= 7) return run_two(argv);
    if (mode == "ba" && argc >= 4) return run_ba(argv);
    if (mode == "ba_persist" && argc >= 4) return run_ba_persist(argv);
    if (mode == "ransac" && argc >= 4) return run_ransac(argv);
    if (mode == "scenegraph" && argc >= 4) return run_scenegraph(argv);
    if (mode == "scenegraph_db" && argc >= 5) return run_scenegraph_db(argv);
    return 2;
}
