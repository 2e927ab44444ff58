// MonocularSfM::Database — the I/O boundary of the matching path (reference: include/Database/Database.h:17-134,
// src/Database/Database.cpp).  Same public names and the same on-disk format (schema of Database.cpp:710-764,
// blobs = rows x cols little-endian: descriptors float32, keypoints float32 x 4(+), matches int32 x 2 stored with
// image_id1 < image_id2 orientation, pair_id = 10000 * min + max), written from scratch against the public SQLite C
// API, which is resolved with dlopen("libsqlite3.so.0") so no SQLite headers or sources are needed.
#ifndef MSFM_HOST_DATABASE_H_
#define MSFM_HOST_DATABASE_H_
#include <string>
#include <utility>
#include <vector>

#include "Common/Types.h"
#include "cvlite/cvlite.h"

namespace MonocularSfM {

class Database {
public:
    struct Image {
        image_t id;
        std::string name;
    };
    const static int kSchemaVersion = 1;

    Database();
    ~Database();
    // ---- life cycle.  Open creates the file and the five tables when missing (schema of Database.cpp:710-764) and sets
    //      the reference's pragmas (synchronous OFF, WAL, :299-302).  Errors are fatal like in the reference (:8-22).
    void Open(const std::string& path);
    void Close();
    void BeginTransaction() const;        // MatchImagePairs wraps a batch of pairs in one transaction (FeatureMatching.cpp:13,72)
    void EndTransaction() const;

    // ---- images table: (image_id AUTOINCREMENT starting at 1, name UNIQUE)
    image_t WriteImage(const Image& image, const bool use_image_id = false) const;
    bool ExistImageById(const image_t image_id) const;
    bool ExistImageByName(const std::string name) const;
    Image ReadImageById(const image_t image_id) const;          // id == INVALID when absent
    Image ReadImageByName(const std::string name) const;
    std::vector<Image> ReadAllImages() const;                   // ascending image_id
    size_t NumImages() const;

    // ---- per-image blobs (rows, cols, little-endian data).  keypoints: rows x 4 float32 (x, y, size, angle);
    //      colors: rows x 3 uint8 (:143-169); descriptors: rows x 128 float32 (:174-199) — what msfm_desc_upload_f32 takes
    void WriteKeyPoints(const image_t image_id, const std::vector<cv::KeyPoint>& keypoints) const;
    void WriteKeyPointsColor(const image_t image_id, const std::vector<cv::Vec3b>& keypoints) const;
    void WriteDescriptors(const image_t image_id, const cv::Mat& descriptors) const;   // CV_32F only (:176)
    bool ExistKeyPoints(const image_t image_id) const;
    bool ExistKeyPointsColor(const image_t image_id) const;
    bool ExistDescriptors(const image_t image_id) const;
    size_t NumKeyPoints(const image_t image_id) const;
    size_t NumKeyPointsColor(const image_t image_id) const;
    size_t NumDescriptors(const image_t image_id) const;
    std::vector<cv::KeyPoint> ReadKeyPoints(const image_t image_id) const;
    std::vector<cv::Vec3b> ReadKeyPointsColor(const image_t image_id) const;
    cv::Mat ReadDescriptors(const image_t image_id) const;      // a fresh CV_32F copy per call (the matcher caches on the device)

    // ---- matches table: rows x 2 int32 (queryIdx, trainIdx) stored under pair_id = 10000 * min(id) + max(id) in the
    //      (min id, max id) orientation — columns are swapped on the way in and out when image_id1 > image_id2
    //      (:93-99, 541-544, 637-640).  A row exists for every processed pair, also with 0 matches: that is the resume
    //      mechanism of MatchImagePairs (:23-27, 68-70).  DMatch::distance is not persisted.
    void WriteMatches(const image_t image_id1, const image_t image_id2, const std::vector<cv::DMatch>& matches) const;
    bool ExistMatches(const image_pair_t pair_id) const;
    bool ExistMatches(const image_t image_id1, const image_t image_id2) const;
    size_t NumMatches(const image_pair_t pair_id) const;
    size_t NumMatches(const image_t image_id1, const image_t image_id2) const;
    std::vector<cv::DMatch> ReadMatches(const image_pair_t pair_id) const;
    std::vector<cv::DMatch> ReadMatches(const image_t image_id1, const image_t image_id2) const;
    std::vector<std::pair<image_pair_t, std::vector<cv::DMatch>>> ReadAllMatches() const;   // what SceneGraph::Load ingests

    static image_pair_t ImagePairToPairId(const image_t image_id1, const image_t image_id2);
    static void PairIdToImagePair(const image_pair_t pair_id, image_t* image_id1, image_t* image_id2);
    static bool SwapImagePair(const image_t image_id1, const image_t image_id2);

private:
    struct Impl;
    Impl* impl_;
};

}  // namespace MonocularSfM
#endif
