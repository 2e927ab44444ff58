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
    void Open(const std::string& path);
    void Close();
    void BeginTransaction() const;
    void EndTransaction() const;

    bool ExistImageById(const image_t image_id) const;
    bool ExistImageByName(const std::string name) const;
    bool ExistKeyPoints(const image_t image_id) const;
    bool ExistKeyPointsColor(const image_t image_id) const;
    bool ExistDescriptors(const image_t image_id) const;
    bool ExistMatches(const image_pair_t pair_id) const;
    bool ExistMatches(const image_t image_id1, const image_t image_id2) const;

    size_t NumImages() const;
    size_t NumKeyPoints(const image_t image_id) const;
    size_t NumKeyPointsColor(const image_t image_id) const;
    size_t NumDescriptors(const image_t image_id) const;
    size_t NumMatches(const image_pair_t pair_id) const;
    size_t NumMatches(const image_t image_id1, const image_t image_id2) const;

    Image ReadImageById(const image_t image_id) const;                           // default Image when absent (Database.cpp:437-452)
    Image ReadImageByName(const std::string name) const;
    std::vector<Image> ReadAllImages() const;
    std::vector<cv::KeyPoint> ReadKeyPoints(const image_t image_id) const;
    std::vector<cv::Vec3b> ReadKeyPointsColor(const image_t image_id) const;     // colors blob: rows x 3 uint8 (:143-169)
    cv::Mat ReadDescriptors(const image_t image_id) const;                       // CV_32F rows x cols
    std::vector<cv::DMatch> ReadMatches(const image_pair_t pair_id) const;       // oriented as (min id, max id)
    std::vector<cv::DMatch> ReadMatches(const image_t image_id1, const image_t image_id2) const;
    std::vector<std::pair<image_pair_t, std::vector<cv::DMatch>>> ReadAllMatches() const;

    image_t WriteImage(const Image& image, const bool use_image_id = false) const;
    void WriteKeyPoints(const image_t image_id, const std::vector<cv::KeyPoint>& keypoints) const;
    void WriteKeyPointsColor(const image_t image_id, const std::vector<cv::Vec3b>& keypoints) const;
    void WriteDescriptors(const image_t image_id, const cv::Mat& descriptors) const;   // CV_32F (Database.cpp:176)
    void WriteMatches(const image_t image_id1, const image_t image_id2, const std::vector<cv::DMatch>& matches) const;

    static image_pair_t ImagePairToPairId(const image_t image_id1, const image_t image_id2);
    static void PairIdToImagePair(const image_pair_t pair_id, image_t* image_id1, image_t* image_id2);
    static bool SwapImagePair(const image_t image_id1, const image_t image_id2);

private:
    struct Impl;
    Impl* impl_;
};

}  // namespace MonocularSfM
#endif
