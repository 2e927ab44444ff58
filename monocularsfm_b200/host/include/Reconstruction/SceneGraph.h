// SceneGraph — correspondence graph built from the matches table the M-path writes (include/Reconstruction/SceneGraph.h
// and src/Reconstruction/SceneGraph.cpp of the reference; consumer of K1's output).  Same public API
// and counting rules as the reference (invalid indices and duplicate correspondences are dropped and un-counted,
// self-matches ignored, num_observations only after Finalize()).
//
// Storage is different on purpose: the reference keeps a vector<vector<Correspondence>> per image and scans a point's
// list linearly for duplicates on every insertion.  Here matches are appended to a flat half-edge list and compiled —
// lazily, on the first query after an insertion — into one CSR per graph (counting sort by image and point, stable, so a
// point's correspondences keep their insertion order like the reference's); duplicates are found inside the (tiny)
// per-point groups.  Loading N pairs is O(total matches) with two passes over contiguous arrays.
#ifndef MSFM_HOST_SCENE_GRAPH_H_
#define MSFM_HOST_SCENE_GRAPH_H_
#include <cstddef>
#include <unordered_map>
#include <vector>

#include "Common/Types.h"
#include "Database/Database.h"

namespace MonocularSfM {

class SceneGraph {
public:
    struct Correspondence {
        Correspondence() : image_id(INVALID), point2D_idx(INVALID) {}
        Correspondence(const image_t image_id, point2D_t point2D_idx) : image_id(image_id), point2D_idx(point2D_idx) {}
        image_t image_id;
        point2D_t point2D_idx;
    };

    SceneGraph() {}
    // every image of the database becomes a node; pairs with fewer than min_num_matches matches are ignored (:11-85)
    void Load(const cv::Ptr<Database> database, const size_t min_num_matches);
    // recounts num_observations and drops the images without any correspondence (:88-117)
    void Finalize();

    size_t NumImages() const;
    bool ExistsImage(const image_t image_id) const;
    point2D_t NumObservationsForImage(image_t image_id) const;
    point2D_t NumCorrespondencesForImage(image_t image_id) const;
    point2D_t NumCorrespondencesBetweenImages(const image_t image_id1, const image_t image_id2) const;

    void AddImage(const image_t image_id, const size_t num_points2D);
    void AddCorrespondences(const image_t image_id1, const image_t image_id2, const std::vector<cv::DMatch>& matches);

    const std::vector<Correspondence> FindCorrespondences(const image_t image_id, const point2D_t point2D_idx) const;
    std::vector<cv::DMatch> FindCorrespondencesBetweenImages(const image_t image_id1, const image_t image_id2) const;
    bool HasCorrespondences(const image_t image_id, const point2D_t point2D_idx) const;
    bool IsTwoViewObservation(const image_t image_id, const point2D_t point2D_idx) const;
    std::vector<image_t> GetAllImageIds() const;
    const std::unordered_map<image_pair_t, point2D_t> ImagePairs();

private:
    struct Node {
        point2D_t num_points = 0;
        point2D_t num_observations = 0;
        point2D_t num_correspondences = 0;
        size_t row0 = 0;            // first row of this image in the CSR (one row per 2-D point)
        bool erased = false;        // dropped by Finalize()
    };
    struct HalfEdge {               // one direction of one match
        int node;                   // dense index of the image that owns the point
        point2D_t point;
        image_t other_image;
        point2D_t other_point;
    };
    void Compile() const;           // half-edges -> CSR (drops duplicates, fixes the counters)
    const Node& NodeOf(image_t image_id) const;

    std::unordered_map<image_t, int> index_of_;     // image id -> dense node index
    mutable std::vector<Node> nodes_;
    std::vector<image_t> ids_;                      // dense node index -> image id
    mutable std::vector<HalfEdge> pending_;         // appended by AddCorrespondences, consumed by Compile()
    mutable std::vector<HalfEdge> edges_;           // compiled half-edges in CSR order
    mutable std::vector<size_t> row_start_;         // CSR offsets over all (image, point) rows, size rows + 1
    mutable std::unordered_map<image_pair_t, point2D_t> image_pairs_;
    mutable bool dirty_ = false;
};

}  // namespace MonocularSfM
#endif  // MSFM_HOST_SCENE_GRAPH_H_
