// SceneGraph on flat half-edge lists compiled into one CSR (see Reconstruction/SceneGraph.h for the why).
// Behaviour follows the reference's src/Reconstruction/SceneGraph.cpp line by line where it is observable:
//   Load             :11-85    every image is a node, pairs below min_num_matches ignored, Finalize() NOT called
//   AddCorrespondences :170-251 self-matches ignored; a match with an out-of-range index or one that repeats an
//                              existing correspondence is dropped and taken back out of all three counters
//   Finalize         :88-117   num_observations = points with at least one correspondence; isolated images erased
#include "Reconstruction/SceneGraph.h"

#include <cassert>
#include <cstdio>
#include <iostream>

using namespace MonocularSfM;

void SceneGraph::Load(const cv::Ptr<Database> database, const size_t min_num_matches) {
    std::cout << "Loading matches..." << std::flush;
    const std::vector<std::pair<image_pair_t, std::vector<cv::DMatch>>> image_pairs = database->ReadAllMatches();
    std::cout << "Total image pairs : " << image_pairs.size() << std::endl;
    const std::vector<Database::Image> images = database->ReadAllImages();
    std::cout << "Total images : " << images.size() << std::endl;
    std::cout << "Building scene graph..." << std::flush;
    for (const Database::Image& image : images) AddImage(image.id, database->NumKeyPoints(image.id));
    std::cout << NumImages() << std::endl;
    size_t ignored = 0;
    for (const auto& pair : image_pairs) {
        if (pair.second.size() >= min_num_matches) {
            image_t id1, id2;
            Database::PairIdToImagePair(pair.first, &id1, &id2);
            AddCorrespondences(id1, id2, pair.second);
        } else {
            ++ignored;
        }
    }
    Compile();
    std::cout << "Total image pairs : " << image_pairs.size() << ".  Ignored : " << ignored << std::endl;
}

void SceneGraph::AddImage(const image_t image_id, const size_t num_points2D) {
    assert(!ExistsImage(image_id));
    index_of_[image_id] = static_cast<int>(nodes_.size());
    Node n;
    n.num_points = static_cast<point2D_t>(num_points2D);
    nodes_.push_back(n);
    ids_.push_back(image_id);
    dirty_ = true;
}

void SceneGraph::AddCorrespondences(const image_t image_id1, const image_t image_id2, const std::vector<cv::DMatch>& matches) {
    if (image_id1 == image_id2) {
        std::fprintf(stderr, "WARNING : Cannot use self-matches for image_id = %d", image_id1);
        return;
    }
    assert(ExistsImage(image_id1));
    assert(ExistsImage(image_id2));
    const int n1 = index_of_.at(image_id1), n2 = index_of_.at(image_id2);
    point2D_t& between = image_pairs_[Database::ImagePairToPairId(image_id1, image_id2)];
    for (const cv::DMatch& m : matches) {
        const point2D_t i1 = m.queryIdx, i2 = m.trainIdx;
        // the reference compares the (signed) index with corrs.size() after conversion to size_t: negative = invalid
        const bool ok1 = i1 >= 0 && i1 < nodes_[n1].num_points, ok2 = i2 >= 0 && i2 < nodes_[n2].num_points;
        if (!ok1) std::fprintf(stderr, "WARNING : point2D_idx = %d in image_id = %d does not exist\n", i1, image_id1);
        if (!ok2) std::fprintf(stderr, "WARNING : point2D_idx = %d in image_id = %d does not exist\n", i2, image_id2);
        if (!ok1 || !ok2) continue;
        // counted now, taken back by Compile() if it turns out to repeat an earlier correspondence
        nodes_[n1].num_correspondences += 1;
        nodes_[n2].num_correspondences += 1;
        between += 1;
        pending_.push_back(HalfEdge{n1, i1, image_id2, i2});
        pending_.push_back(HalfEdge{n2, i2, image_id1, i1});
    }
    dirty_ = true;
}

// Rebuild the CSR from the compiled edges plus the pending ones.  Stable counting sort by (node, point): a point's
// correspondences keep insertion order.  Inside a row, an edge that repeats an earlier (other image, other point) is a
// duplicate: dropped, and un-counted once per MATCH (its mirror edge is dropped in the mirror row).
void SceneGraph::Compile() const {
    if (!dirty_) return;
    size_t rows = 0;
    for (Node& n : nodes_) { n.row0 = rows; rows += static_cast<size_t>(n.num_points); }
    std::vector<HalfEdge> all;
    all.reserve(edges_.size() + pending_.size());
    all.insert(all.end(), edges_.begin(), edges_.end());        // already de-duplicated, still first in their rows
    all.insert(all.end(), pending_.begin(), pending_.end());
    pending_.clear();
    std::vector<size_t> start(rows + 1, 0);
    for (const HalfEdge& e : all) start[nodes_[e.node].row0 + e.point + 1] += 1;
    for (size_t r = 0; r < rows; ++r) start[r + 1] += start[r];
    std::vector<HalfEdge> sorted(all.size());
    {
        std::vector<size_t> cursor(start.begin(), start.end() - 1);
        for (const HalfEdge& e : all) sorted[cursor[nodes_[e.node].row0 + e.point]++] = e;
    }
    // drop duplicates row by row, compacting in place
    edges_.clear();
    edges_.reserve(sorted.size());
    row_start_.assign(rows + 1, 0);
    for (size_t r = 0; r < rows; ++r) {
        row_start_[r] = edges_.size();
        for (size_t k = start[r]; k < start[r + 1]; ++k) {
            const HalfEdge& e = sorted[k];
            bool dup = false;
            for (size_t j = row_start_[r]; j < edges_.size() && !dup; ++j)
                dup = edges_[j].other_image == e.other_image && edges_[j].other_point == e.other_point;
            if (dup) {
                // every duplicate MATCH shows up as one dropped edge in each of its two rows: un-count the image here
                // and half of the pair counter twice
                nodes_[e.node].num_correspondences -= 1;
                if (ids_[e.node] < e.other_image) {
                    image_pairs_[Database::ImagePairToPairId(ids_[e.node], e.other_image)] -= 1;
                    std::fprintf(stderr, "WARNING : Duplicate correspondence betweenpoint2D_idx = %d in image_id = %d and point2D_idx = %d in image_id = %d\n",
                                 e.point, ids_[e.node], e.other_point, e.other_image);
                }
                continue;
            }
            edges_.push_back(e);
        }
    }
    row_start_[rows] = edges_.size();
    dirty_ = false;
}

void SceneGraph::Finalize() {
    Compile();
    bool erased_any = false;
    for (Node& n : nodes_) {
        if (n.erased) continue;
        n.num_observations = 0;
        for (point2D_t p = 0; p < n.num_points; ++p)
            if (row_start_[n.row0 + p + 1] > row_start_[n.row0 + p]) n.num_observations += 1;
        if (n.num_observations == 0) { n.erased = true; erased_any = true; }
    }
    if (erased_any)
        for (size_t i = 0; i < nodes_.size(); ++i)
            if (nodes_[i].erased) index_of_.erase(ids_[i]);
}

const SceneGraph::Node& SceneGraph::NodeOf(image_t image_id) const { return nodes_[index_of_.at(image_id)]; }

size_t SceneGraph::NumImages() const { return index_of_.size(); }
bool SceneGraph::ExistsImage(const image_t image_id) const { return index_of_.count(image_id) > 0; }

point2D_t SceneGraph::NumObservationsForImage(image_t image_id) const {
    assert(ExistsImage(image_id));
    return NodeOf(image_id).num_observations;
}

point2D_t SceneGraph::NumCorrespondencesForImage(image_t image_id) const {
    assert(ExistsImage(image_id));
    Compile();
    return NodeOf(image_id).num_correspondences;
}

point2D_t SceneGraph::NumCorrespondencesBetweenImages(const image_t image_id1, const image_t image_id2) const {
    assert(ExistsImage(image_id1));
    assert(ExistsImage(image_id2));
    Compile();
    const auto it = image_pairs_.find(Database::ImagePairToPairId(image_id1, image_id2));
    return it == image_pairs_.end() ? 0 : it->second;
}

const std::vector<SceneGraph::Correspondence> SceneGraph::FindCorrespondences(const image_t image_id, const point2D_t point2D_idx) const {
    assert(ExistsImage(image_id));
    Compile();
    const Node& n = NodeOf(image_id);
    assert(point2D_idx >= 0 && point2D_idx < n.num_points);
    std::vector<Correspondence> out;
    for (size_t k = row_start_[n.row0 + point2D_idx]; k < row_start_[n.row0 + point2D_idx + 1]; ++k)
        out.push_back(Correspondence(edges_[k].other_image, edges_[k].other_point));
    return out;
}

std::vector<cv::DMatch> SceneGraph::FindCorrespondencesBetweenImages(const image_t image_id1, const image_t image_id2) const {
    Compile();
    std::vector<cv::DMatch> found;
    const Node& n = NodeOf(image_id1);
    for (point2D_t p = 0; p < n.num_points; ++p)
        for (size_t k = row_start_[n.row0 + p]; k < row_start_[n.row0 + p + 1]; ++k)
            if (edges_[k].other_image == image_id2) found.push_back(cv::DMatch(p, edges_[k].other_point, 0));
    return found;
}

bool SceneGraph::HasCorrespondences(const image_t image_id, const point2D_t point2D_idx) const {
    Compile();
    const Node& n = NodeOf(image_id);
    return row_start_[n.row0 + point2D_idx + 1] > row_start_[n.row0 + point2D_idx];
}

bool SceneGraph::IsTwoViewObservation(const image_t image_id, const point2D_t point2D_idx) const {
    Compile();
    const Node& n = NodeOf(image_id);
    const size_t b = row_start_[n.row0 + point2D_idx], e = row_start_[n.row0 + point2D_idx + 1];
    if (e - b != 1) return false;
    const Node& o = NodeOf(edges_[b].other_image);
    return row_start_[o.row0 + edges_[b].other_point + 1] - row_start_[o.row0 + edges_[b].other_point] == 1;
}

std::vector<image_t> SceneGraph::GetAllImageIds() const {
    std::vector<image_t> out;
    for (size_t i = 0; i < nodes_.size(); ++i)
        if (!nodes_[i].erased) out.push_back(ids_[i]);
    return out;
}

const std::unordered_map<image_pair_t, point2D_t> SceneGraph::ImagePairs() {
    Compile();
    return image_pairs_;
}
