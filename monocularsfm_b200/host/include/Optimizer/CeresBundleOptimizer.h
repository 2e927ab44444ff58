// MonocularSfM::CeresBundelOptimizer (sic) — same class name, Parameters and Optimize() contract as the reference
// (include/Optimizer/CeresBundleOptimizer.h:14-40, src/Optimizer/CeresBundleOptimizer.cpp:188-328); the solve runs on
// the B200 through msfm_ba_* instead of Ceres.
#ifndef MSFM_HOST_CERES_BUNDLE_OPTIMIZER_H_
#define MSFM_HOST_CERES_BUNDLE_OPTIMIZER_H_
#include <vector>

#include "Optimizer/BundleData.h"

struct msfm_ba;       // the device-resident problem (include/msfm_b200.h)

void initLogging();   // kept for link compatibility (CeresBundleOptimizer.cpp:11-14); no-op without glog

namespace MonocularSfM {

class CeresBundelOptimizer {
public:
    struct Parameters {
        int min_observation_per_image = 10;   // unused by the reference as well
        bool refine_focal_length = false;
        double loss_function_scale = 1.0;     // unused (loss_function = nullptr, CeresBundleOptimizer.cpp:209)
    };
    struct Statistics {
        bool is_succeed = false;
    };
    CeresBundelOptimizer(const Parameters& params);
    ~CeresBundelOptimizer();
    CeresBundelOptimizer(const CeresBundelOptimizer&) = delete;
    CeresBundelOptimizer& operator=(const CeresBundelOptimizer&) = delete;
    // Mutates bundle_data in place (also when it returns false); true iff the solve terminated with CONVERGENCE (:296).
    bool Optimize(BundleData& bundle_data);

    // extras (not in the reference): last solve summary
    int last_iterations() const { return last_iterations_; }
    double last_initial_cost() const { return last_initial_cost_; }
    double last_final_cost() const { return last_final_cost_; }
    // The device problem persists across Optimize calls (MapBuilder keeps ONE optimizer for the whole reconstruction,
    // MapBuilder.cpp:92): true if the last call found the sparsity pattern of the previous one and uploaded only values;
    // host -> device bytes of the last call.
    bool last_structure_reused() const { return last_reused_; }
    long long last_h2d_bytes() const { return last_h2d_bytes_; }
    // After Optimize: the statistics Map::FilterAllPoints3D recomputes on the host (Map.cpp:793-917), served from the resident
    // problem at the optimised parameters (msfm_ba_filter_stats); arrays in the flattening order documented in BundleData.h.
    // Any output may be nullptr.  false if no problem is resident.
    bool FilterStatistics(double max_reproj_error, std::vector<unsigned char>* obs_keep, std::vector<double>* pt_mean_error,
                          std::vector<int>* pt_kept, std::vector<double>* pt_max_parallax_deg);

private:
    Parameters params_;
    int last_iterations_ = 0;
    double last_initial_cost_ = 0, last_final_cost_ = 0;
    msfm_ba* problem_ = nullptr;
    bool last_reused_ = false;
    long long last_h2d_bytes_ = 0;
};

}  // namespace MonocularSfM
#endif
