// MonocularSfM::CeresBundelOptimizer (sic) — same class name, Parameters and Optimize() contract as the reference
// (include/Optimizer/CeresBundleOptimizer.h:14-40, src/Optimizer/CeresBundleOptimizer.cpp:188-328); the solve runs on
// the B200 through msfm_ba_* instead of Ceres.
#ifndef MSFM_HOST_CERES_BUNDLE_OPTIMIZER_H_
#define MSFM_HOST_CERES_BUNDLE_OPTIMIZER_H_
#include "Optimizer/BundleData.h"

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
    // Mutates bundle_data in place (also when it returns false); true iff the solve terminated with CONVERGENCE (:296).
    bool Optimize(BundleData& bundle_data);

    // extras (not in the reference): last solve summary
    int last_iterations() const { return last_iterations_; }
    double last_initial_cost() const { return last_initial_cost_; }
    double last_final_cost() const { return last_final_cost_; }

private:
    Parameters params_;
    int last_iterations_ = 0;
    double last_initial_cost_ = 0, last_final_cost_ = 0;
};

}  // namespace MonocularSfM
#endif
