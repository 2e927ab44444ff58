// Kernel launch counter of the library: incremented AT every launch site (each `kernel<<<...>>>(...)` statement and the
// cooperative launch of the band Cholesky are followed by MSFM_COUNT_LAUNCH()), so msfm_launch_count() is a count of the
// library's own kernels, not a hand-maintained estimate.  Library routines (cuSOLVER / cuBLAS in the fallback solvers) are not
// counted.  Process-wide: one ctx per process is the intended use; a ctx reports the launches since its creation.
#pragma once
#include <atomic>

namespace msfm {
extern std::atomic<long long> g_kernel_launches;
}
#define MSFM_COUNT_LAUNCH() (::msfm::g_kernel_launches.fetch_add(1, std::memory_order_relaxed))
