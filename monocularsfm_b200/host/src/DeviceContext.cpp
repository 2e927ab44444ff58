#include "DeviceContext.h"

#include <cstdio>
#include <cstdlib>

namespace MonocularSfM {
namespace device {

static msfm_ctx* g_ctx = nullptr;

msfm_ctx* Context() {
    if (g_ctx) return g_ctx;
    int dev = 0;
    if (const char* e = std::getenv("MSFM_DEVICE")) dev = std::atoi(e);
    const int rc = msfm_init(&g_ctx, dev);
    if (rc != MSFM_OK) {
        std::fprintf(stderr, "msfm_init(device %d) failed (%d): %s\n", dev, rc, msfm_last_error(nullptr));
        std::exit(EXIT_FAILURE);
    }
    std::atexit(Shutdown);
    return g_ctx;
}

void Check(int rc, const char* what) {
    if (rc == MSFM_OK) return;
    std::fprintf(stderr, "%s failed (%d): %s\n", what, rc, msfm_last_error(g_ctx));
    std::exit(EXIT_FAILURE);
}

bool Alive() { return g_ctx != nullptr; }

void Shutdown() {
    if (g_ctx) msfm_destroy(g_ctx);
    g_ctx = nullptr;
}

}  // namespace device
}  // namespace MonocularSfM
