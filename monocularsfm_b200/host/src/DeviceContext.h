// Process-wide msfm_ctx used by the drop-in classes (the reference is single-threaded; so is this).
#ifndef MSFM_HOST_DEVICE_CONTEXT_H_
#define MSFM_HOST_DEVICE_CONTEXT_H_
#include "msfm_b200.h"

namespace MonocularSfM {
namespace device {
// Lazily creates the context on device $MSFM_DEVICE (default 0).  Like the reference's SQLite error handling
// (Database.cpp:16-20) a failure is fatal: message on stderr, exit(EXIT_FAILURE) — there is no CPU fallback.
msfm_ctx* Context();
void Check(int rc, const char* what);
void Shutdown();
bool Alive();      // false before the first Context() and after Shutdown()
}  // namespace device
}  // namespace MonocularSfM
#endif
