// Type aliases of the reference API surface kept verbatim (reference: include/Common/Types.h:9-14).
#ifndef MSFM_HOST_TYPES_H_
#define MSFM_HOST_TYPES_H_
namespace MonocularSfM {
#ifndef INVALID
#define INVALID -1
#endif
typedef int image_t;
typedef int image_pair_t;
typedef int point2D_t;
typedef int point3D_t;
}  // namespace MonocularSfM
#endif
