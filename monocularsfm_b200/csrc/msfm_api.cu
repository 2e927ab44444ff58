// C-ABI entry points (include/msfm_b200.h): context + M-path orchestration.
// The product path has no CPU fallback: every entry point needs a live CUDA device.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <new>

#include "ctx.hpp"
#include "launch_count.hpp"

namespace msfm {
struct TempImgs {
    uint8_t* sw; int32_t* nrm; int32_t* perm; int32_t* used; int32_t* flags; int32_t n_pad_t;
};
cudaError_t launch_setup_temp_imgs(ImgDev*, int, const SegDev*, int, TempImgs, cudaStream_t);
cudaError_t launch_gather_candidates(const ImgDev*, const SegDev*, int, const int32_t*, TempImgs, cudaStream_t);
cudaError_t launch_match_tile_kernel(const ImgDev*, const UnitDev*, int, int, int32_t*, int32_t*, int32_t*, void*, int, cudaStream_t);
size_t match_tile_item_bytes(int num_units);
cudaError_t launch_desc_format(const uint8_t*, int, int, uint8_t*, uint8_t*, int32_t*, int32_t*, int32_t*, int32_t*, int32_t*, unsigned long long*,
                               int32_t*, int32_t*, int32_t*, cudaStream_t);
cudaError_t launch_desc_quantize(const float*, int, int, int32_t*, uint8_t*, cudaStream_t);
cudaError_t launch_desc_normalize(float*, int, int, cudaStream_t);
cudaError_t launch_build_units(const SegDev*, int, int, UnitDev*, cudaStream_t);
cudaError_t launch_resolve_rows(const ImgDev*, const UnitDev*, int, int, const int32_t*, const int32_t*, const int32_t*,
                                MatchOpts, int32_t*, int32_t*, int32_t*, int32_t*, int32_t*, unsigned int*, cudaStream_t);
cudaError_t launch_exact_rows(const ImgDev*, const UnitDev*, const int32_t*, const unsigned int*, int, MatchOpts,
                              int32_t*, int32_t*, int32_t*, int32_t*, int, cudaStream_t);
cudaError_t launch_count_scan_write(const ImgDev*, const SegDev*, int, MatchOpts, const int32_t*, const int32_t*,
                                    int32_t*, long long*, long long*, long long, int32_t*, float*, cudaStream_t);
}  // namespace msfm

namespace msfm { void ctx_destroy_solver(msfm_ctx* c); }
using namespace msfm;

static thread_local std::string g_init_error;
static constexpr int kMaxUnitsPerBatch = 1 << 17;     // 131072 units = 16.7 M rows of scratch per batch
static constexpr int32_t kTmpIdA = -1000001, kTmpIdB = -1000002;

namespace msfm {
std::atomic<long long> g_kernel_launches{0};      // launch_count.hpp: incremented at every launch site of the library
}

extern "C" {

const char* msfm_version(void) { return "0.1.0"; }

int msfm_init(msfm_ctx** out, int device_id) {
    if (!out) return MSFM_E_INVALID;
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        g_init_error = std::string("no CUDA device: ") + cudaGetErrorString(e);
        return MSFM_E_NO_DEVICE;
    }
    if (device_id < 0 || device_id >= ndev) {
        g_init_error = "device_id out of range";
        return MSFM_E_INVALID;
    }
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device_id)) != cudaSuccess) {
        g_init_error = cudaGetErrorString(e);
        return MSFM_E_CUDA;
    }
    if (prop.major != 10) {
        g_init_error = "this library contains sm_100a code only (tcgen05/TMEM); device is sm_" +
                       std::to_string(prop.major) + std::to_string(prop.minor);
        return MSFM_E_NO_DEVICE;
    }
    if ((e = cudaSetDevice(device_id)) != cudaSuccess) {
        g_init_error = cudaGetErrorString(e);
        return MSFM_E_CUDA;
    }
    msfm_ctx* c = new (std::nothrow) msfm_ctx();
    if (c) c->launches = static_cast<int64_t>(msfm::g_kernel_launches.load());      // msfm_launch_count reports launches since creation
    if (!c) return MSFM_E_CUDA;
    c->device = device_id;
    c->num_sms = prop.multiProcessorCount;
    c->h_stage.pinned_host = true;
    if ((e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking)) != cudaSuccess) {
        g_init_error = cudaGetErrorString(e);
        if (c->stream) cudaStreamDestroy(c->stream);
        delete c;
        return MSFM_E_CUDA;
    }
    for (int i = 0; i < 2; ++i) {
        cudaEventCreateWithFlags(&c->raw_ready[i], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&c->raw_free[i], cudaEventDisableTiming);
    }
    *out = c;
    return MSFM_OK;
}

void msfm_destroy(msfm_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    msfm_comm_destroy(c);
    msfm::ctx_destroy_solver(c);
    c->prof_collect();
    for (cudaEvent_t e : c->prof_pool) cudaEventDestroy(e);
    for (auto& im : c->imgs)
        if (im.block) cudaFree(im.block);
    GrowBuf* bufs[] = {&c->d_imgs, &c->d_raw, &c->d_fmt, &c->d_temp, &c->h_stage, &c->d_segs, &c->d_units, &c->d_items, &c->d_res, &c->d_m, &c->d_exact,
                       &c->d_counts, &c->d_misc, &c->d_out_offsets, &c->d_out_matches, &c->d_out_dist, &c->d_ba_r, &c->d_ba_J,
                       &c->d_kp_tab, &c->d_kp_slots, &c->d_vf_in, &c->d_vf_out};
    for (auto& kv : c->kps)
        if (kv.second.xy) cudaFree(kv.second.xy);
    for (GrowBuf* b : bufs) b->release();
    for (int i = 0; i < 2; ++i) {
        c->d_rawq[i].release();
        if (c->raw_ready[i]) cudaEventDestroy(c->raw_ready[i]);
        if (c->raw_free[i]) cudaEventDestroy(c->raw_free[i]);
    }
    if (c->copy_stream) { cudaStreamSynchronize(c->copy_stream); cudaStreamDestroy(c->copy_stream); }
    cudaStreamDestroy(c->stream);
    delete c;
}

const char* msfm_last_error(const msfm_ctx* c) { return c ? c->err.c_str() : g_init_error.c_str(); }

int msfm_sync(msfm_ctx* c) {
    if (!c) return MSFM_E_INVALID;
    MSFM_CUDA(c, cudaStreamSynchronize(c->stream));
    return MSFM_OK;
}
void* msfm_stream(msfm_ctx* c) { return c ? static_cast<void*>(c->stream) : nullptr; }
int64_t msfm_launch_count(const msfm_ctx* c) { return c ? static_cast<int64_t>(msfm::g_kernel_launches.load()) - c->launches : 0; }

int msfm_prof_enable(msfm_ctx* c, int on) {
    if (!c) return MSFM_E_INVALID;
    c->prof_on = on != 0;
    return MSFM_OK;
}
int msfm_prof_reset(msfm_ctx* c) {
    if (!c) return MSFM_E_INVALID;
    MSFM_CUDA(c, cudaSetDevice(c->device));
    MSFM_CUDA(c, cudaStreamSynchronize(c->stream));
    c->prof_collect();
    for (int i = 0; i < MSFM_PROF_NCAT; ++i) { c->prof_ms[i] = 0; c->prof_n[i] = 0; }
    return MSFM_OK;
}
int msfm_prof_read(msfm_ctx* c, double ms[MSFM_PROF_NCAT], int64_t n[MSFM_PROF_NCAT]) {
    if (!c || !ms || !n) return MSFM_E_INVALID;
    MSFM_CUDA(c, cudaSetDevice(c->device));
    MSFM_CUDA(c, cudaStreamSynchronize(c->stream));
    c->prof_collect();
    for (int i = 0; i < MSFM_PROF_NCAT; ++i) { ms[i] = c->prof_ms[i]; n[i] = c->prof_n[i]; }
    return MSFM_OK;
}

// ------------------------------------------------------------------------------------------------ uploads
// One allocation per image: sw | ext | cg | nrm | perm (match_types.cuh), every part 256-B aligned.
struct ImgLayout {
    size_t off_ext, off_cg, off_nrm, off_perm, off_inv, off_used, total;
};
static ImgLayout img_layout(int32_t n_pad) {
    ImgLayout L;
    const size_t sw = static_cast<size_t>(n_pad) * 128;
    const size_t ext = static_cast<size_t>((n_pad + 1023) / 1024) * 256 * 128;
    const size_t cg = (static_cast<size_t>(n_pad) / 32 * 4 + 255) / 256 * 256;
    L.off_ext = sw;
    L.off_cg = L.off_ext + ext;
    L.off_nrm = L.off_cg + cg;
    L.off_perm = L.off_nrm + static_cast<size_t>(n_pad) * 4;
    L.off_inv = L.off_perm + static_cast<size_t>(n_pad) * 4;
    L.off_used = L.off_inv + static_cast<size_t>(n_pad) * 4;
    L.total = L.off_used + 256;
    return L;
}
static int32_t padded_count(int32_t n) { return n > 0 ? (n + kMaxPadPerImage + 255) / 256 * 256 : 0; }

// src_f32 != nullptr: float32 host descriptors, converted on the device (f32_mode: 0 decide per set, 1 always quantise)
// normalize: 0 none, 1 L1-root, 2 L2 (raw SIFT rows, extraction-time normalisation on the device); normalized_out: optional host
// copy of the normalised floats (what the reference writes to its database)
static int upload_common(msfm_ctx* c, int32_t image_id, const uint8_t* src, int32_t n, bool src_on_device,
                         const float* src_f32 = nullptr, int f32_mode = 0, int normalize = 0, float* normalized_out = nullptr) {
    if (!c || n < 0 || (n > 0 && !src && !src_f32))
        return c ? c->fail(MSFM_E_INVALID, "msfm_desc_upload: bad arguments") : MSFM_E_INVALID;
    MSFM_CUDA(c, cudaSetDevice(c->device));
    const int32_t n_pad = padded_count(n);
    int slot;
    auto it = c->slot_of.find(image_id);
    if (it != c->slot_of.end()) {
        slot = it->second;
    } else if (!c->free_slots.empty()) {
        slot = c->free_slots.back();
        c->free_slots.pop_back();
        c->slot_of[image_id] = slot;
    } else {
        slot = static_cast<int>(c->imgs.size());
        c->imgs.emplace_back();
        c->slot_of[image_id] = slot;
    }
    ImgHost& im = c->imgs[slot];
    if (im.block && im.n_pad != n_pad) {
        // a kernel of an earlier call may still read the old block
        MSFM_CUDA(c, cudaStreamSynchronize(c->stream));
        MSFM_CUDA(c, cudaFree(im.block));
        im.block = nullptr;
    }
    const ImgLayout L = img_layout(n_pad);
    if (!im.block && n_pad > 0) MSFM_CUDA(c, cudaMalloc(&im.block, L.total));
    im.n = n;
    im.n_pad = n_pad;
    im.live = true;
    c->imgs_dirty = true;
    if (n_pad == 0) return MSFM_OK;
    const uint8_t* raw = src;
    int raw_q = -1;
    if (src_f32) {
        // staging: [n][128] u8 | 256-B aligned [n][128] f32 | flag
        const size_t u8_bytes = (static_cast<size_t>(n) * 128 + 255) / 256 * 256, f32_bytes = static_cast<size_t>(n) * 128 * 4;
        MSFM_CUDA(c, c->d_raw.reserve(u8_bytes + f32_bytes + 256));
        float* stage = reinterpret_cast<float*>(c->d_raw.as<uint8_t>() + u8_bytes);
        MSFM_CUDA(c, cudaMemcpyAsync(stage, src_f32, f32_bytes, cudaMemcpyHostToDevice, c->stream));
        if (normalize) {
            c->prof_begin(MSFM_PROF_DESC_FORMAT);
            MSFM_CUDA(c, launch_desc_normalize(stage, n, normalize, c->stream));
            c->prof_end();
            if (normalized_out) MSFM_CUDA(c, cudaMemcpyAsync(normalized_out, stage, f32_bytes, cudaMemcpyDeviceToHost, c->stream));
        }
        c->prof_begin(MSFM_PROF_DESC_FORMAT);
        MSFM_CUDA(c, launch_desc_quantize(stage, n, f32_mode, reinterpret_cast<int32_t*>(c->d_raw.as<uint8_t>() + u8_bytes + f32_bytes),
                                          c->d_raw.as<uint8_t>(), c->stream));
        c->prof_end();
        raw = c->d_raw.as<uint8_t>();
    } else if (!src_on_device) {
        // double-buffered staging on the copy stream: this copy runs under the formatting kernels of the previous upload
        raw_q = static_cast<int>(c->raw_turn++ & 1u);
        GrowBuf& q = c->d_rawq[raw_q];
        if (c->raw_free_set[raw_q]) MSFM_CUDA(c, cudaStreamWaitEvent(c->copy_stream, c->raw_free[raw_q], 0));
        if (q.cap < static_cast<size_t>(n) * 128) {
            MSFM_CUDA(c, cudaStreamSynchronize(c->stream));        // kernels of an earlier upload may still read the old buffer
            MSFM_CUDA(c, q.reserve(static_cast<size_t>(n) * 128));
        }
        MSFM_CUDA(c, cudaMemcpyAsync(q.p, src, static_cast<size_t>(n) * 128, cudaMemcpyHostToDevice, c->copy_stream));
        MSFM_CUDA(c, cudaEventRecord(c->raw_ready[raw_q], c->copy_stream));
        MSFM_CUDA(c, cudaStreamWaitEvent(c->stream, c->raw_ready[raw_q], 0));
        raw = q.as<uint8_t>();
    }
    uint8_t* sw = static_cast<uint8_t*>(im.block);
    // formatting scratch: keys [n] u64 | nrm_orig [n] | pos_of [n] | bucket counts [8]
    MSFM_CUDA(c, c->d_fmt.reserve(static_cast<size_t>(n) * 16 + 64));
    unsigned long long* keys = c->d_fmt.as<unsigned long long>();
    int32_t* nrm_orig = reinterpret_cast<int32_t*>(keys + n);
    int32_t* pos_of = nrm_orig + n;
    int32_t* bucket_cnt = pos_of + n;
    c->prof_begin(MSFM_PROF_DESC_FORMAT);
    MSFM_CUDA(c, launch_desc_format(raw, n, n_pad, sw, sw + L.off_ext, reinterpret_cast<int32_t*>(sw + L.off_cg),
                                    reinterpret_cast<int32_t*>(sw + L.off_nrm), reinterpret_cast<int32_t*>(sw + L.off_perm),
                                    reinterpret_cast<int32_t*>(sw + L.off_inv), reinterpret_cast<int32_t*>(sw + L.off_used), keys, nrm_orig, pos_of, bucket_cnt, c->stream));
    c->prof_end();
    if (raw_q >= 0) {
        MSFM_CUDA(c, cudaEventRecord(c->raw_free[raw_q], c->stream));
        c->raw_free_set[raw_q] = true;
    }
    // word after `used`: 1 = the float32 set converted exactly (or uint8 upload), 0 = it was quantised
    int32_t* exact_flag = reinterpret_cast<int32_t*>(sw + L.off_used) + 1;
    if (src_f32 && f32_mode == 0) {
        const size_t u8_bytes = (static_cast<size_t>(n) * 128 + 255) / 256 * 256, f32_bytes = static_cast<size_t>(n) * 128 * 4;
        MSFM_CUDA(c, cudaMemcpyAsync(exact_flag, c->d_raw.as<uint8_t>() + u8_bytes + f32_bytes, sizeof(int32_t),
                                     cudaMemcpyDeviceToDevice, c->stream));
    } else {
        MSFM_CUDA(c, cudaMemsetAsync(exact_flag, src_f32 ? 0x00 : 0x01, sizeof(int32_t), c->stream));   // any non-zero word = exact
    }
    return MSFM_OK;
}

int msfm_desc_upload_u8(msfm_ctx* c, int32_t image_id, const uint8_t* desc_host, int32_t n) {
    if (c && image_id < 0) return c->fail(MSFM_E_INVALID, "image_id must be >= 0");
    return upload_common(c, image_id, desc_host, n, false);
}
int msfm_desc_upload_u8_dev(msfm_ctx* c, int32_t image_id, const uint8_t* desc_dev, int32_t n) {
    if (c && image_id < 0) return c->fail(MSFM_E_INVALID, "image_id must be >= 0");
    return upload_common(c, image_id, desc_dev, n, true);
}
int msfm_desc_upload_f32(msfm_ctx* c, int32_t image_id, const float* desc_host, int32_t n, int32_t mode) {
    if (c && image_id < 0) return c->fail(MSFM_E_INVALID, "image_id must be >= 0");
    if (c && mode != 0 && mode != 1) return c->fail(MSFM_E_INVALID, "msfm_desc_upload_f32: mode must be 0 or 1");
    if (c && n > 0 && !desc_host) return c->fail(MSFM_E_INVALID, "msfm_desc_upload_f32: null descriptors");
    return upload_common(c, image_id, nullptr, n, false, desc_host, mode);
}
int msfm_desc_upload_raw_f32(msfm_ctx* c, int32_t image_id, const float* desc_host, int32_t n, int32_t normalization, float* normalized_out) {
    if (c && image_id < 0) return c->fail(MSFM_E_INVALID, "image_id must be >= 0");
    if (c && normalization != MSFM_NORM_L1_ROOT && normalization != MSFM_NORM_L2)
        return c->fail(MSFM_E_INVALID, "msfm_desc_upload_raw_f32: normalization must be MSFM_NORM_L1_ROOT or MSFM_NORM_L2");
    if (c && n > 0 && !desc_host) return c->fail(MSFM_E_INVALID, "msfm_desc_upload_raw_f32: null descriptors");
    int rc = upload_common(c, image_id, nullptr, n, false, desc_host, 1, normalization, normalized_out);
    if (rc == MSFM_OK && normalized_out && n > 0) MSFM_CUDA(c, cudaStreamSynchronize(c->stream));     // the host copy is complete on return
    return rc;
}
int msfm_desc_quantised(msfm_ctx* c, int32_t image_id) {
    if (!c) return MSFM_E_INVALID;
    auto it = c->slot_of.find(image_id);
    if (it == c->slot_of.end()) return c->fail(MSFM_E_NOT_FOUND, "image %d not resident", image_id);
    const ImgHost& im = c->imgs[it->second];
    if (!im.block) return 0;
    MSFM_CUDA(c, cudaSetDevice(c->device));
    int32_t exact = 1;
    const ImgLayout L = img_layout(im.n_pad);
    MSFM_CUDA(c, cudaMemcpyAsync(&exact, static_cast<uint8_t*>(im.block) + L.off_used + 4, sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
    MSFM_CUDA(c, cudaStreamSynchronize(c->stream));
    return exact ? 0 : 1;
}
int msfm_desc_count(msfm_ctx* c, int32_t image_id) {
    if (!c) return MSFM_E_INVALID;
    auto it = c->slot_of.find(image_id);
    if (it == c->slot_of.end()) return c->fail(MSFM_E_NOT_FOUND, "image %d not resident", image_id);
    return c->imgs[it->second].n;
}
int msfm_desc_release(msfm_ctx* c, int32_t image_id) {
    if (!c) return MSFM_E_INVALID;
    auto it = c->slot_of.find(image_id);
    if (it == c->slot_of.end()) return c->fail(MSFM_E_NOT_FOUND, "image %d not resident", image_id);
    MSFM_CUDA(c, cudaSetDevice(c->device));
    MSFM_CUDA(c, cudaStreamSynchronize(c->stream));
    ImgHost& im = c->imgs[it->second];
    if (im.block) MSFM_CUDA(c, cudaFree(im.block));
    im = ImgHost();
    c->free_slots.push_back(it->second);
    c->slot_of.erase(it);
    c->imgs_dirty = true;
    return MSFM_OK;
}
int msfm_desc_release_all(msfm_ctx* c) {
    if (!c) return MSFM_E_INVALID;
    MSFM_CUDA(c, cudaSetDevice(c->device));
    MSFM_CUDA(c, cudaStreamSynchronize(c->stream));
    for (auto& im : c->imgs)
        if (im.block) cudaFree(im.block);
    c->imgs.clear();
    c->free_slots.clear();
    c->slot_of.clear();
    c->imgs_dirty = true;
    return MSFM_OK;
}

// ------------------------------------------------------------------------------------------------ matching core
static int sync_img_table(msfm_ctx* c) {
    if (!c->imgs_dirty) return MSFM_OK;
    std::vector<ImgDev> tab(c->imgs.size());
    for (size_t i = 0; i < c->imgs.size(); ++i) {
        const ImgHost& im = c->imgs[i];
        uint8_t* sw = static_cast<uint8_t*>(im.block);
        const ImgLayout L = img_layout(im.n_pad);
        tab[i].sw = sw;
        tab[i].ext = sw ? sw + L.off_ext : nullptr;
        tab[i].cg = sw ? reinterpret_cast<const int32_t*>(sw + L.off_cg) : nullptr;
        tab[i].nrm = sw ? reinterpret_cast<const int32_t*>(sw + L.off_nrm) : nullptr;
        tab[i].perm = sw ? reinterpret_cast<const int32_t*>(sw + L.off_perm) : nullptr;
        tab[i].inv = sw ? reinterpret_cast<const int32_t*>(sw + L.off_inv) : nullptr;
        tab[i].used = sw ? reinterpret_cast<const int32_t*>(sw + L.off_used) : nullptr;
        tab[i].n = im.n;
        tab[i].n_pad = im.n_pad;
    }
    MSFM_CUDA(c, cudaStreamSynchronize(c->stream));
    MSFM_CUDA(c, c->d_imgs.reserve((tab.size() + 1) * sizeof(ImgDev)));    // never shrinks: match_core reserves the temp slots
    if (!tab.empty())
        MSFM_CUDA(c, cudaMemcpy(c->d_imgs.p, tab.data(), tab.size() * sizeof(ImgDev), cudaMemcpyHostToDevice));
    c->imgs_dirty = false;
    return MSFM_OK;
}

struct RowDump {          // optional per-row readback for the knn2 API (single pair, direction 0)
    int32_t* j0 = nullptr;    // [n][2]
    int32_t* d1 = nullptr;
    int32_t* d2 = nullptr;
    int32_t n = 0;
};

// mode: 0 tensor path, 1 exact scan of every row.
static int match_core(msfm_ctx* c, const int32_t* pairs, int32_t P, const msfm_match_options* uopt, int exact_second,
                      int mode, long long* out_offsets_dev, int32_t* out_matches_dev, float* out_dist_dev,
                      long long capacity, long long* total_out, RowDump* dump) {
    if (!c) return MSFM_E_INVALID;
    if (P < 0 || (P > 0 && !pairs) || !uopt) return c->fail(MSFM_E_INVALID, "msfm_match_pairs: bad arguments");
    if (!(uopt->distance_ratio >= 0.0f) || uopt->reserved != 0) return c->fail(MSFM_E_INVALID, "bad match options");
    MSFM_CUDA(c, cudaSetDevice(c->device));
    MatchOpts opt;
    opt.max_distance = uopt->max_distance;
    opt.ratio = uopt->distance_ratio;
    opt.cross_check = uopt->cross_check ? 1 : 0;
    opt.quirks = uopt->opencv_quirks ? 1 : 0;
    opt.exact_second = exact_second;
    const int spp = opt.cross_check ? 2 : 1;   // segments per pair

    // ---- segments + batches (host).  Per batch the segment table is [forward segments of its pairs][reverse segments]:
    //      the forward pass runs first; the reverse pass (cross-check only) runs on the train rows the forward pass
    //      matched, gathered into temporary query images (match_post.cu: gather_candidates_kernel).
    struct Batch { int first_pair, npairs, units_fwd, units_rev; };
    std::vector<Batch> batches;
    std::vector<int> slot1(P), slot2(P);
    int n2_max = 0;
    {
        Batch cur{0, 0, 0, 0};
        for (int p = 0; p < P; ++p) {
            auto i1 = c->slot_of.find(pairs[2 * p]), i2 = c->slot_of.find(pairs[2 * p + 1]);
            if (i1 == c->slot_of.end() || i2 == c->slot_of.end())
                return c->fail(MSFM_E_NOT_FOUND, "pair %d: image %d or %d not resident", p, pairs[2 * p], pairs[2 * p + 1]);
            slot1[p] = i1->second; slot2[p] = i2->second;
            const int n2 = c->imgs[slot2[p]].n;
            n2_max = std::max(n2_max, n2);
            const int u12 = c->imgs[slot1[p]].n_pad / kUnitRows;
            const int u21 = opt.cross_check ? 2 * ((n2 + 2 * kUnitRows - 1) / (2 * kUnitRows)) : 0;   // K1 works on unit pairs
            if (cur.npairs > 0 && cur.units_fwd + cur.units_rev + u12 + u21 > kMaxUnitsPerBatch) {
                batches.push_back(cur);
                cur = Batch{p, 0, 0, 0};
            }
            cur.units_fwd += u12; cur.units_rev += u21; cur.npairs += 1;
        }
        if (cur.npairs > 0) batches.push_back(cur);
    }
    int max_units = 1, max_pairs = 1;
    for (const Batch& b : batches) {
        max_units = std::max(max_units, b.units_fwd + b.units_rev);
        max_pairs = std::max(max_pairs, b.npairs);
    }
    const int first_temp_slot = static_cast<int>(c->imgs.size());
    std::vector<SegDev> segs(static_cast<size_t>(P) * spp);
    for (const Batch& b : batches) {
        SegDev* bs = segs.data() + static_cast<size_t>(b.first_pair) * spp;
        int ub = 0;
        for (int k = 0; k < b.npairs; ++k) {
            const int p = b.first_pair + k;
            SegDev& a = bs[k];
            a.q_slot = slot1[p]; a.t_slot = slot2[p]; a.unit_base = ub; a.n_units = c->imgs[slot1[p]].n_pad / kUnitRows;
            ub += a.n_units;
        }
        if (opt.cross_check)
            for (int k = 0; k < b.npairs; ++k) {
                const int p = b.first_pair + k;
                SegDev& r = bs[b.npairs + k];
                r.q_slot = first_temp_slot + k;          // temporary query image of pair k: matched rows of image 2
                r.t_slot = slot1[p];
                r.unit_base = ub;
                r.n_units = 2 * ((c->imgs[slot2[p]].n + 2 * kUnitRows - 1) / (2 * kUnitRows));
                ub += r.n_units;
            }
    }
    // image table: persistent slots + one temporary entry per pair of a batch
    {
        const size_t need = (static_cast<size_t>(first_temp_slot) + max_pairs + 1) * sizeof(ImgDev);
        if (need > c->d_imgs.cap) {
            MSFM_CUDA(c, cudaStreamSynchronize(c->stream));
            MSFM_CUDA(c, c->d_imgs.reserve(need));
            c->imgs_dirty = true;
        }
    }
    int rc = sync_img_table(c);
    if (rc) return rc;
    MSFM_CUDA(c, cudaStreamSynchronize(c->stream));   // scratch of an earlier call is free now

    const size_t rows = static_cast<size_t>(max_units) * kUnitRows;
    const size_t nb = std::max<size_t>(1, batches.size());
    MSFM_CUDA(c, c->d_segs.reserve(std::max<size_t>(1, segs.size()) * sizeof(SegDev)));
    MSFM_CUDA(c, c->d_units.reserve(static_cast<size_t>(max_units) * sizeof(UnitDev)));
    MSFM_CUDA(c, c->d_items.reserve(match_tile_item_bytes(max_units)));
    MSFM_CUDA(c, c->d_res.reserve(rows * 3 * sizeof(int32_t)));
    MSFM_CUDA(c, c->d_m.reserve(rows * 5 * sizeof(int32_t)));
    MSFM_CUDA(c, c->d_exact.reserve(rows * sizeof(int32_t)));
    MSFM_CUDA(c, c->d_counts.reserve(static_cast<size_t>(max_pairs) * sizeof(int32_t)));
    MSFM_CUDA(c, c->d_misc.reserve(16 + nb * 4 * sizeof(unsigned int)));
    TempImgs T{};
    if (opt.cross_check && P > 0) {
        T.n_pad_t = std::max(2 * kUnitRows, (n2_max + 2 * kUnitRows - 1) / (2 * kUnitRows) * (2 * kUnitRows));
        const size_t per = static_cast<size_t>(T.n_pad_t);
        const size_t sw_bytes = static_cast<size_t>(max_pairs) * per * 128;
        const size_t arr_bytes = static_cast<size_t>(max_pairs) * per * 4;
        MSFM_CUDA(c, c->d_temp.reserve(sw_bytes + 3 * arr_bytes + static_cast<size_t>(max_pairs) * 4 + 256));
        uint8_t* base = c->d_temp.as<uint8_t>();
        T.sw = base;
        T.nrm = reinterpret_cast<int32_t*>(base + sw_bytes);
        T.perm = reinterpret_cast<int32_t*>(base + sw_bytes + arr_bytes);
        T.flags = reinterpret_cast<int32_t*>(base + sw_bytes + 2 * arr_bytes);
        T.used = reinterpret_cast<int32_t*>(base + sw_bytes + 3 * arr_bytes);
    }
    if (!segs.empty())
        MSFM_CUDA(c, cudaMemcpy(c->d_segs.p, segs.data(), segs.size() * sizeof(SegDev), cudaMemcpyHostToDevice));
    MSFM_CUDA(c, cudaMemsetAsync(c->d_misc.p, 0, 16 + nb * 4 * sizeof(unsigned int), c->stream));

    long long* running_total = c->d_misc.as<long long>();
    unsigned int* counters = reinterpret_cast<unsigned int*>(c->d_misc.as<uint8_t>() + 16);
    ImgDev* d_imgs = c->d_imgs.as<ImgDev>();
    int32_t* res_j = c->d_res.as<int32_t>();
    int32_t* res_d1 = res_j + rows;
    int32_t* res_u = res_d1 + rows;
    int32_t* m_j = c->d_m.as<int32_t>();
    int32_t* m_d1 = m_j + rows;
    int32_t* m_d2 = m_d1 + rows;
    int32_t* m_j0 = m_d2 + rows;
    if (P == 0) MSFM_CUDA(c, cudaMemsetAsync(out_offsets_dev, 0, sizeof(long long), c->stream));

    int64_t total_units = 0, total_rows = 0;
    for (size_t bi = 0; bi < batches.size(); ++bi) {
        const Batch& b = batches[bi];
        const SegDev* bsegs = c->d_segs.as<SegDev>() + static_cast<size_t>(b.first_pair) * spp;
        UnitDev* units = c->d_units.as<UnitDev>();
        const int nunits = b.units_fwd + b.units_rev;
        unsigned int* bcnt = counters + 4 * bi;         // [0,1] forward pass: exact rows, rescans; [2,3] reverse pass
        if (opt.cross_check) {
            MSFM_CUDA(c, launch_setup_temp_imgs(d_imgs, first_temp_slot, bsegs, b.npairs, T, c->stream));
            // reverse rows that are never computed (train rows nobody matched) must read as "no match"
            MSFM_CUDA(c, cudaMemsetAsync(m_j, 0xFF, static_cast<size_t>(nunits) * kUnitRows * sizeof(int32_t), c->stream));
        }
        c->prof_begin(MSFM_PROF_BUILD_UNITS);
        MSFM_CUDA(c, launch_build_units(bsegs, b.npairs * spp, nunits, units, c->stream));
        c->prof_end();
        if (mode == 0) {
            for (int pass = 0; pass < (opt.cross_check ? 2 : 1); ++pass) {
                const int u0 = pass == 0 ? 0 : b.units_fwd;
                const int nu = pass == 0 ? b.units_fwd : b.units_rev;
                if (pass == 1) {
                    c->prof_begin(MSFM_PROF_COMPACT);
                    MSFM_CUDA(c, launch_gather_candidates(d_imgs, bsegs, b.npairs, m_j, T, c->stream));
                    c->prof_end();
                }
                c->prof_begin(MSFM_PROF_MATCH_TILE);
                MSFM_CUDA(c, launch_match_tile_kernel(d_imgs, units, u0, nu, res_j, res_d1, res_u, c->d_items.p, c->num_sms, c->stream));
                c->prof_end();
                c->prof_begin(MSFM_PROF_RESOLVE);
                MSFM_CUDA(c, launch_resolve_rows(d_imgs, units, u0, nu, res_j, res_d1, res_u, opt, m_j, m_d1, m_d2, m_j0,
                                                 c->d_exact.as<int32_t>(), bcnt + 2 * pass, c->stream));
                c->prof_end();
                c->prof_begin(MSFM_PROF_EXACT);
                MSFM_CUDA(c, launch_exact_rows(d_imgs, units, c->d_exact.as<int32_t>(), bcnt + 2 * pass, -1, opt, m_j, m_d1,
                                               m_d2, m_j0, c->num_sms, c->stream));
                c->prof_end();
            }
        } else {
            // exact CUDA-core scan of every row (knn2 parity API; never combined with cross-check)
            c->prof_begin(MSFM_PROF_EXACT);
            MSFM_CUDA(c, launch_exact_rows(d_imgs, units, nullptr, nullptr, nunits * kUnitRows, opt, m_j, m_d1, m_d2, m_j0,
                                           c->num_sms, c->stream));
            c->prof_end();
        }
        c->prof_begin(MSFM_PROF_COMPACT);
        MSFM_CUDA(c, launch_count_scan_write(d_imgs, bsegs, b.npairs, opt, m_j, m_d1, c->d_counts.as<int32_t>(),
                                             out_offsets_dev + b.first_pair, running_total, capacity, out_matches_dev,
                                             out_dist_dev, c->stream));
        c->prof_end();
        total_units += nunits;
        total_rows += static_cast<int64_t>(nunits) * kUnitRows;
        if (dump && bi == 0 && dump->n > 0) {
            MSFM_CUDA(c, cudaMemcpyAsync(dump->j0, m_j0, static_cast<size_t>(dump->n) * 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
            MSFM_CUDA(c, cudaMemcpyAsync(dump->d1, m_d1, static_cast<size_t>(dump->n) * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
            MSFM_CUDA(c, cudaMemcpyAsync(dump->d2, m_d2, static_cast<size_t>(dump->n) * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
        }
    }
    // ---- totals (one small readback; this is the call's only host<->device sync besides the entry one)
    MSFM_CUDA(c, c->h_stage.reserve(16 + nb * 4 * sizeof(unsigned int)));
    MSFM_CUDA(c, cudaMemcpyAsync(c->h_stage.p, c->d_misc.p, 16 + nb * 4 * sizeof(unsigned int), cudaMemcpyDeviceToHost, c->stream));
    MSFM_CUDA(c, cudaStreamSynchronize(c->stream));
    const long long total = *c->h_stage.as<long long>();
    const unsigned int* hc = reinterpret_cast<const unsigned int*>(c->h_stage.as<uint8_t>() + 16);
    int64_t n_exact = 0, n_rescan = 0;
    for (size_t bi = 0; bi < batches.size(); ++bi) { n_exact += hc[4 * bi] + hc[4 * bi + 2]; n_rescan += hc[4 * bi + 1] + hc[4 * bi + 3]; }
    c->stats[0] = total_rows; c->stats[1] = n_rescan; c->stats[2] = n_exact; c->stats[3] = total_units;
    if (total_out) *total_out = total;
    if (total > capacity) return c->fail(MSFM_E_CAPACITY, "output capacity %lld < %lld matches", capacity, total);
    return MSFM_OK;
}

int msfm_match_pairs_dev(msfm_ctx* c, const int32_t* pairs_host, int32_t P, const msfm_match_options* opt,
                         int64_t* out_offsets_dev, int32_t* out_matches_dev, float* out_dist_dev, int64_t capacity,
                         int64_t* total_out) {
    if (!c) return MSFM_E_INVALID;
    if (!out_offsets_dev || capacity < 0 || (capacity > 0 && !out_matches_dev))
        return c->fail(MSFM_E_INVALID, "msfm_match_pairs_dev: bad output arguments");
    long long total = 0;
    int rc = match_core(c, pairs_host, P, opt, 0, 0, reinterpret_cast<long long*>(out_offsets_dev), out_matches_dev,
                        out_dist_dev, capacity, &total, nullptr);
    if (total_out) *total_out = total;
    return rc;
}

int msfm_match_pairs(msfm_ctx* c, const int32_t* pairs, int32_t P, const msfm_match_options* opt, int64_t* out_offsets,
                     int32_t* out_matches, float* out_dist, int64_t capacity, int64_t* total_out) {
    if (!c) return MSFM_E_INVALID;
    if (!out_offsets || capacity < 0 || (capacity > 0 && !out_matches) || P < 0 || (P > 0 && !pairs))
        return c->fail(MSFM_E_INVALID, "msfm_match_pairs: bad arguments");
    MSFM_CUDA(c, cudaSetDevice(c->device));
    // device capacity: never more than one match per query row
    long long bound = 0;
    for (int p = 0; p < P; ++p) {
        auto it = c->slot_of.find(pairs[2 * p]);
        if (it == c->slot_of.end()) return c->fail(MSFM_E_NOT_FOUND, "pair %d: image %d not resident", p, pairs[2 * p]);
        bound += c->imgs[it->second].n;
    }
    const long long devcap = std::min<long long>(capacity, bound);
    MSFM_CUDA(c, cudaStreamSynchronize(c->stream));
    MSFM_CUDA(c, c->d_out_offsets.reserve((static_cast<size_t>(P) + 1) * sizeof(long long)));
    MSFM_CUDA(c, c->d_out_matches.reserve(std::max<size_t>(1, static_cast<size_t>(devcap)) * 2 * sizeof(int32_t)));
    if (out_dist) MSFM_CUDA(c, c->d_out_dist.reserve(std::max<size_t>(1, static_cast<size_t>(devcap)) * sizeof(float)));
    long long total = 0;
    int rc = match_core(c, pairs, P, opt, 0, 0, c->d_out_offsets.as<long long>(), c->d_out_matches.as<int32_t>(),
                        out_dist ? c->d_out_dist.as<float>() : nullptr, devcap, &total, nullptr);
    if (total_out) *total_out = total;
    if (rc != MSFM_OK && rc != MSFM_E_CAPACITY) return rc;
    MSFM_CUDA(c, cudaMemcpyAsync(out_offsets, c->d_out_offsets.p, (static_cast<size_t>(P) + 1) * sizeof(long long),
                                 cudaMemcpyDeviceToHost, c->stream));
    if (rc == MSFM_OK && total > 0) {
        MSFM_CUDA(c, cudaMemcpyAsync(out_matches, c->d_out_matches.p, static_cast<size_t>(total) * 2 * sizeof(int32_t),
                                     cudaMemcpyDeviceToHost, c->stream));
        if (out_dist)
            MSFM_CUDA(c, cudaMemcpyAsync(out_dist, c->d_out_dist.p, static_cast<size_t>(total) * sizeof(float),
                                         cudaMemcpyDeviceToHost, c->stream));
    }
    MSFM_CUDA(c, cudaStreamSynchronize(c->stream));
    return rc;
}

int msfm_match_knn2_u8(msfm_ctx* c, const uint8_t* A, int32_t nA, const uint8_t* B, int32_t nB, int32_t mode,
                       int32_t* idx, float* dist, int32_t* d2) {
    if (!c) return MSFM_E_INVALID;
    if (nA < 0 || nB < 0 || (nA > 0 && (!A || !idx || !dist)) || (nB > 0 && !B) || (mode != 0 && mode != 1))
        return c->fail(MSFM_E_INVALID, "msfm_match_knn2_u8: bad arguments");
    if (nA == 0) return MSFM_OK;
    int rc;
    if ((rc = upload_common(c, kTmpIdA, A, nA, false))) return rc;
    MSFM_CUDA(c, cudaStreamSynchronize(c->stream));       // d_raw staging is reused by the next upload
    if ((rc = upload_common(c, kTmpIdB, B, nB, false))) return rc;
    msfm_match_options o;
    o.max_distance = -1.0; o.distance_ratio = 0.8f; o.cross_check = 0; o.opencv_quirks = 0; o.reserved = 0;
    const int32_t pair[2] = {kTmpIdA, kTmpIdB};
    std::vector<int32_t> hj(static_cast<size_t>(nA) * 2), hd1(nA), hd2(nA);
    RowDump dump; dump.j0 = hj.data(); dump.d1 = hd1.data(); dump.d2 = hd2.data(); dump.n = nA;
    MSFM_CUDA(c, c->d_out_offsets.reserve(2 * sizeof(long long)));
    MSFM_CUDA(c, c->d_out_matches.reserve(std::max<size_t>(1, nA) * 2 * sizeof(int32_t)));
    long long total = 0;
    rc = match_core(c, pair, 1, &o, 1, mode, c->d_out_offsets.as<long long>(), c->d_out_matches.as<int32_t>(), nullptr, nA,
                    &total, &dump);
    int rc2 = msfm_desc_release(c, kTmpIdA);
    int rc3 = msfm_desc_release(c, kTmpIdB);
    if (rc) return rc;
    if (rc2) return rc2;
    if (rc3) return rc3;
    for (int i = 0; i < nA; ++i) {
        const int32_t a = hd1[i], b = hd2[i];
        const bool h0 = hj[2 * i] >= 0 && a != kIntInf, h1 = h0 && b != kIntInf;
        idx[2 * i] = h0 ? hj[2 * i] : -1;
        idx[2 * i + 1] = h1 ? hj[2 * i + 1] : -1;
        dist[2 * i] = h0 ? sqrtf(static_cast<float>(a)) : INFINITY;
        dist[2 * i + 1] = h1 ? sqrtf(static_cast<float>(b)) : INFINITY;
        if (d2) { d2[2 * i] = h0 ? a : -1; d2[2 * i + 1] = h1 ? b : -1; }
    }
    return MSFM_OK;
}

int msfm_match_stats(msfm_ctx* c, int64_t stats[4]) {
    if (!c || !stats) return MSFM_E_INVALID;
    for (int i = 0; i < 4; ++i) stats[i] = c->stats[i];
    return MSFM_OK;
}

}  // extern "C"
