// msfm_ctx — host-side state of one (process, device): stream, resident descriptor sets, scratch.
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/msfm_b200.h"
#include "match_types.cuh"

namespace msfm {

// A device (or pinned-host) buffer that only ever grows.
struct GrowBuf {
    void* p = nullptr;
    size_t cap = 0;
    bool pinned_host = false;
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) {
            cudaError_t e = pinned_host ? cudaFreeHost(p) : cudaFree(p);
            p = nullptr; cap = 0;
            if (e != cudaSuccess) return e;
        }
        size_t want = bytes + bytes / 4 + 256;
        cudaError_t e = pinned_host ? cudaMallocHost(&p, want) : cudaMalloc(&p, want);
        if (e != cudaSuccess) { p = nullptr; return e; }
        cap = want;
        return cudaSuccess;
    }
    void release() {
        if (p) { if (pinned_host) cudaFreeHost(p); else cudaFree(p); }
        p = nullptr; cap = 0;
    }
    template <class T> T* as() const { return static_cast<T*>(p); }
};

struct ImgHost {
    void* block = nullptr;     // one allocation: sw [n_pad*128] then cj [n_pad]
    int32_t n = 0, n_pad = 0;
    bool live = false;
};

}  // namespace msfm

struct msfm_ctx {
    int device = 0;
    int num_sms = 0;
    cudaStream_t stream = nullptr;
    std::string err;
    int64_t launches = 0;          // value of the library's launch counter (launch_count.hpp) when the ctx was created

    // ---- M-path
    std::unordered_map<int32_t, int> slot_of;      // image_id -> slot
    std::vector<msfm::ImgHost> imgs;               // slot -> allocation
    std::vector<int> free_slots;
    msfm::GrowBuf d_imgs;                          // ImgDev[slots]
    bool imgs_dirty = true;
    msfm::GrowBuf d_raw;                           // upload staging (device): float32 uploads, knn2
    // uint8 host uploads: two staging buffers filled on a copy stream, so that the H2D copy of image k+1 runs under the
    // formatting kernels of image k
    cudaStream_t copy_stream = nullptr;
    msfm::GrowBuf d_rawq[2];
    cudaEvent_t raw_ready[2] = {nullptr, nullptr}, raw_free[2] = {nullptr, nullptr};
    bool raw_free_set[2] = {false, false};
    uint32_t raw_turn = 0;
    msfm::GrowBuf d_fmt;                           // upload formatting scratch (sort keys, ranks)
    msfm::GrowBuf d_temp;                          // temporary query images of the reverse (cross-check) pass
    msfm::GrowBuf h_stage;                         // pinned host staging (segments, offsets readback)
    msfm::GrowBuf d_segs, d_units, d_items, d_res, d_m, d_exact, d_counts, d_misc;
    msfm::GrowBuf d_out_offsets, d_out_matches, d_out_dist;
    msfm::GrowBuf d_ba_r, d_ba_J;                  // msfm_ba_evaluate parity dumps
    // ---- geometric verification: resident keypoint positions per image, call tables, staged match lists / masks
    struct KpHost { void* xy = nullptr; int32_t n = 0; };
    std::unordered_map<int32_t, KpHost> kps;
    msfm::GrowBuf d_kp_tab, d_kp_slots, d_vf_in, d_vf_out;
    int64_t stats[4] = {0, 0, 0, 0};

    // ---- B-path / multi-GPU
    void* cusolver = nullptr;      // cusolverDnHandle_t (reduced camera system Cholesky)
    void* comm = nullptr;          // ncclComm_t
    int comm_ranks = 1, comm_rank = 0;

    // ---- optional per-kernel-class timing with CUDA events on the ctx stream (bench.py roofline)
    struct ProfRec { int cat; cudaEvent_t a, b; };
    bool prof_on = false;
    std::vector<cudaEvent_t> prof_pool;
    std::vector<ProfRec> prof_recs;
    double prof_ms[MSFM_PROF_NCAT] = {0};
    int64_t prof_n[MSFM_PROF_NCAT] = {0};
    cudaEvent_t prof_event() {
        if (!prof_pool.empty()) { cudaEvent_t e = prof_pool.back(); prof_pool.pop_back(); return e; }
        cudaEvent_t e = nullptr;
        cudaEventCreate(&e);
        return e;
    }
    void prof_begin(int cat) {
        if (!prof_on) return;
        ProfRec r{cat, prof_event(), prof_event()};
        cudaEventRecord(r.a, stream);
        prof_recs.push_back(r);
    }
    void prof_end() {
        if (!prof_on || prof_recs.empty()) return;
        cudaEventRecord(prof_recs.back().b, stream);
    }
    void prof_collect() {      // stream must be idle
        for (ProfRec& r : prof_recs) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) { prof_ms[r.cat] += ms; prof_n[r.cat] += 1; }
            prof_pool.push_back(r.a);
            prof_pool.push_back(r.b);
        }
        prof_recs.clear();
    }

    int fail(int code, const char* fmt, ...) {
        char buf[512];
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(buf, sizeof buf, fmt, ap);
        va_end(ap);
        err = buf;
        return code;
    }
    int cuda_fail(cudaError_t e, const char* what) {
        return fail(MSFM_E_CUDA, "%s: %s", what, cudaGetErrorString(e));
    }
};

#define MSFM_CUDA(ctx, call)                                            \
    do {                                                                \
        cudaError_t e__ = (call);                                       \
        if (e__ != cudaSuccess) return (ctx)->cuda_fail(e__, #call);    \
    } while (0)
