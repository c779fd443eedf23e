// libdsa — shared definitions: error type, device buffers, launch accounting.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>
#include "../../include/dsa.h"

namespace dsa {

// Gap sentinel of the SoA layout: keys[p] == GAP_KEY <=> the reference's `array[p] === nothing`
// (DynamicSparseArrays.jl:18).  In-array keys are >= 0 (0 = semaphore key, pcsr.jl:23), so INT64_MIN is free.
constexpr int64_t GAP_KEY = INT64_MIN;
constexpr int MAX_LEVELS = 34;   // height + 1 <= 33 for capacity <= 2^32 with segment capacity >= 1

struct DsaError {
    int code;
    std::string msg;
};

#define DSA_CUDA(expr)                                                                                            \
    do {                                                                                                          \
        cudaError_t _e = (expr);                                                                                  \
        if (_e != cudaSuccess) {                                                                                  \
            int _code = (_e == cudaErrorMemoryAllocation) ? DSA_ERR_OOM : DSA_ERR_CUDA;                           \
            throw ::dsa::DsaError{_code, std::string(#expr) + ": " + cudaGetErrorString(_e)};                      \
        }                                                                                                         \
    } while (0)

// ---- launch accounting + optional per-kernel event timing ----------------------------------
struct ProfEntry {
    int64_t count = 0;
    double ms = 0.0;
};
struct Prof {
    bool enabled = false;
    int64_t launches = 0;
    std::map<std::string, ProfEntry> entries;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
};
Prof& prof();

struct LaunchScope {   // brackets ONE kernel launch
    const char* name;
    cudaStream_t st;
    LaunchScope(const char* n, cudaStream_t s) : name(n), st(s) {
        Prof& p = prof();
        p.launches += 1;
        if (p.enabled) {
            if (!p.e0) { cudaEventCreate(&p.e0); cudaEventCreate(&p.e1); }
            cudaEventRecord(p.e0, st);
        }
    }
    ~LaunchScope() {
        Prof& p = prof();
        if (p.enabled) {
            cudaEventRecord(p.e1, st);
            cudaEventSynchronize(p.e1);
            float ms = 0.f;
            cudaEventElapsedTime(&ms, p.e0, p.e1);
            ProfEntry& e = p.entries[name];
            e.count += 1;
            e.ms += ms;
        }
    }
};

#define DSA_LAUNCH(name, kernel, grid, block, smem, stream, ...)                   \
    do {                                                                           \
        ::dsa::LaunchScope _ls(name, stream);                                      \
        kernel<<<grid, block, smem, stream>>>(__VA_ARGS__);                        \
        DSA_CUDA(cudaGetLastError());                                              \
    } while (0)

// ---- device buffer that only grows ------------------------------------------------------------
template <typename T>
struct DBuf {
    T* p = nullptr;
    size_t cap = 0;   // elements
    DBuf() {}
    DBuf(const DBuf&) = delete;
    DBuf& operator=(const DBuf&) = delete;
    ~DBuf() { release(); }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    // contents are NOT preserved
    T* ensure(size_t n) {
        if (n > cap) {
            release();
            size_t want = n + n / 4 + 64;
            DSA_CUDA(cudaMalloc(&p, want * sizeof(T)));
            cap = want;
        }
        return p;
    }
    // contents preserved (copy on stream)
    T* grow_keep(size_t n, size_t keep, cudaStream_t st) {
        if (n > cap) {
            size_t want = n + n / 2 + 64;
            T* q = nullptr;
            DSA_CUDA(cudaMalloc(&q, want * sizeof(T)));
            if (p && keep) DSA_CUDA(cudaMemcpyAsync(q, p, keep * sizeof(T), cudaMemcpyDeviceToDevice, st));
            DSA_CUDA(cudaStreamSynchronize(st));
            if (p) cudaFree(p);
            p = q;
            cap = want;
        }
        return p;
    }
    void swap(DBuf& o) {
        std::swap(p, o.p);
        std::swap(cap, o.cap);
    }
};

// pinned host scratch for small status read-backs
template <typename T>
struct HPinned {
    T* p = nullptr;
    size_t cap = 0;
    ~HPinned() { if (p) cudaFreeHost(p); }
    T* ensure(size_t n) {
        if (n > cap) {
            if (p) cudaFreeHost(p);
            p = nullptr;
            DSA_CUDA(cudaMallocHost(&p, n * sizeof(T)));
            cap = n;
        }
        return p;
    }
};

static inline int ilog2_i64(int64_t x) {
    int r = 0;
    while ((int64_t(1) << r) < x) ++r;
    return r;
}
static inline unsigned grid_for(int64_t n, int block) { return (unsigned)((n + block - 1) / block); }

}  // namespace dsa
