// libdsa — shared definitions: error type, device buffers, launch accounting.
#pragma once
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <ctime>
#include <map>
#include <mutex>
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

// ---- NVTX range per C-ABI call (header-only NVTX v3: a no-op unless a tool is attached) ---------------
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

// ---- launch accounting + optional per-kernel event timing ----------------------------------
struct ProfEntry {
    int64_t count = 0;
    double ms = 0.0;
};
// Per-kernel timing (dsa_prof_enable): every launch is bracketed by two CUDA events on its stream and waited for, so the
// kernels of a step run ONE AT A TIME (no overlap of the two orientations' streams): the durations are those of each kernel
// alone, comparable with an ncu launch list; their sum is larger than the step.
struct Prof {
    bool enabled = false;
    int64_t launches = 0;
    std::map<std::string, ProfEntry> entries;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    void resolve() {}
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

// host-timed allocator calls show up in the profile dump as "(cudaMalloc)" / "(cudaFree)"
struct AllocTimer {
    const char* name;
    double t0;
    static double now() {
        timespec ts;
        clock_gettime(CLOCK_MONOTONIC, &ts);
        return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
    }
    explicit AllocTimer(const char* n) : name(n), t0(prof().enabled ? now() : 0.0) {}
    ~AllocTimer() {
        if (prof().enabled) {
            ProfEntry& e = prof().entries[name];
            e.count += 1;
            e.ms += now() - t0;
        }
    }
};

// ---- caching device allocator ---------------------------------------------------------------------
// cudaMalloc / cudaFree of the 100 MB-class buffers of a growing structure cost milliseconds each (measured: 230 ms per batch
// while a matrix doubles, profiles/prof_c5.py).  Blocks are therefore recycled: sizes are rounded to 4 classes per octave and a
// released block is handed out again to the next request of a compatible size.  One pool PER DEVICE (a process may drive
// several GPUs through dsa_set_device).  Releasing does not synchronise: a released block waits in `pending` and becomes
// reusable at the first device synchronisation that an allocation actually needs (or that dsa_trim_memory performs) — the
// guarantee cudaFree gives (nobody is still using the block), paid only when a block is really recycled.
// dsa_trim_memory() returns the cached blocks to the driver.
struct DevicePool {
    std::multimap<size_t, void*> free_blocks, pending;
    size_t cached_bytes = 0;
    std::mutex mu;   // handles of different host threads share the pool
    static size_t size_class(size_t bytes) {
        if (bytes < 4096) return 4096;
        int lg = 63 - __builtin_clzll((unsigned long long)bytes);
        const size_t step = size_t(1) << (lg - 2);
        return (bytes + step - 1) / step * step;
    }
    void* take_locked(size_t want) {
        auto it = free_blocks.lower_bound(want);
        if (it != free_blocks.end() && it->first <= want + want / 2) {
            void* p = it->second;
            cached_bytes -= it->first;
            free_blocks.erase(it);
            return p;
        }
        return nullptr;
    }
    void settle_locked() {   // everything released so far is idle after this
        if (pending.empty()) return;
        cudaDeviceSynchronize();
        for (auto& kv : pending) free_blocks.emplace(kv.first, kv.second);
        pending.clear();
    }
    void* get(size_t bytes) {
        std::lock_guard<std::mutex> lock(mu);
        const size_t want = size_class(bytes);
        if (void* p = take_locked(want)) return p;
        auto it = pending.lower_bound(want);
        if (it != pending.end() && it->first <= want + want / 2) {   // a released block fits: wait for its last users once
            settle_locked();
            if (void* p = take_locked(want)) return p;
        }
        void* p = nullptr;
        AllocTimer t("(cudaMalloc)");
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {   // give the cache back to the driver and retry once
            cudaGetLastError();
            trim_locked();
            e = cudaMalloc(&p, want);
        }
        if (e != cudaSuccess) {
            cudaGetLastError();
            throw DsaError{DSA_ERR_OOM, "cudaMalloc(" + std::to_string(want) + " bytes): " + cudaGetErrorString(e)};
        }
        return p;
    }
    void put(void* p, size_t bytes) {
        std::lock_guard<std::mutex> lock(mu);
        const size_t c = size_class(bytes);
        pending.emplace(c, p);
        cached_bytes += c;
    }
    void trim_locked() {
        settle_locked();
        AllocTimer t("(cudaFree)");
        for (auto& kv : free_blocks) cudaFree(kv.second);
        free_blocks.clear();
        cached_bytes = 0;
    }
    void trim() {
        std::lock_guard<std::mutex> lock(mu);
        trim_locked();
    }
};
DevicePool& device_pool(int device);   // the pool of a device
inline int current_device() {
    int d = 0;
    cudaGetDevice(&d);
    return d;
}

// ---- device buffer that only grows ------------------------------------------------------------
template <typename T>
struct DBuf {
    T* p = nullptr;
    size_t cap = 0;   // elements
    int device = 0;   // where the block lives: it goes back to that device's pool
    DBuf() {}
    DBuf(const DBuf&) = delete;
    DBuf& operator=(const DBuf&) = delete;
    ~DBuf() { release(); }
    void release() {
        if (p) device_pool(device).put(p, cap * sizeof(T));
        p = nullptr;
        cap = 0;
    }
    // contents are NOT preserved
    T* ensure(size_t n) {
        if (n > cap) {
            release();
            const size_t bytes = DevicePool::size_class((n + 64) * sizeof(T));
            device = current_device();
            p = (T*)device_pool(device).get(bytes);
            cap = bytes / sizeof(T);
        }
        return p;
    }
    void swap(DBuf& o) {
        std::swap(p, o.p);
        std::swap(cap, o.cap);
        std::swap(device, o.device);
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
