// libdsa K1 — device radix sort of the pending batch.
//
// LSD radix sort of (uint64 key, uint32 payload) pairs over the bit range [0, nbits), 8 bits per
// pass, stable.  Each pass: (1) per-tile digit histogram, (2) exclusive scan of the (digit-major,
// tile-minor) count table, (3) stable scatter: every warp ranks its keys with __match_any_sync
// in tile order, so equal digits keep their input order.  Tiles are 256 threads x 16 keys.
#pragma once
#include "common.cuh"
#include "primitives.cuh"

namespace dsa {

constexpr int RS_THREADS = 256;
constexpr int RS_ITEMS = 16;
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;   // 4096 keys per tile
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_RADIX = 256;

// counts[digit * ntiles + tile]
__global__ void __launch_bounds__(RS_THREADS) k_rs_histogram(const uint64_t* __restrict__ keys, int64_t n, int shift,
                                                              int32_t* __restrict__ counts, int64_t ntiles) {
    __shared__ int32_t h[RS_RADIX];
    h[threadIdx.x] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * RS_TILE;
#pragma unroll
    for (int i = 0; i < RS_ITEMS; ++i) {
        int64_t idx = base + (int64_t)i * RS_THREADS + threadIdx.x;
        if (idx < n) atomicAdd(&h[(unsigned)((keys[idx] >> shift) & 0xff)], 1);
    }
    __syncthreads();
    counts[(int64_t)threadIdx.x * ntiles + blockIdx.x] = h[threadIdx.x];
}

// offsets = exclusive scan of counts (digit-major).  Warp w owns keys [w*512, (w+1)*512) of the tile, processed in
// 16 rounds of 32 consecutive keys, so (warp, round, lane) order == input order.
__global__ void __launch_bounds__(RS_THREADS) k_rs_scatter(const uint64_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                                                            uint64_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out, int64_t n,
                                                            int shift, const int32_t* __restrict__ offsets, int64_t ntiles) {
    __shared__ int32_t wcnt[RS_WARPS][RS_RADIX];   // per-warp digit counts, then exclusive per-warp bases
    __shared__ int32_t dbase[RS_RADIX];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int d = threadIdx.x; d < RS_WARPS * RS_RADIX; d += RS_THREADS) (&wcnt[0][0])[d] = 0;
    dbase[threadIdx.x] = offsets[(int64_t)threadIdx.x * ntiles + blockIdx.x];
    __syncthreads();
    const int64_t wbase = (int64_t)blockIdx.x * RS_TILE + (int64_t)wid * (RS_ITEMS * 32);
    uint64_t k[RS_ITEMS];
    int32_t rank[RS_ITEMS];
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {
        int64_t idx = wbase + r * 32 + lane;
        bool valid = idx < n;
        k[r] = valid ? keys_in[idx] : ~uint64_t(0);
        unsigned d = (unsigned)((k[r] >> shift) & 0xff);
        // invalid lanes use a digit id outside the radix so they never match valid ones
        unsigned dm = valid ? d : 0x100u;
        unsigned peers = __match_any_sync(0xffffffffu, dm);
        int before = __popc(peers & ((1u << lane) - 1));
        int32_t basecnt = 0;
        if (valid) basecnt = wcnt[wid][d];
        __syncwarp();
        if (valid && before == 0) wcnt[wid][d] = basecnt + __popc(peers);
        __syncwarp();
        rank[r] = basecnt + before;
    }
    __syncthreads();
    // exclusive scan over warps per digit (thread d handles digit d)
    {
        int d = threadIdx.x;
        int32_t run = 0;
#pragma unroll
        for (int w = 0; w < RS_WARPS; ++w) {
            int32_t c = wcnt[w][d];
            wcnt[w][d] = run;
            run += c;
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {
        int64_t idx = wbase + r * 32 + lane;
        if (idx < n) {
            unsigned d = (unsigned)((k[r] >> shift) & 0xff);
            int64_t dst = (int64_t)dbase[d] + wcnt[wid][d] + rank[r];
            keys_out[dst] = k[r];
            vals_out[dst] = vals_in[idx];
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Single-sweep variant (used for n < 2^30): the digit histograms of ALL passes are computed by one read of the keys;
// each pass is then ONE kernel in which a tile publishes its digit counts and resolves its global offsets by decoupled
// look-back over the preceding tiles (tiles take their index from an atomic ticket, so a tile only ever waits on tiles
// that are already running).  6 launches for a 34-bit key instead of 25.
// ---------------------------------------------------------------------------------------------
constexpr int OS_THREADS = 256;
constexpr int OS_ROUNDS = 16;                        // keys per thread
constexpr int OS_TILE = OS_THREADS * OS_ROUNDS;      // 4096 keys per tile: a 1M-key batch is ONE wave of 245 CTAs (2 per SM)
constexpr int OS_WARPS = OS_THREADS / 32;
constexpr int OS_MAX_PASS = 8;
constexpr int OS_LOOK = 8;
constexpr uint32_t OS_FLAG_LOCAL = 1u << 30, OS_FLAG_INCL = 2u << 30, OS_VALUE_MASK = (1u << 30) - 1u;

// ghist[pass * 256 + digit] += count
__global__ void __launch_bounds__(256) k_os_histogram(const uint64_t* __restrict__ keys, int64_t n, int npass, uint32_t* __restrict__ ghist) {
    __shared__ uint32_t h[OS_MAX_PASS * RS_RADIX];
    for (int i = threadIdx.x; i < npass * RS_RADIX; i += blockDim.x) h[i] = 0;
    __syncthreads();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t k = keys[i];
        for (int p = 0; p < npass; ++p) atomicAdd(&h[p * RS_RADIX + (unsigned)((k >> (8 * p)) & 0xff)], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < npass * RS_RADIX; i += blockDim.x)
        if (h[i]) atomicAdd(&ghist[i], h[i]);
}
// exclusive scan of each pass's 256 bins, in place (one block of 256 threads)
__global__ void __launch_bounds__(256) k_os_scan_hist(uint32_t* __restrict__ ghist, int npass) {
    __shared__ uint32_t s[RS_RADIX];
    for (int p = 0; p < npass; ++p) {
        const uint32_t v = ghist[p * RS_RADIX + threadIdx.x];
        s[threadIdx.x] = v;
        __syncthreads();
        for (int o = 1; o < RS_RADIX; o <<= 1) {
            uint32_t t = threadIdx.x >= (unsigned)o ? s[threadIdx.x - o] : 0;
            __syncthreads();
            s[threadIdx.x] += t;
            __syncthreads();
        }
        ghist[p * RS_RADIX + threadIdx.x] = s[threadIdx.x] - v;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(OS_THREADS, 2) k_os_pass(const uint64_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                                                         uint64_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out, int64_t n,
                                                         int shift, const uint32_t* __restrict__ gbase, volatile uint32_t* status,
                                                         unsigned* __restrict__ ticket) {
    __shared__ int32_t wcnt[OS_WARPS][RS_RADIX];
    __shared__ uint32_t dbase[RS_RADIX];
    __shared__ unsigned tile_s;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) tile_s = atomicAdd(ticket, 1u);
    for (int d = threadIdx.x; d < OS_WARPS * RS_RADIX; d += OS_THREADS) (&wcnt[0][0])[d] = 0;
    __syncthreads();
    const unsigned tile = tile_s;
    const int64_t wbase = (int64_t)tile * OS_TILE + (int64_t)wid * (OS_ROUNDS * 32);
    uint64_t k[OS_ROUNDS];
    uint32_t v[OS_ROUNDS];
    int32_t rank[OS_ROUNDS];
    unsigned peers[OS_ROUNDS];
#pragma unroll
    for (int r = 0; r < OS_ROUNDS; ++r) {   // all loads in flight together
        const int64_t idx = wbase + r * 32 + lane;
        k[r] = idx < n ? keys_in[idx] : ~uint64_t(0);
        v[r] = idx < n ? vals_in[idx] : 0u;
    }
#pragma unroll
    for (int r = 0; r < OS_ROUNDS; ++r) {   // independent matches pipeline through the MIO queue
        const int64_t idx = wbase + r * 32 + lane;
        const unsigned d = (unsigned)((k[r] >> shift) & 0xff);
        peers[r] = __match_any_sync(0xffffffffu, idx < n ? d : 0x100u);
    }
#pragma unroll
    for (int r = 0; r < OS_ROUNDS; ++r) {   // serial part: running per-warp digit counters
        const int64_t idx = wbase + r * 32 + lane;
        const bool valid = idx < n;
        const unsigned d = (unsigned)((k[r] >> shift) & 0xff);
        const int before = __popc(peers[r] & ((1u << lane) - 1u));
        int32_t basecnt = 0;
        if (valid) basecnt = wcnt[wid][d];
        __syncwarp();
        if (valid && before == 0) wcnt[wid][d] = basecnt + __popc(peers[r]);
        __syncwarp();
        rank[r] = basecnt + before;
    }
    __syncthreads();
    {   // thread d owns digit d: tile count, per-warp exclusive offsets, publish, look back
        const int d = threadIdx.x;
        uint32_t run = 0;
#pragma unroll
        for (int w = 0; w < OS_WARPS; ++w) {
            const int32_t c = wcnt[w][d];
            wcnt[w][d] = (int32_t)run;
            run += (uint32_t)c;
        }
        volatile uint32_t* mine = status + (size_t)tile * RS_RADIX + d;
        uint32_t excl = 0;
        if (tile == 0) {
            *mine = OS_FLAG_INCL | run;
        } else {
            *mine = OS_FLAG_LOCAL | run;
            // decoupled look-back, OS_LOOK predecessors per round trip (all tiles of a small sort run concurrently, so the
            // nearest inclusive prefix can be tens of tiles away: batching the polls keeps the chain short)
            int64_t t = (int64_t)tile - 1;
            bool done = false;
            while (!done) {
                uint32_t sv[OS_LOOK];
#pragma unroll
                for (int b = 0; b < OS_LOOK; ++b) sv[b] = (t - b >= 0) ? status[(size_t)(t - b) * RS_RADIX + d] : (uint32_t)(2u << 30);
#pragma unroll
                for (int b = 0; b < OS_LOOK; ++b) {
                    if (done) break;
                    const uint32_t f = sv[b] >> 30;
                    if (f == 0) break;          // not published yet: poll again from this tile
                    excl += sv[b] & OS_VALUE_MASK;
                    --t;
                    if (f == 2u) done = true;
                }
            }
            *mine = OS_FLAG_INCL | (excl + run);
        }
        dbase[d] = gbase[d] + excl;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < OS_ROUNDS; ++r) {
        const int64_t idx = wbase + r * 32 + lane;
        if (idx < n) {
            const unsigned d = (unsigned)((k[r] >> shift) & 0xff);
            const int64_t dst = (int64_t)dbase[d] + wcnt[wid][d] + rank[r];
            keys_out[dst] = k[r];
            vals_out[dst] = v[r];
        }
    }
}

struct SortWorkspace {
    DBuf<int32_t> counts;
    DBuf<uint64_t> keys_alt;
    DBuf<uint32_t> vals_alt;
    DBuf<uint32_t> os_state;   // [ghist: npass*256][tickets: npass][status: npass*ntiles*256]
    ScanWorkspace scan;
};

// Sorts in place (result ends in d_keys / d_vals). nbits = number of significant low bits of the keys.
inline void radix_sort_pairs(SortWorkspace& ws, uint64_t* d_keys, uint32_t* d_vals, int64_t n, int nbits, cudaStream_t st) {
    if (n <= 1 || nbits <= 0) return;
    if (n < (int64_t(1) << 30) && (nbits + 7) / 8 <= OS_MAX_PASS) {
        const int npass = (nbits + 7) / 8;
        const int64_t ntiles = (n + OS_TILE - 1) / OS_TILE;
        const size_t nstate = (size_t)npass * RS_RADIX + 64 + (size_t)npass * ntiles * RS_RADIX;
        uint32_t* state = ws.os_state.ensure(nstate);
        DSA_CUDA(cudaMemsetAsync(state, 0, nstate * sizeof(uint32_t), st));
        uint32_t* ghist = state;
        unsigned* tickets = state + (size_t)npass * RS_RADIX;
        uint32_t* status = state + (size_t)npass * RS_RADIX + 64;
        uint64_t* ka = ws.keys_alt.ensure((size_t)n);
        uint32_t* va = ws.vals_alt.ensure((size_t)n);
        const unsigned hgrid = (unsigned)std::min<int64_t>((n + 255) / 256, 148 * 4);
        DSA_LAUNCH("os_histogram", k_os_histogram, hgrid, 256, 0, st, d_keys, n, npass, ghist);
        DSA_LAUNCH("os_scan_hist", k_os_scan_hist, 1, 256, 0, st, ghist, npass);
        uint64_t *kin = d_keys, *kout = ka;
        uint32_t *vin = d_vals, *vout = va;
        for (int pass = 0; pass < npass; ++pass) {
            DSA_LAUNCH("os_pass", k_os_pass, (unsigned)ntiles, OS_THREADS, 0, st, kin, vin, kout, vout, n, pass * 8, ghist + pass * RS_RADIX,
                       status + (size_t)pass * ntiles * RS_RADIX, tickets + pass);
            std::swap(kin, kout);
            std::swap(vin, vout);
        }
        if (kin != d_keys) {
            DSA_CUDA(cudaMemcpyAsync(d_keys, kin, (size_t)n * sizeof(uint64_t), cudaMemcpyDeviceToDevice, st));
            DSA_CUDA(cudaMemcpyAsync(d_vals, vin, (size_t)n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
        }
        return;
    }
    const int64_t ntiles = (n + RS_TILE - 1) / RS_TILE;
    int32_t* counts = ws.counts.ensure((size_t)(ntiles * RS_RADIX));
    uint64_t* ka = ws.keys_alt.ensure((size_t)n);
    uint32_t* va = ws.vals_alt.ensure((size_t)n);
    uint64_t* kin = d_keys;
    uint32_t* vin = d_vals;
    uint64_t* kout = ka;
    uint32_t* vout = va;
    const int npass = (nbits + 7) / 8;
    for (int pass = 0; pass < npass; ++pass) {
        const int shift = pass * 8;
        DSA_LAUNCH("rs_histogram", k_rs_histogram, (unsigned)ntiles, RS_THREADS, 0, st, kin, n, shift, counts, ntiles);
        exclusive_scan_i32<int32_t>(ws.scan, counts, counts, ntiles * RS_RADIX, nullptr, st);
        DSA_LAUNCH("rs_scatter", k_rs_scatter, (unsigned)ntiles, RS_THREADS, 0, st, kin, vin, kout, vout, n, shift, counts, ntiles);
        std::swap(kin, kout);
        std::swap(vin, vout);
    }
    if (kin != d_keys) {
        DSA_CUDA(cudaMemcpyAsync(d_keys, kin, (size_t)n * sizeof(uint64_t), cudaMemcpyDeviceToDevice, st));
        DSA_CUDA(cudaMemcpyAsync(d_vals, vin, (size_t)n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
    }
}

}  // namespace dsa
