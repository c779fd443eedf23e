// libdsa — column-range sharding of a DynamicSparseMatrix over the GPUs of one box, one process per GPU.
//
// Rank r owns the column-major PCSR of the columns in [col_split[r], col_split[r+1]) and the row-major PCSR of the rows in
// [row_split[r], row_split[r+1]).  A logical update A[i, j] = v is therefore routed twice: to owner(j) and to owner(i).
//
// Exchange of a routed batch — ONE fused route+push kernel over NVLink peer memory, no host round trip:
//   k_route_count   per tile of 1024 ops: how many go to each owner, for both orientations
//   k_route_scan    exclusive scan of the tile counts (stable partition offsets) + this rank's send counts
//   k_route_push    every op is STORED DIRECTLY into the owner's receive region (peer pointer obtained through CUDA IPC), at
//                   region(src = me) + stable rank: the all-to-all is the tail of the routing kernel, not a separate collective
//   ncclAllGather   of the 2 x W send counts: the only collective of the exchange; it is also the barrier that makes every
//                   peer's stores visible (a rank contributes after its push kernel has completed)
//   k_unpack        the receiver compacts its W regions into dense (rows, cols, vals) arrays in rank-major, arrival-within-rank
//                   order (= the batch's global op order: last writer wins stays well defined) and leaves the count on the device
//   local batch     the per-GPU pipeline takes the device-side count (k_col_lookup is grid-stride over an upper bound); the host
//                   learns it at the pipeline's own first synchronisation point.
// Receive regions are double-buffered by batch parity; the count all-gather of batch s+1 orders "peer consumed batch s-1" before
// "I overwrite its region for batch s+1".  Regions hold a full share per (src, dst) pair, so no skew can overflow them.
// Fallback transport (DSA_DIST_TRANSPORT=nccl, or when IPC mapping is unavailable): counts all-gather -> host -> grouped
// ncclSend/ncclRecv of exact sizes.
// SpMV: every rank computes its y slice straight into the gather buffer, then one in-place ncclAllGather.
//
// NCCL is bound at run time (dlopen "libnccl.so.2": the copy the host process already loaded, e.g. PyTorch's, else the system
// one), so single-GPU users of libdsa do not need NCCL at all.
#pragma once
#include <dlfcn.h>
#include <nccl.h>
#include "pcsr.cuh"

namespace dsa {

struct NcclApi {
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    decltype(&ncclAllGather) AllGather = nullptr;
    decltype(&ncclAllReduce) AllReduce = nullptr;
    decltype(&ncclSend) Send = nullptr;
    decltype(&ncclRecv) Recv = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclGetVersion) GetVersion = nullptr;
    void* handle = nullptr;
};

inline NcclApi& nccl() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        void* h = nullptr;
        for (const char* nm : names) {   // the copy already mapped into the process first (PyTorch bundles its own)
            h = dlopen(nm, RTLD_NOW | RTLD_NOLOAD);
            if (h) break;
        }
        for (int i = 0; !h && i < 2; ++i) h = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
        if (!h) return;
        api.handle = h;
#define DSA_NCCL_SYM(f) api.f = (decltype(api.f))dlsym(h, "nccl" #f)
        DSA_NCCL_SYM(GetUniqueId); DSA_NCCL_SYM(CommInitRank); DSA_NCCL_SYM(CommDestroy); DSA_NCCL_SYM(GetErrorString);
        DSA_NCCL_SYM(AllGather); DSA_NCCL_SYM(AllReduce); DSA_NCCL_SYM(Send); DSA_NCCL_SYM(Recv); DSA_NCCL_SYM(GroupStart);
        DSA_NCCL_SYM(GroupEnd); DSA_NCCL_SYM(GetVersion);
#undef DSA_NCCL_SYM
    });
    if (!api.handle || !api.AllGather || !api.CommInitRank)
        throw DsaError{DSA_ERR_NCCL, "NCCL is not available: dlopen(\"libnccl.so.2\") failed (multi-GPU entry points need it)"};
    return api;
}

#define DSA_NCCL(expr)                                                                                         \
    do {                                                                                                       \
        ncclResult_t _r = (expr);                                                                              \
        if (_r != ncclSuccess)                                                                                 \
            throw ::dsa::DsaError{DSA_ERR_NCCL, std::string(#expr) + ": " + ::dsa::nccl().GetErrorString(_r)};  \
    } while (0)

constexpr int DIST_MAX_RANKS = 16;      // one box: 8 GPUs (the bins of the routing kernels live in registers / shared memory)
constexpr int RT_THREADS = 256, RT_ITEMS = 4, RT_TILE = RT_THREADS * RT_ITEMS;

struct RouteTables {
    int64_t split[2][DIST_MAX_RANKS];   // interior splitters: [0] by column (column-major owner), [1] by row (row-major owner)
    int nsplit;                         // world - 1
    int world, me;
    int omask;                          // bit o set = orientation o takes part (deletes route to the twin orientation only)
};

// owner(key) = number of interior splitters <= key
__device__ __forceinline__ int route_owner(const int64_t* __restrict__ sp, int ns, int64_t k) {
    int o = 0;
#pragma unroll 4
    for (int i = 0; i < ns; ++i) o += sp[i] <= k;
    return o;
}

__global__ void __launch_bounds__(RT_THREADS) k_route_count(const int64_t* __restrict__ rows, const int64_t* __restrict__ cols, int64_t n,
                                                             RouteTables T, int32_t* __restrict__ tile_cnt, int64_t* __restrict__ bad_flag) {
    // bad_flag[0]: a key < 1 (refused); bad_flag[1]: a key >= 2^32 (this rank's share travels as 24-byte triples instead of
    // 16-byte {row << 32 | col, value} pairs)
    __shared__ int hist[2][DIST_MAX_RANKS];
    int bad = 0, wide = 0;
    if (threadIdx.x < 2 * DIST_MAX_RANKS) (&hist[0][0])[threadIdx.x] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * RT_TILE;
#pragma unroll
    for (int j = 0; j < RT_ITEMS; ++j) {
        const int64_t i = base + j * RT_THREADS + threadIdx.x;
        int oc = -1, orr = -1;
        if (i < n) {
            const int64_t c = cols[i], r = rows[i];
            if (c < 1 || r < 1) bad = 1;   // device contract: keys >= 1 (key 0 is the semaphore key, pcsr.jl:23)
            if (((unsigned long long)c | (unsigned long long)r) >> 32) wide = 1;
            oc = route_owner(T.split[0], T.nsplit, c);
            orr = route_owner(T.split[1], T.nsplit, r);
        }
        // one shared atomic per (warp, owner) instead of one per op
        const unsigned act = __ballot_sync(0xffffffffu, i < n);
        if (i < n) {
            if (T.omask & 1) {
                const unsigned mc = __match_any_sync(act, oc);
                if ((threadIdx.x & 31) == __ffs(mc) - 1) atomicAdd(&hist[0][oc], __popc(mc));
            }
            if (T.omask & 2) {
                const unsigned mr = __match_any_sync(act, orr);
                if ((threadIdx.x & 31) == __ffs(mr) - 1) atomicAdd(&hist[1][orr], __popc(mr));
            }
        }
    }
    if (bad) bad_flag[0] = 1;
    if (wide) bad_flag[1] = 1;
    __syncthreads();
    if (threadIdx.x < 2 * T.world) {
        const int o = threadIdx.x / T.world, d = threadIdx.x % T.world;
        tile_cnt[((int64_t)blockIdx.x * 2 + o) * T.world + d] = hist[o][d];
    }
}

// one CTA per (orientation, owner) bin: exclusive scan of the bin's tile counts (stable partition offsets); the total is this
// rank's send count for that bin
__global__ void __launch_bounds__(1024) k_route_scan(const int32_t* __restrict__ tile_cnt, int64_t ntiles, int world,
                                                      int32_t* __restrict__ tile_off, int64_t* __restrict__ send_counts) {
    __shared__ int warp_tot[32];
    __shared__ int carry_s;
    const int bin = blockIdx.x, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int64_t t0 = 0; t0 < ntiles; t0 += 1024) {
        const int64_t t = t0 + threadIdx.x;
        const int v = t < ntiles ? tile_cnt[t * 2 * world + bin] : 0;
        int s = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= o) s += u;
        }
        if (lane == 31) warp_tot[wid] = s;
        __syncthreads();
        if (wid == 0) {
            int wv = warp_tot[lane], ws = wv;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int u = __shfl_up_sync(0xffffffffu, ws, o);
                if (lane >= o) ws += u;
            }
            warp_tot[lane] = ws - wv;   // exclusive over the warps
        }
        __syncthreads();
        const int carry = carry_s;
        if (t < ntiles) tile_off[t * 2 * world + bin] = carry + warp_tot[wid] + s - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = carry + warp_tot[wid] + s;
        __syncthreads();
    }
    if (threadIdx.x == 0) send_counts[bin] = carry_s;   // bin = o * world + dst
}

struct PushTargets {
    int64_t* base[DIST_MAX_RANKS];   // receive buffer (this batch's parity) of every rank, as mapped into this process
    int region[DIST_MAX_RANKS];      // region index written at destination d: this rank (peer stores) or d (local send staging)
};
// receive buffer layout (int64 words): [orientation o][source rank][array a in rows, cols, vals][region_cap]
__host__ __device__ __forceinline__ int64_t region_word(int o, int src, int a, int world, int64_t region_cap) {
    return (((int64_t)o * world + src) * 3 + a) * region_cap;
}

// The tile's ops are first laid out in shared memory in (owner, arrival) order, then copied out: consecutive threads store
// consecutive words of one owner's run, so the peer stores leave the SM as full 128-byte lines (about RT_TILE / world ops = 1 KB
// per owner, array and tile).  Storing straight from registers split every warp store into `world` 32-byte pieces: 196 us per
// launch at 8 GPUs against 70 us at 2 (NVLink packets of 32 bytes), the limiter of the 8-GPU step.
__global__ void __launch_bounds__(RT_THREADS) k_route_push(const int64_t* __restrict__ rows, const int64_t* __restrict__ cols,
                                                            const double* __restrict__ vals, int64_t n, RouteTables T,
                                                            const int32_t* __restrict__ tile_off, PushTargets P, int64_t region_cap,
                                                            int64_t* __restrict__ wide_flag) {
    // every key of this rank's share fits 32 bits (the common case): {row << 32 | col, value} = 16 bytes per op and owner instead of 24.
    // wide_flag[0] (from k_route_count) decides; wide_flag[1] <- what was used (the receivers' unpack reads it after the all-gather)
    const bool packed = wide_flag && wide_flag[0] == 0;
    if (wide_flag && blockIdx.x == 0 && threadIdx.x == 0) wide_flag[1] = packed ? 1 : 0;
    constexpr int NSEG = RT_ITEMS * (RT_THREADS / 32);
    // counts of every (slab j, warp w) per bin, then their exclusive prefix in (j, w) order = index order inside the tile
    __shared__ int wcnt[NSEG][2][DIST_MAX_RANKS];
    __shared__ int binstart[2][DIST_MAX_RANKS + 1];   // first staged index of every owner's run
    __shared__ int toff[2][DIST_MAX_RANKS];           // offset of the tile's run inside the owner's region
    __shared__ int64_t s_r[RT_TILE], s_c[RT_TILE];
    __shared__ double s_v[RT_TILE];
    for (int t = threadIdx.x; t < NSEG * 2 * DIST_MAX_RANKS; t += RT_THREADS) (&wcnt[0][0][0])[t] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const unsigned lt = (1u << lane) - 1u;
    const int64_t base = (int64_t)blockIdx.x * RT_TILE;
    int own[RT_ITEMS][2], rk[RT_ITEMS][2];
    int64_t r_[RT_ITEMS], c_[RT_ITEMS];
    double v_[RT_ITEMS];
#pragma unroll
    for (int j = 0; j < RT_ITEMS; ++j) {
        const int64_t i = base + j * RT_THREADS + threadIdx.x;
        const bool in = i < n;
        own[j][0] = own[j][1] = -1;
        rk[j][0] = rk[j][1] = 0;
        if (in) {
            r_[j] = rows[i]; c_[j] = cols[i]; v_[j] = vals[i];
            own[j][0] = route_owner(T.split[0], T.nsplit, c_[j]);
            own[j][1] = route_owner(T.split[1], T.nsplit, r_[j]);
        }
        const unsigned act = __ballot_sync(0xffffffffu, in);
        if (in) {
#pragma unroll
            for (int o = 0; o < 2; ++o) {
                if (!(T.omask & (1 << o))) continue;
                const unsigned m = __match_any_sync(act, own[j][o]);
                rk[j][o] = __popc(m & lt);
                if (lane == __ffs(m) - 1) wcnt[j * (RT_THREADS / 32) + w][o][own[j][o]] = __popc(m);
            }
        }
    }
    __syncthreads();
    if (threadIdx.x < 2 * T.world) {   // serial exclusive prefix over the 32 (slab, warp) segments of one bin
        const int o = threadIdx.x / T.world, d = threadIdx.x % T.world;
        int run = 0;
        for (int sgm = 0; sgm < NSEG; ++sgm) {
            const int c = wcnt[sgm][o][d];
            wcnt[sgm][o][d] = run;
            run += c;
        }
        binstart[o][d + 1] = run;   // count, turned into a start below
        toff[o][d] = tile_off[((int64_t)blockIdx.x * 2 + o) * T.world + d];
    }
    __syncthreads();
    if (threadIdx.x < 2) {
        const int o = threadIdx.x;
        int run = 0;
        binstart[o][0] = 0;
        for (int d = 0; d < T.world; ++d) {
            const int c = binstart[o][d + 1];
            binstart[o][d + 1] = run + c;
            run += c;
        }
    }
    __syncthreads();
    for (int o = 0; o < 2; ++o) {
        if (!(T.omask & (1 << o))) continue;
#pragma unroll
        for (int j = 0; j < RT_ITEMS; ++j) {
            const int d = own[j][o];
            if (d < 0) continue;
            const int li = binstart[o][d] + wcnt[j * (RT_THREADS / 32) + w][o][d] + rk[j][o];
            s_r[li] = packed ? (int64_t)(((unsigned long long)r_[j] << 32) | (unsigned long long)c_[j]) : r_[j];
            if (!packed) s_c[li] = c_[j];
            s_v[li] = v_[j];
        }
        __syncthreads();
        const int cnt = binstart[o][T.world];
        for (int t = threadIdx.x; t < cnt; t += RT_THREADS) {
            int d = 0;
            while (t >= binstart[o][d + 1]) ++d;
            const int64_t pos = toff[o][d] + (t - binstart[o][d]);
            int64_t* dst = P.base[d] + region_word(o, P.region[d], 0, T.world, region_cap) + pos;   // NVLink peer store (local when d == me)
            dst[0] = s_r[t];
            if (!packed) dst[region_cap] = s_c[t];
            dst[2 * region_cap] = __double_as_longlong(s_v[t]);
        }
        __syncthreads();
    }
    // One system-scope fence per CTA (cumulative over the CTA's stores through the barrier); kernel completion + the count
    // all-gather order the rest.
    if (threadIdx.x == 0) __threadfence_system();
}

// counts[src * row_stride + o * world + dst] (all-gathered).  The receiver's dense arrays: ops of source 0 first, then source 1, ... (rank-major,
// arrival order within a rank).  grid.y = orientation.  n_out[o] = total.
__global__ void __launch_bounds__(256) k_dist_unpack(const int64_t* __restrict__ counts, int row_stride, int world, int me, const int64_t* __restrict__ rbuf,
                                                      int64_t region_cap, int64_t* __restrict__ out_rows0, int64_t* __restrict__ out_cols0,
                                                      double* __restrict__ out_vals0, int64_t* __restrict__ out_rows1,
                                                      int64_t* __restrict__ out_cols1, double* __restrict__ out_vals1,
                                                      int64_t* __restrict__ n_out) {
    __shared__ int64_t off[DIST_MAX_RANKS + 1];
    __shared__ int pk[DIST_MAX_RANKS];   // source s sent {row << 32 | col, value} pairs
    const int o = blockIdx.y;
    if (threadIdx.x == 0) {
        int64_t run = 0;
        for (int s = 0; s < world; ++s) {
            off[s] = run;
            run += counts[(int64_t)s * row_stride + o * world + me];
            pk[s] = counts[(int64_t)s * row_stride + 2 * world + 2] != 0;
        }
        off[world] = run;
        if (blockIdx.x == 0) n_out[o] = run;
    }
    __syncthreads();
    int64_t* orows = o ? out_rows1 : out_rows0;
    int64_t* ocols = o ? out_cols1 : out_cols0;
    double* ovals = o ? out_vals1 : out_vals0;
    const int64_t tot = off[world];
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < tot; i += (int64_t)gridDim.x * blockDim.x) {
        int s = 0;
        while (i >= off[s + 1]) ++s;
        const int64_t* src = rbuf + region_word(o, s, 0, world, region_cap) + (i - off[s]);
        const int64_t a = src[0];
        orows[i] = pk[s] ? (int64_t)((unsigned long long)a >> 32) : a;
        ocols[i] = pk[s] ? (int64_t)((unsigned long long)a & 0xffffffffull) : src[region_cap];
        ovals[i] = __longlong_as_double(src[2 * region_cap]);
    }
}

// y of the padded gather buffer (slice r at r * per) -> dense y indexed by key
__global__ void __launch_bounds__(256) k_dist_assemble_y(const double* __restrict__ ybuf, int64_t per, int world, RouteTables T, int which,
                                                          int64_t key_hi_last, double* __restrict__ y, int64_t ny) {
    const int64_t k0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // key - 1
    if (k0 >= ny) return;
    const int64_t key = k0 + 1;
    const int r = route_owner(T.split[which], T.nsplit, key);
    const int64_t lo = r == 0 ? 1 : T.split[which][r - 1];
    (void)key_hi_last;
    y[k0] = ybuf[(int64_t)r * per + (key - lo)];
}

}  // namespace dsa
