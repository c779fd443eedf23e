// K7 SpMV, experimental variants (opt-in with DSA_SPMV_BULK; k_spmv_flat stays the default).
//
// Measured on B200, config 2 (profiles/exp_r01_spmv_bulk.log): flat 86 us; bulk-staged 122 us with 16 consumer warps per SM,
// 91.5 us with 32 — the results are bit-identical, but the kernel is NOT bound by the bytes the stream keeps in flight (the
// hypothesis below was wrong): time scales with the number of warps issuing gathers, and the floor that fits both kernels is the
// L1 replay rate of divergent loads (~2 cycles per distinct 128 B line, B300_MICROARCH.md: 1e7 lines / 148 SMs x 2.07 cycles = 71 us).
// k_spmv_dsmem (mode 8) moves the gathers off the L1 path: x is spread over the shared memory of an 8-CTA cluster and read
// through distributed shared memory.  Measured: 186 us (profiles/exp_r01_spmv_dsmem.log) — random 8-byte DSMEM reads run at
// ~0.2 words/cycle/SM, far below the L1 path.  Both variants stay here as measured dead ends with their switches.
//
// --- bulk-copy pipelined variant (modes 1..5) ---
// Why: k_spmv_flat is bound by the bytes each warp keeps in flight.  A warp loads its chunk (4 x 32 cells x 16 B = 2 KB), waits a
// DRAM round trip, gathers x, waits an L2 round trip, reduces.  With ~30 resident warps per SM about half of them are in the
// stream phase at any time: ~30 KB in flight per SM against the ~44 KB that 6.5 TB/s x 1 us / 148 SMs needs, and the L1 wavefronts
// of the x gathers (1e7 lines, ~34 us of L1 time per SM at config 2) do not overlap the stream of the same warp
// (profiles/README.md, DESIGN.md §10.1: stream alone 55 us, kernel 88 us).
//
// Here the stream is decoupled from the warps: one producer lane per CTA moves tiles of TILE cells (keys + values) from HBM to a
// ring of STAGES shared-memory stages with cp.async.bulk (the 1-D TMA path: no tensor map, completion counted in bytes on an
// mbarrier), so STAGES x TILE x 16 B (128 KB by default) stay in flight per SM regardless of what the consumer warps are doing;
// the consumer warps read their chunk from shared memory and spend their time on the x gathers and the segmented reduction.
// Chunk numbering, per-chunk arithmetic and summation order are those of k_spmv_flat<.,4>, so the two kernels produce the same
// bits and share k_spmv_fixup.
#pragma once

namespace dsa {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// A protocol error must not hang the GPU: after 2^20 failed probes (a legitimate wait is microseconds; a probe already blocks for a hardware-defined
// time slice) the kernel traps, which surfaces as a CUDA error on the next synchronisation.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity))
        if (++spins == (1u << 20)) __trap();
}
// 1-D bulk copy global -> shared, completion signalled on `bar` in bytes (dst, src and bytes are multiples of 16)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
                 "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// One chunk of 4 x 32 cells held in registers (k = keys, t = values): identical arithmetic to k_spmv_flat<.,4>.
struct GatherLdg {   // x[kk - 1] through the read-only path (what k_spmv_flat does)
    const double* __restrict__ x;
    __device__ __forceinline__ double operator()(int64_t kk) const { return __ldg(x + (kk - 1)); }
};

template <bool SPARSE_X, typename Gather>
__device__ __forceinline__ void spmv_chunk4(int64_t (&k)[4], double (&t)[4], int lane, unsigned lt, int64_t chunk, const Gather& gather,
                                            const uint8_t* __restrict__ xmask, int64_t nx, double* __restrict__ yslot,
                                            int32_t* __restrict__ ycnt, double* __restrict__ carry, int32_t* __restrict__ carry_cnt,
                                            int32_t* __restrict__ chunk_last_slot) {
    int32_t tc[4];
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        tc[s] = 0;
        const int64_t kk = k[s];
        if (kk > 0) {
            double xv = 0.0;
            bool present = kk <= nx;
            if (SPARSE_X) present = present && xmask[kk - 1] != 0;
            if (present) {
                xv = gather(kk);
                tc[s] = 1;
            }
            t[s] = present ? __dmul_rn(xv, t[s]) : 0.0;
        } else if (kk != 0) {
            t[s] = 0.0;
        }
    }
    int32_t cur_slot = -1;
    double acc = 0.0;
    int32_t acc_cnt = 0;
    double lacc = 0.0;
    int32_t lcnt = 0;
    bool prefix_open = true;
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        const bool head = k[s] == 0;
        const unsigned hb = __ballot_sync(0xffffffffu, head);
        if (hb == 0) {
            lacc = __dadd_rn(lacc, t[s]);
            lcnt += tc[s];
            continue;
        }
        acc = __dadd_rn(acc, warp_sum_f64(lacc));
        acc_cnt += warp_sum_i32(lcnt);
        lacc = 0.0;
        lcnt = 0;
        const double tv = head ? 0.0 : t[s];
        const unsigned hle = hb & (lt | (1u << lane));
        const int seg_lo = hle ? 31 - __clz(hle) : 0;
        double st = tv;
        int32_t sc = tc[s];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double ot = __shfl_up_sync(0xffffffffu, st, o);
            const int32_t oc = __shfl_up_sync(0xffffffffu, sc, o);
            if (lane >= o && lane - o >= seg_lo) {
                st = __dadd_rn(ot, st);
                sc += oc;
            }
        }
        const double prev_t = __shfl_up_sync(0xffffffffu, st, 1);
        const int32_t prev_c = __shfl_up_sync(0xffffffffu, sc, 1);
        const unsigned hlt = hb & lt;
        const int prev_head_lane = hlt ? 31 - __clz(hlt) : -1;
        const int32_t my_slot = head ? (int32_t)t[s] - 1 : -1;
        const int32_t prev_head_slot = __shfl_sync(0xffffffffu, my_slot, prev_head_lane < 0 ? 0 : prev_head_lane);
        if (head) {
            double tot = lane > 0 ? prev_t : 0.0;
            int32_t totc = lane > 0 ? prev_c : 0;
            if (prev_head_lane < 0) {
                tot = lane > 0 ? __dadd_rn(acc, tot) : acc;
                totc += acc_cnt;
                if (cur_slot >= 0) {
                    yslot[cur_slot] = tot;
                    ycnt[cur_slot] = totc;
                } else if (prefix_open) {
                    carry[chunk] = tot;
                    carry_cnt[chunk] = totc;
                }
            } else {
                yslot[prev_head_slot] = tot;
                ycnt[prev_head_slot] = totc;
            }
        }
        const int last_head_lane = 31 - __clz(hb);
        cur_slot = __shfl_sync(0xffffffffu, my_slot, last_head_lane);
        acc = __shfl_sync(0xffffffffu, st, 31);
        acc_cnt = __shfl_sync(0xffffffffu, sc, 31);
        prefix_open = false;
    }
    acc = __dadd_rn(acc, warp_sum_f64(lacc));
    acc_cnt += warp_sum_i32(lcnt);
    if (lane == 0) {
        if (cur_slot >= 0) {
            yslot[cur_slot] = acc;
            ycnt[cur_slot] = acc_cnt;
        } else {
            carry[chunk] = acc;
            carry_cnt[chunk] = acc_cnt;
        }
        chunk_last_slot[chunk] = cur_slot;
    }
}

// Persistent CTAs (grid = #SMs x CTAs per SM): CTA b streams tiles b, b + grid, b + 2 grid, ...
// Warps 0..NCONS-1 consume, warp NCONS produces (one lane).  cap must be a multiple of TILE (the host checks it).
template <bool SPARSE_X, int TILE, int STAGES, int NCONS>
__global__ void __launch_bounds__((NCONS + 1) * 32) k_spmv_bulk(const int64_t* __restrict__ keys, const double* __restrict__ vals, int64_t ntiles,
                                                                 const double* __restrict__ x, const uint8_t* __restrict__ xmask, int64_t nx,
                                                                 double* __restrict__ yslot, int32_t* __restrict__ ycnt,
                                                                 double* __restrict__ carry, int32_t* __restrict__ carry_cnt,
                                                                 int32_t* __restrict__ chunk_last_slot) {
    constexpr int CHUNK = 128;                       // cells per warp chunk (k_spmv_flat<., 4>)
    constexpr int CHUNKS_PER_TILE = TILE / CHUNK;
    static_assert(TILE % CHUNK == 0, "a tile is a whole number of chunks");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[STAGES];
    __shared__ __align__(8) uint64_t empty_bar[STAGES];
    int64_t* skeys = reinterpret_cast<int64_t*>(smem_raw);                                   // [STAGES][TILE]
    double* svals = reinterpret_cast<double*>(smem_raw + (size_t)STAGES * TILE * 8);         // [STAGES][TILE]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);          // the producer's arrive.expect_tx; the copies complete the byte count
            mbar_init(&empty_bar[s], NCONS);     // one arrive per consumer warp
        }
        mbar_fence_init();
    }
    __syncthreads();
    const int64_t first = blockIdx.x, stride = gridDim.x;
    const int64_t nmine = first < ntiles ? (ntiles - first + stride - 1) / stride : 0;
    if (warp == NCONS) {
        if (lane == 0) {
            for (int64_t i = 0; i < nmine; ++i) {
                const int s = (int)(i % STAGES);
                const uint32_t round = (uint32_t)(i / STAGES);
                mbar_wait(&empty_bar[s], (round & 1u) ^ 1u);   // passes at once in round 0 (fresh barrier), then waits for the consumers
                const int64_t cell0 = (first + i * stride) * TILE;
                mbar_arrive_expect_tx(&full_bar[s], (uint32_t)(2 * TILE * 8));
                bulk_g2s(skeys + (size_t)s * TILE, keys + cell0, (uint32_t)(TILE * 8), &full_bar[s]);
                bulk_g2s(svals + (size_t)s * TILE, vals + cell0, (uint32_t)(TILE * 8), &full_bar[s]);
            }
        }
        return;
    }
    const unsigned lt = lanemask_lt();
    for (int64_t i = 0; i < nmine; ++i) {
        const int s = (int)(i % STAGES);
        const uint32_t round = (uint32_t)(i / STAGES);
        mbar_wait(&full_bar[s], round & 1u);
        const int64_t tile = first + i * stride;
        const int64_t* tk = skeys + (size_t)s * TILE;
        const double* tv = svals + (size_t)s * TILE;
        for (int c = warp; c < CHUNKS_PER_TILE; c += NCONS) {
            int64_t k[4];
            double t[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                k[q] = tk[c * CHUNK + q * 32 + lane];
                t[q] = tv[c * CHUNK + q * 32 + lane];
            }
            spmv_chunk4<SPARSE_X>(k, t, lane, lt, tile * CHUNKS_PER_TILE + c, GatherLdg{x}, xmask, nx, yslot, ycnt, carry, carry_cnt,
                                  chunk_last_slot);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[s]);   // this warp no longer reads stage s
    }
}

// --- x in distributed shared memory (mode 8) ---
// A cluster of CL CTAs holds x: CTA r keeps x[r * 2^slice_lg .. (r+1) * 2^slice_lg) in its shared memory.  The stream is read as in
// k_spmv_flat (registers, evict-first); the gather of x[kk-1] becomes mapa + ld.shared::cluster on CTA (kk-1) >> slice_lg.
// Persistent warps: chunk = global warp, + total warps, ...  Same chunking and arithmetic as k_spmv_flat<., 4>: same bits.
struct GatherDsmem {
    uint32_t sx_addr;   // shared::cta address of this CTA's slice (every CTA uses the same offset)
    int slice_lg;
    __device__ __forceinline__ double operator()(int64_t kk) const {
        const uint64_t j = (uint64_t)(kk - 1);
        const uint32_t owner = (uint32_t)(j >> slice_lg);
        const uint32_t local = sx_addr + (uint32_t)(j & ((1ull << slice_lg) - 1)) * 8u;
        uint32_t remote;
        double v;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(owner));
        asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(remote) : "memory");
        return v;
    }
};

template <int CL>
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(1024, 1)
    k_spmv_dsmem(const int64_t* __restrict__ keys, const double* __restrict__ vals, int64_t cap, const double* __restrict__ x, int64_t nx,
                 int slice_lg, double* __restrict__ yslot, int32_t* __restrict__ ycnt, double* __restrict__ carry,
                 int32_t* __restrict__ carry_cnt, int32_t* __restrict__ chunk_last_slot, int64_t nchunks) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* sx = reinterpret_cast<double*>(smem_raw);
    uint32_t rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const int64_t slice = int64_t(1) << slice_lg;
    for (int64_t i = threadIdx.x; i < slice; i += blockDim.x) {
        const int64_t j = (int64_t)rank * slice + i;
        sx[i] = j < nx ? x[j] : 0.0;
    }
    // every CTA's slice is complete before anyone reads it (release/acquire at cluster scope)
    __syncwarp();
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    const GatherDsmem gather{smem_u32(sx), slice_lg};
    const int lane = threadIdx.x & 31;
    const unsigned lt = lanemask_lt();
    const int64_t warps_total = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t chunk = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); chunk < nchunks; chunk += warps_total) {
        const int64_t base = chunk * 128;
        int64_t k[4];
        double t[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int64_t p = base + q * 32 + lane;
            k[q] = GAP_KEY;
            t[q] = 0.0;
            if (p < cap) {
                k[q] = __ldcs(keys + p);
                t[q] = __ldcs(vals + p);
            }
        }
        spmv_chunk4<false>(k, t, lane, lt, chunk, gather, nullptr, nx, yslot, ycnt, carry, carry_cnt, chunk_last_slot);
    }
    // no CTA may exit while its slice can still be read by the others
    __syncwarp();
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

}  // namespace dsa
