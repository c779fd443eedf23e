// EXPERIMENTAL (opt-in with DSA_ILP=2 or 4; not validated on hardware yet): several ops per thread for the latency-bound per-op
// kernels of the update pipeline.
//
// Why (profiles/ncu_full_r01d.md + `ncu --page source` of the same report): k_locate, k_col_lookup, k_bucket_scatter,
// k_apply_hits and k_compact_inserts spend 50-80 % of their issue slots stalled on long-scoreboard dependencies (a chain of 2-10
// dependent loads per op, one op per thread), with DRAM at 10-43 % and no unit above 60 %.  A 1M-op batch is 3.3 waves of
// 256-thread CTAs, so the kernel time is ~3.3 x the latency of one op's chain.  With ITEMS ops per thread the chains of ITEMS ops
// are in flight together and the batch fits one wave.
//
// Also here: k_get_ilp (batched getindex, same multi-search) and k_insert_leaf_info_gallop (run end by galloping instead of a
// ~19-step binary search; enabled by any DSA_ILP value).
//
// Every kernel here computes exactly what its one-op-per-thread twin computes (same outputs for every op; the only difference is
// the arrival order of the bucket-count atomics, which the bucket path is independent of by construction: lidx is only used as a
// unique slot inside the bucket and the in-bucket rank is by (key, arrival)).  tests/test_zz_experimental.py compares the final
// layouts bit for bit.
#pragma once

namespace dsa {

inline int ilp_items() {
    static const int items = [] {
        const char* e = getenv("DSA_ILP");
        const int v = e ? atoi(e) : 0;
        return (v == 2 || v == 4) ? v : 0;
    }();
    return items;
}

// ---- K6 locate: ITEMS gapped binary searches per thread, advanced in lock step (one load per search per round) ----------------
// The reference's search (finds.jl:29-57) as a state machine, so that the loads of ITEMS searches are issued together:
//   state 1 = probing: load keys[cur]; a gap with cur > lo walks left (finds.jl:33-35); otherwise the probe is decided
//             (gap: lo = mid + 1; greater: hi = cur - 1; smaller: lo = mid + 1; equal: hit) and the next probe or the final walk starts
//   state 2 = final walk left from hi to the nearest element or off the front (finds.jl:49-56): the predecessor
// go[k] = false skips search k.  Same (pos, hit) as gapped_find for every query.  __host__ too: dsa_find_multi_host runs it on the CPU
// for tests/test_hostlogic.py.
#if defined(__CUDA_ARCH__)
#define DSA_UNROLL _Pragma("unroll")
#else
#define DSA_UNROLL   // host pass of a __host__ __device__ function: gcc does not know the pragma
#endif
template <int ITEMS>
__host__ __device__ __forceinline__ void gapped_find_multi(const int64_t* __restrict__ keys, const int64_t* key, const int64_t* from,
                                                           const int64_t* to, const bool* go, int64_t* pos, bool* hit) {
    int64_t lo[ITEMS], hi[ITEMS], mid[ITEMS], cur[ITEMS];
    int state[ITEMS];
DSA_UNROLL
    for (int k = 0; k < ITEMS; ++k) {
        hit[k] = false;
        state[k] = 0;
        lo[k] = hi[k] = mid[k] = cur[k] = 0;
        if (!go[k]) continue;
        pos[k] = -1;
        lo[k] = from[k];
        hi[k] = to[k];
        if (lo[k] <= hi[k]) {
            mid[k] = (lo[k] + hi[k]) >> 1;
            cur[k] = mid[k];
            state[k] = 1;
        } else {
            cur[k] = hi[k];
            state[k] = 2;
        }
    }
    while (true) {
        bool any = false;
        int64_t kv[ITEMS];
DSA_UNROLL
        for (int k = 0; k < ITEMS; ++k) {   // one load per unfinished search, all in flight together
            kv[k] = 0;
            if (state[k] == 2 && cur[k] < 0) {   // walked off the front: no predecessor
                pos[k] = cur[k];
                state[k] = 0;
            }
            if (state[k] != 0) {
                kv[k] = keys[cur[k]];
                any = true;
            }
        }
        if (!any) break;
DSA_UNROLL
        for (int k = 0; k < ITEMS; ++k) {
            if (state[k] == 1) {
                if (kv[k] == GAP_KEY && cur[k] > lo[k]) {   // walk left to the nearest element (finds.jl:33-35)
                    cur[k] -= 1;
                    continue;
                }
                if (kv[k] == GAP_KEY) {
                    lo[k] = mid[k] + 1;
                } else if (kv[k] > key[k]) {
                    hi[k] = cur[k] - 1;
                } else if (kv[k] < key[k]) {
                    lo[k] = mid[k] + 1;
                } else {
                    hit[k] = true;
                    pos[k] = cur[k];
                    state[k] = 0;
                    continue;
                }
                if (lo[k] <= hi[k]) {
                    mid[k] = (lo[k] + hi[k]) >> 1;
                    cur[k] = mid[k];
                } else {
                    cur[k] = hi[k];
                    state[k] = 2;
                }
            } else if (state[k] == 2) {
                if (kv[k] == GAP_KEY) {
                    cur[k] -= 1;
                } else {
                    pos[k] = cur[k];
                    state[k] = 0;
                }
            }
        }
    }
}

// host driver of the state machine (dsa_find_multi_host): queries in groups of ITEMS
template <int ITEMS>
inline void find_multi_host(const int64_t* keys, const int64_t* q, const int64_t* from, const int64_t* to, int64_t nq, int64_t* pos_out,
                            uint8_t* hit_out) {
    for (int64_t b = 0; b < nq; b += ITEMS) {
        int64_t key[ITEMS], f[ITEMS], t[ITEMS], pos[ITEMS];
        bool go[ITEMS], hit[ITEMS];
        for (int k = 0; k < ITEMS; ++k) {
            go[k] = b + k < nq;
            key[k] = go[k] ? q[b + k] : 0;
            f[k] = go[k] ? from[b + k] : 0;
            t[k] = go[k] ? to[b + k] : -1;
            pos[k] = -1;
        }
        gapped_find_multi<ITEMS>(keys, key, f, t, go, pos, hit);
        for (int k = 0; k < ITEMS; ++k)
            if (go[k]) {
                pos_out[b + k] = pos[k];
                hit_out[b + k] = hit[k] ? 1 : 0;
            }
    }
}
template <int ITEMS>
__global__ void __launch_bounds__(256) k_locate_ilp(const int64_t* __restrict__ keys, int64_t cap, const int32_t* __restrict__ op_pid,
                                                     const int64_t* __restrict__ op_key, const double* __restrict__ op_val, int64_t nops,
                                                     const int64_t* __restrict__ sem, const int32_t* __restrict__ next_slot,
                                                     int64_t* __restrict__ op_pos, uint8_t* __restrict__ op_flag,
                                                     const int64_t* __restrict__ n_dev, const uint8_t* __restrict__ op_dead) {
    const int64_t count = n_dev ? *n_dev : nops;
    const int64_t base = (int64_t)blockIdx.x * (256 * ITEMS) + threadIdx.x;
    int64_t key[ITEMS], from[ITEMS], to[ITEMS], pos[ITEMS];
    int32_t pid[ITEMS];
    bool act[ITEMS];      // the op exists and is not superseded
    bool go[ITEMS];       // ... and needs a search
    bool is_set[ITEMS], hit[ITEMS];
    // the op's own fields: ITEMS independent loads in flight
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
        const int64_t i = base + (int64_t)k * 256;
        act[k] = i < count;
        if (act[k] && op_dead && op_dead[i]) {   // overwritten by a later op of the same batch (last writer wins)
            op_flag[i] = 0;
            act[k] = false;
        }
        key[k] = 0;
        is_set[k] = false;
        pid[k] = -1;
        if (act[k]) {
            key[k] = op_key[i];
            is_set[k] = op_val[i] != 0.0;
            if (op_pid) pid[k] = op_pid[i];
        }
    }
    // the partition spans: two more rounds of independent loads
    int64_t s[ITEMS], e[ITEMS];
    int32_t ns[ITEMS];
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
        s[k] = 0;
        ns[k] = -1;
        if (act[k] && op_pid) {
            s[k] = sem[pid[k]];
            ns[k] = next_slot[pid[k]];
        }
    }
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
        e[k] = cap;
        if (act[k] && op_pid && ns[k] >= 0) e[k] = sem[ns[k]];
    }
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
        pos[k] = -1;
        from[k] = to[k] = 0;
        go[k] = act[k];
        if (!act[k]) continue;
        if (op_pid && (s[k] < 0 || key[k] == 0)) {
            pos[k] = e[k] - 1;   // new partition (or its semaphore): before the next live semaphore (pcsr.jl:121-126,101)
            go[k] = false;
            continue;
        }
        from[k] = op_pid ? (is_set[k] ? s[k] + 1 : s[k]) : 0;   // inserts search (sem, end], deletes [sem, end] (pcsr.jl:305-307)
        to[k] = op_pid ? e[k] - 1 : cap - 1;
    }
    gapped_find_multi<ITEMS>(keys, key, from, to, go, pos, hit);
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
        if (!act[k]) continue;
        const int64_t i = base + (int64_t)k * 256;
        op_pos[i] = pos[k];
        op_flag[i] = hit[k] ? (is_set[k] ? FL_OVERWRITE : FL_DELETE) : ((is_set[k] || (op_pid && key[k] == 0)) ? FL_INSERT : 0);
    }
}

// ---- batched getindex: ITEMS finds per thread (pma.jl:189-193 / pcsr.jl:228-232) --------------------------------------------------
template <int ITEMS>
__global__ void __launch_bounds__(256) k_get_ilp(const int64_t* __restrict__ keys, const double* __restrict__ vals, int64_t cap,
                                                  const int32_t* __restrict__ q_pid, const int64_t* __restrict__ q_key, int64_t nq,
                                                  const int64_t* __restrict__ sem, const int32_t* __restrict__ next_slot,
                                                  double* __restrict__ out) {
    const int64_t base = (int64_t)blockIdx.x * (256 * ITEMS) + threadIdx.x;
    int64_t key[ITEMS], from[ITEMS], to[ITEMS], pos[ITEMS], s[ITEMS];
    int32_t pid[ITEMS], ns[ITEMS];
    bool in[ITEMS], go[ITEMS], hit[ITEMS];
DSA_UNROLL
    for (int k = 0; k < ITEMS; ++k) {
        const int64_t i = base + (int64_t)k * 256;
        in[k] = i < nq;
        key[k] = 0;
        pid[k] = -1;
        if (in[k]) {
            key[k] = q_key[i];
            if (q_pid) pid[k] = q_pid[i];
        }
    }
DSA_UNROLL
    for (int k = 0; k < ITEMS; ++k) {
        s[k] = -1;
        ns[k] = -1;
        if (in[k] && q_pid && pid[k] >= 0) {   // pid < 0: the column is absent (pcsr.jl:263-265)
            s[k] = sem[pid[k]];
            ns[k] = next_slot[pid[k]];
        }
    }
DSA_UNROLL
    for (int k = 0; k < ITEMS; ++k) {
        pos[k] = -1;
        hit[k] = false;
        from[k] = 0;
        to[k] = cap - 1;
        go[k] = in[k];
        if (q_pid) {
            go[k] = in[k] && pid[k] >= 0 && s[k] >= 0;
            if (go[k]) {
                from[k] = s[k];
                to[k] = (ns[k] >= 0 ? sem[ns[k]] : cap) - 1;
            }
        }
    }
    gapped_find_multi<ITEMS>(keys, key, from, to, go, pos, hit);
    double v[ITEMS];
DSA_UNROLL
    for (int k = 0; k < ITEMS; ++k) v[k] = (go[k] && hit[k]) ? vals[pos[k]] : 0.0;
DSA_UNROLL
    for (int k = 0; k < ITEMS; ++k) {
        const int64_t i = base + (int64_t)k * 256;
        if (in[k]) out[i] = v[k];
    }
}

// ---- per-leaf bookkeeping of the compacted inserts: the run end by galloping -----------------------------------------------------
// k_insert_leaf_info finds the end of a leaf's run of inserts with a binary search over [j+1, n): ~19 dependent loads per leaf at
// 500k inserts, although a run is 1-3 inserts long.  Galloping (probe j+1, j+2, j+4, ...) brackets the end in 1-3 loads, then the
// binary search runs inside the bracket.  Same result: the first index >= j+1 whose predecessor position is >= leaf_end.
__global__ void __launch_bounds__(256) k_insert_leaf_info_gallop(const int64_t* __restrict__ ins_pos, const int64_t* __restrict__ nins_dev,
                                                                  int32_t* __restrict__ inscnt, int32_t* __restrict__ ins_first,
                                                                  uint8_t* __restrict__ touched, int lgS, ActiveLeaf* __restrict__ act,
                                                                  int64_t* __restrict__ nact_dev) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t n = *nins_dev;
    if (j >= n) return;
    const int64_t pos = ins_pos[j];
    const int64_t leaf = (pos < 0 ? 0 : pos) >> lgS;
    if (j > 0) {
        const int64_t pq = ins_pos[j - 1];
        if (((pq < 0 ? 0 : pq) >> lgS) == leaf) return;
    }
    const int64_t leaf_end = (leaf + 1) << lgS;
    int64_t lo = j + 1, hi = n;
    for (int64_t step = 1;; step <<= 1) {   // invariant: ins_pos[lo - 1] < leaf_end (positions are sorted)
        const int64_t p = j + step;
        if (p >= n) break;
        if (ins_pos[p] >= leaf_end) {
            hi = p;
            break;
        }
        lo = p + 1;
    }
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (ins_pos[mid] < leaf_end) lo = mid + 1;
        else hi = mid;
    }
    const int cnt = (int)(lo - j);
    inscnt[leaf] = cnt;
    ins_first[leaf] = (int32_t)j;
    touched[leaf] = 1;
    const unsigned long long slot = atomicAdd((unsigned long long*)nact_dev, 1ull);
    act[slot] = ActiveLeaf{(int32_t)leaf, cnt, (int32_t)j, 0};
}

// ---- hits / deletes in place ------------------------------------------------------------------------------------------------------
template <int ITEMS>
__global__ void __launch_bounds__(256) k_apply_hits_ilp(int64_t* __restrict__ keys, double* __restrict__ vals,
                                                         const int64_t* __restrict__ op_pos, const uint8_t* __restrict__ op_flag,
                                                         const double* __restrict__ op_val, int64_t nops, int32_t* __restrict__ leafcnt,
                                                         uint8_t* __restrict__ touched, int lgS, const int64_t* __restrict__ n_dev,
                                                         int32_t* __restrict__ ins_flag) {
    const int64_t count = n_dev ? *n_dev : nops;
    const int64_t base = (int64_t)blockIdx.x * (256 * ITEMS) + threadIdx.x;
    uint8_t f[ITEMS];
    int64_t p[ITEMS];
    double v[ITEMS];
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
        const int64_t i = base + (int64_t)k * 256;
        f[k] = 0;
        p[k] = 0;
        v[k] = 0.0;
        if (i < count) f[k] = op_flag[i];
    }
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
        const int64_t i = base + (int64_t)k * 256;
        if (f[k] == FL_OVERWRITE || f[k] == FL_DELETE) p[k] = op_pos[i];
        if (f[k] == FL_OVERWRITE) v[k] = op_val[i];
    }
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
        const int64_t i = base + (int64_t)k * 256;
        if (i >= nops) continue;
        ins_flag[i] = f[k] == FL_INSERT ? 1 : 0;
        if (f[k] == FL_OVERWRITE) {
            vals[p[k]] = v[k];
        } else if (f[k] == FL_DELETE) {
            keys[p[k]] = GAP_KEY;
            atomicSub(&leafcnt[p[k] >> lgS], 1);
            touched[p[k] >> lgS] = 1;
        }
    }
}

// ---- order-preserving compaction of the inserts --------------------------------------------------------------------------------
template <int ITEMS>
__global__ void __launch_bounds__(256) k_compact_inserts_ilp(const int64_t* __restrict__ op_key, const double* __restrict__ op_val,
                                                              const int64_t* __restrict__ op_pos, const uint8_t* __restrict__ op_flag,
                                                              const int32_t* __restrict__ ins_idx, int64_t nops,
                                                              int64_t* __restrict__ ins_key, double* __restrict__ ins_val,
                                                              int64_t* __restrict__ ins_pos, const int64_t* __restrict__ n_dev) {
    const int64_t count = n_dev ? *n_dev : nops;
    const int64_t base = (int64_t)blockIdx.x * (256 * ITEMS) + threadIdx.x;
    bool ins[ITEMS];
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
        const int64_t i = base + (int64_t)k * 256;
        ins[k] = i < count && op_flag[i] == FL_INSERT;
    }
    int32_t j[ITEMS];
    int64_t kk[ITEMS], pp[ITEMS];
    double vv[ITEMS];
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
        const int64_t i = base + (int64_t)k * 256;
        j[k] = 0;
        kk[k] = pp[k] = 0;
        vv[k] = 0.0;
        if (ins[k]) {
            j[k] = ins_idx[i];
            kk[k] = op_key[i];
            vv[k] = op_val[i];
            pp[k] = op_pos[i];
        }
    }
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
        if (ins[k]) {
            ins_key[j[k]] = kk[k];
            ins_val[j[k]] = vv[k];
            ins_pos[j[k]] = pp[k];
        }
    }
}

}  // namespace dsa
