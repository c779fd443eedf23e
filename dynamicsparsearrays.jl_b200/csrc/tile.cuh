// libdsa — tile-streamed batched setindex! of one PCSR orientation (dense batches: a batch that touches most leaves).
//
// The random-access pipeline (per-partition buckets -> rank + locate in HBM -> apply -> leaf bookkeeping -> leaf merge) moves
// more bytes than one pass over the array once a batch touches about half of the leaves, and moves them as 32-byte sectors.
// Here the array is cut into tiles of TILE_CELLS cells and a batch is applied in two kernels:
//
//   k_tile_assign   one thread per op: column lookup (find(col_keys), pcsr.jl:342) + batch statistics, then the TILE that holds
//                   the op's predecessor cell.  A partition span that lies inside one tile names the tile by itself (two loads of
//                   the semaphore table); only spans that straddle a tile border run the gapped search (finds.jl:29-57) in HBM.
//                   The op is dropped into its tile's bucket (fixed capacity: a batch that overflows one falls back).
//   k_tile_merge    one CTA per tile with ops: the tile's cells are streamed into shared memory once; every op is located there
//                   (same gapped binary search, on shared memory), hits overwrite / blank their cell (last writer wins by arrival),
//                   misses are grouped per leaf, and every leaf whose post-batch count stays inside its own density bounds
//                   (pma.jl:119-123 with h = 0) is re-laid at its spread! positions on the spot (pack! + spread!, moves.jl:94-172).
//                   Leaves that fail their bounds hand their inserts, in order, to the density tree / window kernels of pma.cuh,
//                   exactly as the random-access pipeline does.  Only modified leaves are written back.
//
// The result is the batch policy's layout, bit for bit: a leaf accepted at its own level and later covered by a
// larger window is simply re-laid twice (the window kernel reads the merged leaf with no pending inserts).
#pragma once
#include "pcsr.cuh"

namespace dsa {

// ---------------------------------------------------------------------------------------------
// phase 1 of a tile-streamed batch: lookup + statistics (as k_col_lookup) + tile bucket
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 6) k_tile_assign(const int64_t* __restrict__ partkeys, const int64_t* __restrict__ inkeys,
                                                        const double* __restrict__ vals, int64_t n_host, const int64_t* __restrict__ n_dev,
                                                        const int64_t* __restrict__ live_keys, const int32_t* __restrict__ live_slot,
                                                        int64_t nlive, const int32_t* __restrict__ keymap, int64_t keymap_min,
                                                        int64_t keymap_len, const int64_t* __restrict__ keys, int64_t cap,
                                                        const int64_t* __restrict__ sem, const int32_t* __restrict__ next_slot,
                                                        int64_t nslots, int64_t* __restrict__ cs, int32_t* __restrict__ tcnt,
                                                        TileRec* __restrict__ rec) {
    int64_t n = n_host;
    if (n_dev) n = *n_dev < n_host ? *n_dev : n_host;
    if (blockIdx.x == 0 && threadIdx.x == 0) cs[CS_N] = n;
    int64_t mink = INT64_MAX, maxk = INT64_MIN, maxp = INT64_MIN, maxknz = INT64_MIN, minp = INT64_MAX;
    int miss = 0;
    int bmax = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t pk = partkeys[i];
        const int64_t k = inkeys[i];
        const double v = vals[i];
        minp = pk < minp ? pk : minp;
        int32_t s;
        if (keymap) {
            const int64_t r = pk - keymap_min;
            s = (r >= 0 && r < keymap_len) ? keymap[r] : -1;
        } else {
            s = live_lookup(live_keys, live_slot, nlive, pk);
        }
        mink = k < mink ? k : mink;
        maxk = k > maxk ? k : maxk;
        const bool is_set = v != 0.0;   // pcsr.jl:301
        if (is_set) {
            maxp = pk > maxp ? pk : maxp;
            maxknz = k > maxknz ? k : maxknz;
        }
        int64_t ps = -1;
        if (s >= 0) ps = sem[s];
        if (ps < 0) {   // absent column (or a slot without a placed semaphore): the host falls back to the general path
            miss += 1;
            continue;
        }
        if (k < 1) continue;   // refused by the host's validation (statistics) before anything is applied
        const int32_t ns = next_slot ? next_slot[s] : (s + 1 < nslots ? s + 1 : -1);
        const int64_t pe = ns >= 0 ? sem[ns] : cap;   // exclusive end of the span (pcsr.jl:177-186)
        const int64_t from = is_set ? ps + 1 : ps;    // inserts search (sem, end], deletes [sem, end]  (pcsr.jl:305-307)
        const int64_t to = pe - 1;
        int64_t t = ps >> TILE_LG, lo = from, hi = to;
        if (t != (to >> TILE_LG)) {   // the span straddles a tile border: find the predecessor's tile in HBM
            bool hit = false;
            int64_t pos = gapped_find(keys, k, from, to, &hit);
            if (pos < ps) pos = ps;
            t = pos >> TILE_LG;
            lo = hi = pos;
        }
        const int64_t tb = t << TILE_LG;
        const int li = atomicAdd(&tcnt[t], 1);
        bmax = li + 1 > bmax ? li + 1 : bmax;
        if (li < TILE_CAP) {
            int4* out = reinterpret_cast<int4*>(rec + t * TILE_CAP + li);
            const uint64_t ku = (uint64_t)k, vu = (uint64_t)__double_as_longlong(v);
            out[0] = make_int4((int)(uint32_t)ku, (int)(uint32_t)(ku >> 32), (int)(uint32_t)vu, (int)(uint32_t)(vu >> 32));
            out[1] = make_int4((int)(uint32_t)i, s, (int)((uint32_t)(lo - tb) | ((uint32_t)(hi - tb) << 16)), 0);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        int64_t a = __shfl_xor_sync(0xffffffffu, mink, o); mink = a < mink ? a : mink;
        a = __shfl_xor_sync(0xffffffffu, maxk, o); maxk = a > maxk ? a : maxk;
        a = __shfl_xor_sync(0xffffffffu, maxp, o); maxp = a > maxp ? a : maxp;
        a = __shfl_xor_sync(0xffffffffu, maxknz, o); maxknz = a > maxknz ? a : maxknz;
        a = __shfl_xor_sync(0xffffffffu, minp, o); minp = a < minp ? a : minp;
        miss += __shfl_xor_sync(0xffffffffu, miss, o);
        const int b2 = __shfl_xor_sync(0xffffffffu, bmax, o);
        bmax = b2 > bmax ? b2 : bmax;
    }
    __shared__ int64_t sh[5][8];
    __shared__ int shm[8], shb[8];
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) {
        sh[0][wid] = mink; sh[1][wid] = maxk; sh[2][wid] = maxp; sh[3][wid] = maxknz; sh[4][wid] = minp;
        shm[wid] = miss;
        shb[wid] = bmax;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) {
            mink = sh[0][w] < mink ? sh[0][w] : mink;
            maxk = sh[1][w] > maxk ? sh[1][w] : maxk;
            maxp = sh[2][w] > maxp ? sh[2][w] : maxp;
            maxknz = sh[3][w] > maxknz ? sh[3][w] : maxknz;
            minp = sh[4][w] < minp ? sh[4][w] : minp;
            miss += shm[w];
            bmax = shb[w] > bmax ? shb[w] : bmax;
        }
        unsigned long long* u = (unsigned long long*)cs;
        if (bmax) atomicMax((long long*)&cs[CS_MAXBUCKET], (long long)bmax);
        if (miss) atomicAdd(&u[CS_MISSING], (unsigned long long)miss);
        if (mink != INT64_MAX) atomicMax(&u[CS_MINKEY], ~cs_code(mink));
        if (maxk != INT64_MIN) atomicMax(&u[CS_MAXKEY], cs_code(maxk));
        if (maxp != INT64_MIN) atomicMax(&u[CS_MAXPART_NZ], cs_code(maxp));
        if (maxknz != INT64_MIN) atomicMax(&u[CS_MAXKEY_NZ], cs_code(maxknz));
        if (minp != INT64_MAX) atomicMax(&u[CS_MINPART], ~cs_code(minp));
    }
}

// the reference's gapped binary search (finds.jl:29-57) on the tile in shared memory: position of the hit or of the predecessor
__device__ __forceinline__ int tile_find(const int64_t* sk, int64_t key, int lo, int hi, bool* hit) {
    const int from = lo;
    while (lo <= hi) {
        const int mid = (lo + hi) >> 1;
        int i = mid;
        int64_t k = sk[i];
        while (k == GAP_KEY && i > lo) {
            --i;
            k = sk[i];
        }
        if (k == GAP_KEY) {
            lo = mid + 1;
        } else if (k > key) {
            hi = i - 1;
        } else if (k < key) {
            lo = mid + 1;
        } else {
            *hit = true;
            return i;
        }
    }
    (void)from;
    *hit = false;
    int i = hi;
    while (i > 0 && sk[i] == GAP_KEY) --i;   // finds.jl:49-56 (the predecessor lies in this tile by construction)
    return i < 0 ? 0 : i;
}

__device__ __forceinline__ unsigned mask_le(int x) { return x >= 31 ? 0xffffffffu : ((2u << x) - 1u); }

struct TileSmem {
    int64_t sk[TILE_CELLS];
    double sv[TILE_CELLS];
    int64_t rkey[TILE_CAP];
    double rval[TILE_CAP];
    uint32_t claim[TILE_CELLS];   // per cell: 1 + arrival of the last op that hits it
    uint32_t rarr[TILE_CAP];
    int32_t rslot[TILE_CAP];
    int lcnt[TILE_MAX_LEAVES];    // per leaf: ops that miss (insert candidates and deletes of absent keys)
    int ndel[TILE_MAX_LEAVES];    // per leaf: cells blanked
    uint16_t rpos[TILE_CAP];      // tile-local position of the hit / predecessor; bit 15 = hit
    uint16_t rli[TILE_CAP];       // index inside the leaf's list, later the merged rank
    uint16_t llist[TILE_CAP];     // misses grouped by leaf
    uint16_t loff[TILE_MAX_LEAVES];
    uint16_t alist[TILE_MAX_LEAVES];   // leaves with misses
    uint8_t rstat[TILE_CAP];      // 1 = live insert
    uint8_t dvf[TILE_MAX_LEAVES]; // leaf has an overwritten value
    uint8_t mrg[TILE_MAX_LEAVES]; // leaf was re-laid
    int wtot[8];
    int nactive;
};

struct TileArgs {
    int64_t* keys;
    double* vals;
    const TileRec* rec;
    const int32_t* tcnt;
    int64_t* sem;
    const uint8_t* destpos;
    int32_t* leafcnt;
    uint8_t* touched;
    int32_t* inscnt;
    int32_t* ins_first;
    int64_t* ins_key;
    double* ins_val;
    int64_t* ins_pos;
    int64_t* status;
};

__global__ void __launch_bounds__(TILE_THREADS, 3) k_tile_merge(TileArgs A, Levels L) {
    extern __shared__ __align__(16) unsigned char tile_smem_raw[];
    TileSmem& s = *reinterpret_cast<TileSmem*>(tile_smem_raw);
    const int t = blockIdx.x;
    int nrec = A.tcnt[t];
    if (nrec <= 0) return;
    if (nrec > TILE_CAP) nrec = TILE_CAP;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t tbase = (int64_t)t << TILE_LG;
    const int lgS = L.lgS, S = 1 << lgS, NL = TILE_CELLS >> lgS;
    constexpr unsigned FULL = 0xffffffffu;

    // ---- A: the tile's cells, once ----------------------------------------------------------------------------------
    {
        const longlong2* gk = reinterpret_cast<const longlong2*>(A.keys + tbase);
        const double2* gv = reinterpret_cast<const double2*>(A.vals + tbase);
        longlong2* dk = reinterpret_cast<longlong2*>(s.sk);
        double2* dv = reinterpret_cast<double2*>(s.sv);
#pragma unroll
        for (int i = 0; i < TILE_CELLS / 2 / TILE_THREADS; ++i) {
            dk[tid + i * TILE_THREADS] = gk[tid + i * TILE_THREADS];
            dv[tid + i * TILE_THREADS] = gv[tid + i * TILE_THREADS];
        }
        for (int i = tid; i < TILE_CELLS; i += TILE_THREADS) s.claim[i] = 0;
        for (int i = tid; i < NL; i += TILE_THREADS) {
            s.lcnt[i] = 0;
            s.ndel[i] = 0;
            s.dvf[i] = 0;
            s.mrg[i] = 0;
        }
    }
    __syncthreads();

    // ---- B: locate every op in shared memory ------------------------------------------------------------------------
    for (int j = tid; j < nrec; j += TILE_THREADS) {
        const int4* rp = reinterpret_cast<const int4*>(A.rec + (int64_t)t * TILE_CAP + j);
        const int4 a = rp[0], b = rp[1];
        const int64_t key = (int64_t)((uint64_t)(uint32_t)a.x | ((uint64_t)(uint32_t)a.y << 32));
        const double val = __longlong_as_double((long long)((uint64_t)(uint32_t)a.z | ((uint64_t)(uint32_t)a.w << 32)));
        const uint32_t arr = (uint32_t)b.x;
        const int lo = (int)((uint32_t)b.z & 0xffffu), hi = (int)((uint32_t)b.z >> 16);
        bool hit = false;
        const int pos = tile_find(s.sk, key, lo, hi, &hit);
        s.rkey[j] = key;
        s.rval[j] = val;
        s.rarr[j] = arr;
        s.rslot[j] = b.y;
        s.rpos[j] = (uint16_t)(pos | (hit ? 0x8000 : 0));
        if (hit) atomicMax(&s.claim[pos], arr + 1u);
        else s.rli[j] = (uint16_t)atomicAdd(&s.lcnt[pos >> lgS], 1);
    }
    __syncthreads();

    // ---- C: hits (the last arrival wins, writes.jl:16-19 / 65-68) + offsets of the per-leaf miss lists --------------
    for (int j = tid; j < nrec; j += TILE_THREADS) {
        const int pp = s.rpos[j];
        if (pp & 0x8000) {
            const int pos = pp & 0x7fff;
            if (s.claim[pos] == s.rarr[j] + 1u) {
                const double v = s.rval[j];
                if (v != 0.0) {
                    s.sv[pos] = v;
                    s.dvf[pos >> lgS] = 1;
                } else {
                    s.sk[pos] = GAP_KEY;
                    atomicAdd(&s.ndel[pos >> lgS], 1);
                }
            }
        }
    }
    {   // block scan of (misses, leaf has misses) packed in one int; NL <= TILE_THREADS
        const int x = tid < NL ? s.lcnt[tid] : 0;
        const int mine = x | ((x > 0 ? 1 : 0) << 16);
        int inc = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(FULL, inc, o);
            if (lane >= o) inc += y;
        }
        if (lane == 31) s.wtot[warp] = inc;
        __syncthreads();
        int before = 0;
        for (int w = 0; w < warp; ++w) before += s.wtot[w];
        const int excl = before + inc - mine;
        if (tid < NL) {
            s.loff[tid] = (uint16_t)(excl & 0xffff);
            if (x > 0) s.alist[excl >> 16] = (uint16_t)tid;
        }
        if (tid == TILE_THREADS - 1) s.nactive = (before + inc) >> 16;
    }
    __syncthreads();

    // ---- D: misses grouped by leaf -----------------------------------------------------------------------------------
    for (int j = tid; j < nrec; j += TILE_THREADS) {
        const int pp = s.rpos[j];
        if (!(pp & 0x8000)) s.llist[s.loff[pp >> lgS] + s.rli[j]] = (uint16_t)j;
    }
    __syncthreads();

    // ---- E: one S-lane group per leaf with misses (32/S leaves per warp; control flow is warp-uniform) ---------------
    {
        const int G = 32 >> lgS, grp = lane >> lgS, q = lane & (S - 1), gshift = grp << lgS;
        const unsigned gmask = S >= 32 ? 0xffffffffu : ((1u << S) - 1u);
        const int mn0 = (int)L.mn[0], mx0 = (int)L.mx[0];
        const int nact = s.nactive;
        for (int a0 = warp * G; a0 < nact; a0 += (TILE_THREADS / 32) * G) {
            const int a = a0 + grp;
            const bool have = a < nact;
            const int l = have ? s.alist[a] : 0;
            const int n_l = have ? s.lcnt[l] : 0;
            const int base = s.loff[l];
            const int cell = (l << lgS) + q;
            const int64_t ck = s.sk[cell];   // after the deletes and overwrites of phase C
            const double cv = s.sv[cell];
            const unsigned lm = (__ballot_sync(FULL, have && ck != GAP_KEY) >> gshift) & gmask;
            const int maxn = __reduce_max_sync(FULL, n_l);
            // sweep 1: last writer wins among the misses of one key; the surviving non-zero writes are the inserts
            int nins = 0;
            for (int j0 = 0; j0 < maxn; j0 += S) {
                const int j = j0 + q;
                const bool v = j < n_l;
                const int idx = v ? s.llist[base + j] : 0;
                const int64_t key = s.rkey[idx];
                const int32_t slot = s.rslot[idx];
                const uint32_t arr = s.rarr[idx];
                bool dead = false;
                for (int o = 0; o < maxn; ++o) {
                    if (o < n_l) {
                        const int io = s.llist[base + o];
                        dead |= (s.rkey[io] == key && s.rslot[io] == slot && s.rarr[io] > arr);
                    }
                }
                const bool ins = v && !dead && s.rval[idx] != 0.0;
                if (v) s.rstat[idx] = ins ? 1 : 0;
                nins += __popc((__ballot_sync(FULL, ins) >> gshift) & gmask);
            }
            __syncwarp();
            if (!__any_sync(FULL, nins > 0)) continue;
            const int m = __popc(lm) + nins;
            const bool acc = nins > 0 && m >= mn0 && m <= mx0;   // the leaf is its own window (pma.jl:119-123, h = 0)
            const bool left = nins > 0 && !acc;                   // a larger window (or a resize) takes the leaf's inserts
            unsigned long long gb = 0;
            if (left && q == 0) gb = atomicAdd((unsigned long long*)&A.status[ST_NINS], (unsigned long long)nins);
            gb = __shfl_sync(FULL, gb, gshift);
            // sweep 2: order of the inserts (predecessor cell, key) -> merged rank / slot in the hand-over arrays
            unsigned insmask = 0;
            for (int j0 = 0; j0 < maxn; j0 += S) {
                const int j = j0 + q;
                const bool v = j < n_l;
                const int idx = v ? s.llist[base + j] : 0;
                const bool ins = v && s.rstat[idx];
                const int pp = s.rpos[idx] & 0x7fff;
                const int64_t key = s.rkey[idx];
                int rk = 0;
                for (int o = 0; o < maxn; ++o) {
                    if (o < n_l) {
                        const int io = s.llist[base + o];
                        if (s.rstat[io]) {
                            const int po = s.rpos[io] & 0x7fff;
                            rk += (po < pp) || (po == pp && s.rkey[io] < key);
                        }
                    }
                }
                unsigned bit = 0;
                if (ins && acc) {
                    const int R = __popc(lm & mask_le(pp & (S - 1))) + rk;   // survivors up to the predecessor + earlier inserts
                    s.rli[idx] = (uint16_t)R;
                    bit = 1u << R;
                }
                if (ins && left) {
                    const int64_t g = (int64_t)gb + rk;
                    A.ins_key[g] = key;
                    A.ins_val[g] = s.rval[idx];
                    A.ins_pos[g] = tbase + pp;
                }
                insmask |= (__reduce_or_sync(FULL, bit << gshift) >> gshift) & gmask;
            }
            __syncwarp();
            // merge: survivor with rank r among the survivors takes the r-th merged rank not taken by an insert
            const int mm = acc ? m : 0;
            const unsigned mask = L.leafmask[mm];
            const uint8_t* __restrict__ dtab = A.destpos + mm * 32;
            const bool isurv = acc && ck != GAP_KEY;
            const int srank = __popc(lm & ((1u << q) - 1u));
            int R = srank;
            if (isurv) {
                while (true) {
                    const int Rn = srank + __popc(insmask & mask_le(R));
                    if (Rn == R) break;
                    R = Rn;
                }
            }
            __syncwarp();
            if (acc && !((mask >> q) & 1u)) {
                s.sk[cell] = GAP_KEY;
                s.sv[cell] = 0.0;
            }
            if (isurv) {
                const int d = (l << lgS) + dtab[R];
                s.sk[d] = ck;
                s.sv[d] = cv;
            }
            for (int j0 = 0; j0 < maxn; j0 += S) {
                const int j = j0 + q;
                const bool v = j < n_l;
                const int idx = v ? s.llist[base + j] : 0;
                if (v && acc && s.rstat[idx]) {
                    const int d = (l << lgS) + dtab[s.rli[idx]];
                    s.sk[d] = s.rkey[idx];
                    s.sv[d] = s.rval[idx];
                }
            }
            if (have && q == 0 && nins > 0) {
                if (acc) {
                    s.mrg[l] = 1;
                } else {
                    const int64_t lg = (int64_t)t * NL + l;
                    A.inscnt[lg] = nins;
                    A.ins_first[lg] = (int32_t)gb;
                    A.touched[lg] = 1;
                }
            }
            __syncwarp();
        }
    }
    __syncthreads();

    // ---- F: modified leaves go back (keys if cells were blanked or re-laid, values if overwritten or re-laid) --------
    {
        const int G = 32 >> lgS, grp = lane >> lgS, q = lane & (S - 1), gshift = grp << lgS;
        const unsigned gmask = S >= 32 ? 0xffffffffu : ((1u << S) - 1u);
        for (int l0 = warp * G; l0 < NL; l0 += (TILE_THREADS / 32) * G) {
            const int l = l0 + grp;
            const bool merged = s.mrg[l] != 0;
            const bool dK = merged || s.ndel[l] > 0;
            const bool dV = merged || s.dvf[l] != 0;
            const int cell = (l << lgS) + q;
            const int64_t k = s.sk[cell];
            const int64_t p = tbase + cell;
            if (dK) A.keys[p] = k;
            if (dV) {
                const double v = s.sv[cell];
                A.vals[p] = v;
                if (merged && k == 0) A.sem[(int64_t)v - 1] = p;   // moves.jl:160-166
            }
            const unsigned b = (__ballot_sync(FULL, k != GAP_KEY) >> gshift) & gmask;
            if (dK && q == 0) {
                const int64_t lg = (int64_t)t * NL + l;
                A.leafcnt[lg] = __popc(b);
                A.touched[lg] = 1;
            }
        }
    }
}

}  // namespace dsa
