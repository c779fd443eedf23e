// EXPERIMENTAL (DSA_ILP=2|4, see ilp.cuh): several ops per thread for the column lookup and the bucket scatter.
#pragma once

namespace dsa {

template <int ITEMS>
__global__ void __launch_bounds__(256) k_col_lookup_ilp(const int64_t* __restrict__ partkeys, const int64_t* __restrict__ inkeys,
                                                         const double* __restrict__ vals, int64_t n,
                                                         const int64_t* __restrict__ live_keys, const int32_t* __restrict__ live_slot,
                                                         int64_t nlive, const int32_t* __restrict__ keymap, int64_t keymap_min,
                                                         int64_t keymap_len, int32_t* __restrict__ op_slot, int64_t* __restrict__ cs,
                                                         int32_t* __restrict__ bcnt, int32_t* __restrict__ lidx) {
    const int64_t base = (int64_t)blockIdx.x * (256 * ITEMS) + threadIdx.x;
    int64_t mink = INT64_MAX, maxk = INT64_MIN, maxp = INT64_MIN, maxknz = INT64_MIN, minp = INT64_MAX;
    int miss = 0;
    int bmax = 0;
    int64_t pk[ITEMS], ik[ITEMS];
    double v[ITEMS];
    int32_t s[ITEMS];
    bool in[ITEMS];
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {   // the op's fields: independent loads
        const int64_t i = base + (int64_t)k * 256;
        in[k] = i < n;
        pk[k] = ik[k] = 0;
        v[k] = 0.0;
        if (in[k]) {
            pk[k] = partkeys[i];
            if (inkeys) ik[k] = inkeys[i];
            if (inkeys && vals) v[k] = vals[i];
        }
    }
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {   // partition key -> slot
        s[k] = -1;
        if (!in[k]) continue;
        if (keymap) {
            const int64_t r = pk[k] - keymap_min;
            s[k] = (r >= 0 && r < keymap_len) ? keymap[r] : -1;
        } else {
            s[k] = live_lookup(live_keys, live_slot, nlive, pk[k]);
        }
    }
    int li[ITEMS];
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {   // bucket bookkeeping: ITEMS atomics in flight
        li[k] = -1;
        if (in[k] && bcnt && s[k] >= 0) li[k] = atomicAdd(&bcnt[s[k]], 1);
    }
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
        if (!in[k]) continue;
        const int64_t i = base + (int64_t)k * 256;
        op_slot[i] = s[k];
        miss += s[k] < 0;
        minp = pk[k] < minp ? pk[k] : minp;
        if (li[k] >= 0) {
            lidx[i] = li[k];
            bmax = li[k] + 1 > bmax ? li[k] + 1 : bmax;
        }
        if (inkeys) {
            mink = ik[k] < mink ? ik[k] : mink;
            maxk = ik[k] > maxk ? ik[k] : maxk;
            if (vals && v[k] != 0.0) {
                maxp = pk[k] > maxp ? pk[k] : maxp;
                maxknz = ik[k] > maxknz ? ik[k] : maxknz;
            }
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
        if (bmax) atomicMax((long long*)&cs[CS_MAXBUCKET], (long long)bmax);
        if (miss) atomicAdd((unsigned long long*)&cs[CS_MISSING], (unsigned long long)miss);
        if (mink != INT64_MAX) atomicMin((long long*)&cs[CS_MINKEY], (long long)mink);
        if (maxk != INT64_MIN) atomicMax((long long*)&cs[CS_MAXKEY], (long long)maxk);
        if (maxp != INT64_MIN) atomicMax((long long*)&cs[CS_MAXPART_NZ], (long long)maxp);
        if (maxknz != INT64_MIN) atomicMax((long long*)&cs[CS_MAXKEY_NZ], (long long)maxknz);
        if (minp != INT64_MAX) atomicMin((long long*)&cs[CS_MINPART], (long long)minp);
    }
}

template <int ITEMS>
__global__ void __launch_bounds__(256) k_bucket_scatter_ilp(const int32_t* __restrict__ op_slot, const int32_t* __restrict__ lidx,
                                                             const int64_t* __restrict__ inkeys, int64_t n, const int32_t* __restrict__ boff,
                                                             BucketRec* __restrict__ rec) {
    const int64_t base = (int64_t)blockIdx.x * (256 * ITEMS) + threadIdx.x;
    int32_t s[ITEMS], li[ITEMS];
    int64_t ik[ITEMS];
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
        const int64_t i = base + (int64_t)k * 256;
        s[k] = -1;
        li[k] = 0;
        ik[k] = 0;
        if (i < n) {
            s[k] = op_slot[i];
            li[k] = lidx[i];
            ik[k] = inkeys[i];
        }
    }
    int32_t bo[ITEMS];
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
        bo[k] = 0;
        if (s[k] >= 0) bo[k] = boff[s[k]];
    }
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
        const int64_t i = base + (int64_t)k * 256;
        if (s[k] >= 0) rec[(int64_t)bo[k] + li[k]] = BucketRec{ik[k], (uint32_t)i, s[k]};
    }
}

// in-bucket ranking, ITEMS ops per thread: the four dependent rounds (own record -> bucket bounds -> bucket scan -> value gather)
// of ITEMS ops overlap.  Same outputs as k_bucket_rank.
template <int ITEMS>
__global__ void __launch_bounds__(256) k_bucket_rank_ilp(const BucketRec* __restrict__ rec, const int32_t* __restrict__ boff,
                                                          const int32_t* __restrict__ bcnt, int64_t n, const double* __restrict__ vals,
                                                          int32_t* __restrict__ u_pid, int64_t* __restrict__ u_key, double* __restrict__ u_val,
                                                          uint8_t* __restrict__ u_dead) {
    const int64_t base = (int64_t)blockIdx.x * (256 * ITEMS) + threadIdx.x;
    BucketRec me[ITEMS];
    bool in[ITEMS];
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
        const int64_t p = base + (int64_t)k * 256;
        in[k] = p < n;
        me[k] = BucketRec{0, 0u, 0};
        if (in[k]) me[k] = rec[p];
    }
    int64_t lo[ITEMS], hi[ITEMS];
    double v[ITEMS];
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
        lo[k] = hi[k] = 0;
        v[k] = 0.0;
        if (in[k]) {
            lo[k] = boff[me[k].slot];
            hi[k] = lo[k] + bcnt[me[k].slot];
            v[k] = vals[me[k].arr];          // independent of the ranking: issued now, consumed at the end
        }
    }
    int64_t r[ITEMS];
    bool dead[ITEMS];
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
        r[k] = lo[k];
        dead[k] = false;
        for (int64_t q = lo[k]; q < hi[k]; ++q) {
            const BucketRec o = rec[q];
            r[k] += (o.key < me[k].key) || (o.key == me[k].key && o.arr < me[k].arr);
            dead[k] |= (o.key == me[k].key && o.arr > me[k].arr);
        }
    }
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
        if (!in[k]) continue;
        u_pid[r[k]] = me[k].slot;
        u_key[r[k]] = me[k].key;
        u_val[r[k]] = v[k];
        u_dead[r[k]] = dead[k] ? 1 : 0;
    }
}

}  // namespace dsa
