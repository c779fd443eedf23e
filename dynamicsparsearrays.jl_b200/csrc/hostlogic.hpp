// libdsa host logic (no GPU needed): PMA geometry, integer density bounds, column-map planning.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <vector>

namespace dsa {

struct Geometry {
    int64_t capacity, segment_capacity, nb_segments, height;
    double t_d, p_d;
};

constexpr double T_H = 0.7, T_0 = 0.92, P_H = 0.3, P_0 = 0.08;   // pma.jl:58,70,87

// capacity = 2^ceil(Int, log2(ceil(n / t_h)))   (pma.jl:64,81,88)
inline int64_t capacity_for(int64_t n) {
    double c = std::ceil((double)n / T_H);
    int64_t e = (int64_t)std::ceil(std::log2(c));
    return int64_t(1) << e;
}

// _pma (pma.jl:42-55): nb_segs = 2^ceil(log2(cap / log2(cap))), seg = cap / nb_segs, height = log2(nb_segs)
inline Geometry geometry_for_capacity(int64_t capacity) {
    Geometry g;
    g.capacity = capacity;
    double lc = std::log2((double)capacity);
    g.nb_segments = int64_t(1) << (int64_t)std::ceil(std::log2((double)capacity / lc));
    g.segment_capacity = capacity / g.nb_segments;
    int64_t h = 0;
    while ((int64_t(1) << h) < g.nb_segments) ++h;
    g.height = h;
    g.t_d = (T_H - T_0) / (double)h;
    g.p_d = (P_H - P_0) / (double)h;
    return g;
}

// bulk build of n elements (pma.jl:57-84); n == 0 -> empty constructor with expected_nb_elems (pma.jl:86-91)
inline Geometry geometry_for_build(int64_t n, int64_t expected_nb_elems = 100) {
    return geometry_for_capacity(capacity_for(n > 0 ? n : expected_nb_elems));
}

// _extend! / _shrink! keep the segment capacity (pma.jl:143-161)
inline Geometry geometry_resized(const Geometry& g, int64_t capacity, int64_t height) {
    Geometry r = g;
    r.capacity = capacity;
    r.nb_segments = capacity / g.segment_capacity;
    r.height = height;
    r.t_d = (T_H - T_0) / (double)height;
    r.p_d = (P_H - P_0) / (double)height;
    return r;
}

// accept iff p_0 + p_d*h <= cnt / window_capacity <= t_0 + t_d*h   (pma.jl:119-123), as integer bounds on cnt
inline void level_bounds(int64_t S, int64_t H, double t_d, double p_d, int64_t* mn, int64_t* mx) {
    for (int64_t h = 0; h <= H; ++h) {
        const int64_t wc = (int64_t(1) << h) * S;
        const double p = P_0 + p_d * (double)h;
        const double t = T_0 + t_d * (double)h;
        int64_t lo = (int64_t)std::ceil(p * (double)wc);
        if (lo < 0) lo = 0;
        while ((double)lo / (double)wc < p) ++lo;
        while (lo > 0 && (double)(lo - 1) / (double)wc >= p) --lo;
        int64_t hi = (int64_t)std::floor(t * (double)wc);
        if (hi > wc) hi = wc;
        while (hi >= 0 && (double)hi / (double)wc > t) --hi;
        while (hi < wc && (double)(hi + 1) / (double)wc <= t) ++hi;
        mn[h] = lo;
        mx[h] = hi;
    }
}

// Capacity after a batch whose root window failed (batch policy): repeat _extend! while the root
// density exceeds t_h, or _shrink! while it is below p_h and height > 1 (pma.jl:132-139).
// single_step: the batch was ONE op -> exactly one _extend! / _shrink! like the reference's setindex! (pma.jl:132-139).
inline Geometry geometry_after_root_failure(const Geometry& g, int64_t N, bool single_step = false) {
    int64_t mn[40], mx[40];
    int64_t cap = g.capacity, h = g.height;
    level_bounds(g.segment_capacity, h, g.t_d, g.p_d, mn, mx);
    if (N > mx[h]) {
        do {
            cap *= 2;
            h += 1;
            level_bounds(g.segment_capacity, h, (T_H - T_0) / (double)h, (P_H - P_0) / (double)h, mn, mx);
        } while (N > mx[h] && !single_step);
    } else {
        while (h > 1) {
            level_bounds(g.segment_capacity, h, (T_H - T_0) / (double)h, (P_H - P_0) / (double)h, mn, mx);
            if (N >= mn[h]) break;
            cap /= 2;
            h -= 1;
            if (single_step) break;
        }
    }
    return geometry_resized(g, cap, h);
}

// Column-map planning: the final col_keys after inserting `new_keys` (distinct, absent, in first-arrival order)
// one by one with addcolumn! (pcsr.jl:148-169).  Between two live slots the tombstones form a run of T slots; the
// reference re-uses a tombstone only when the arriving key is larger than every key already inserted in that run
// (it then sits right after its predecessor, pcsr.jl:155-156), otherwise it shifts (pcsr.jl:158-163).  The result
// per run is therefore [inserted keys ascending][T - min(T, R) tombstones], R = number of left-to-right maxima of
// the arrival sequence inside the run.  Unlike the reference, tombstones to the right of a shift and re-use of a
// trailing tombstone are supported (reference bugs (i) and (ii)).
inline int64_t colmap_plan(const int64_t* slot_key, const uint8_t* slot_live, int64_t nslots, const int64_t* new_keys,
                           int64_t nnew, std::vector<int64_t>& out_key, std::vector<uint8_t>& out_live,
                           std::vector<int64_t>& out_old) {
    std::vector<int64_t> live_slot, live_key;
    for (int64_t s = 0; s < nslots; ++s)
        if (slot_live[s]) { live_slot.push_back(s); live_key.push_back(slot_key[s]); }
    const int64_t nl = (int64_t)live_slot.size();
    struct NewCol { int64_t interval, key, arrival; };
    std::vector<NewCol> nc((size_t)nnew);
    for (int64_t a = 0; a < nnew; ++a) {
        int64_t iv = std::lower_bound(live_key.begin(), live_key.end(), new_keys[a]) - live_key.begin();
        nc[a] = NewCol{iv, new_keys[a], a};
    }
    std::sort(nc.begin(), nc.end(), [](const NewCol& x, const NewCol& y) {
        if (x.interval != y.interval) return x.interval < y.interval;
        return x.key < y.key;
    });
    out_key.clear(); out_live.clear(); out_old.clear();
    size_t q = 0;
    for (int64_t iv = 0; iv <= nl; ++iv) {
        const int64_t run_begin = iv == 0 ? 0 : live_slot[iv - 1] + 1;
        const int64_t run_end = iv == nl ? nslots : live_slot[iv];   // tombstones are [run_begin, run_end)
        const int64_t T = run_end - run_begin;
        size_t q0 = q;
        while (q < nc.size() && nc[q].interval == iv) ++q;
        // left-to-right maxima of the arrival sequence == keys whose arrival precedes every larger key's arrival
        int64_t R = 0;
        int64_t suffix_min = INT64_MAX;
        for (size_t k = q; k-- > q0;) {
            if (nc[k].arrival < suffix_min) R += 1;
            suffix_min = std::min(suffix_min, nc[k].arrival);
        }
        for (size_t k = q0; k < q; ++k) { out_key.push_back(nc[k].key); out_live.push_back(1); out_old.push_back(0); }
        const int64_t remaining = T - std::min(T, R);
        for (int64_t t = 0; t < remaining; ++t) { out_key.push_back(0); out_live.push_back(0); out_old.push_back(0); }
        if (iv < nl) { out_key.push_back(live_key[iv]); out_live.push_back(1); out_old.push_back(live_slot[iv] + 1); }
    }
    return (int64_t)out_key.size();
}

}  // namespace dsa
