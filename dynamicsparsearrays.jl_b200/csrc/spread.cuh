// Closed form of the reference's spread! (moves.jl:120-172) for data-parallel use.
//
// spread! walks a window of c cells right-to-left and leaves the e = c - m gaps at the 1-based
// window offsets floor(k * (c / e)), k = 1..e, where c / e and k * (c / e) are evaluated in
// Float64 (moves.jl:121-131).  With G(k) = floor(fl(k * fl(c / e))) (strictly increasing for
// every window the PMA ever spreads, because c / e >= 1 / (1 - p_0) > 1.08), the m elements
// fill the complement in order, hence
//     dest(r)   = r + #{k : G(k) - k <= r}            (0-based offset of the element of rank r)
//     gaps<=(p) = #{k : G(k) <= p + 1},  rank(p) = p - gaps<=(p),  p is a gap iff G(gaps<=(p)) == p + 1.
// Both counts are monotone in k, so an O(1) estimate plus an exact fix-up loop evaluates them
// with the reference's own IEEE operations (division and multiplication rounded to nearest,
// never contracted into an FMA, no reciprocal).
#pragma once
#include <cstdint>
#include <cmath>

#if defined(__CUDACC__)
#define DSA_HD __host__ __device__ __forceinline__
#else
#define DSA_HD inline
#endif

namespace dsa {

struct Spread {
    int64_t c;   // window capacity
    int64_t m;   // elements
    int64_t e;   // gaps
    double f;    // fl(c / e)   (moves.jl:123)
};

DSA_HD Spread spread_make(int64_t c, int64_t m) {
    Spread s;
    s.c = c;
    s.m = m;
    s.e = c - m;
#if defined(__CUDA_ARCH__)
    s.f = s.e > 0 ? __ddiv_rn((double)c, (double)s.e) : 0.0;
#else
    s.f = s.e > 0 ? (double)c / (double)s.e : 0.0;
#endif
    return s;
}

// G(k) = floor(k * f), 1 <= k <= e    (moves.jl:124,130)
DSA_HD int64_t spread_G(const Spread& s, int64_t k) {
#if defined(__CUDA_ARCH__)
    return (int64_t)floor(__dmul_rn((double)k, s.f));
#else
    volatile double prod = (double)k * s.f;   // volatile: keep the product rounded to double before floor
    return (int64_t)std::floor(prod);
#endif
}

// number of gaps placed before the element of rank r: #{k in [1,e] : G(k) - k <= r}
DSA_HD int64_t spread_gaps_before_rank(const Spread& s, int64_t r) {
    if (s.e <= 0) return 0;
    // G(k) - k ~ k * m / e  =>  k ~ (r + 1) * e / m
    int64_t k = (int64_t)(((double)(r + 1) * (double)s.e) / (double)s.m);
    if (k < 0) k = 0;
    if (k > s.e) k = s.e;
    while (k < s.e && spread_G(s, k + 1) - (k + 1) <= r) ++k;
    while (k > 0 && spread_G(s, k) - k > r) --k;
    return k;
}

DSA_HD int64_t spread_dest(const Spread& s, int64_t r) { return r + spread_gaps_before_rank(s, r); }

// number of gaps at 0-based offsets <= p: #{k in [1,e] : G(k) <= p + 1}
DSA_HD int64_t spread_gaps_upto(const Spread& s, int64_t p) {
    if (s.e <= 0) return 0;
    int64_t k = (int64_t)((double)(p + 1) / s.f);
    if (k < 0) k = 0;
    if (k > s.e) k = s.e;
    while (k < s.e && spread_G(s, k + 1) <= p + 1) ++k;
    while (k > 0 && spread_G(s, k) > p + 1) --k;
    return k;
}

// rank of the element at 0-based offset p, or -1 if p is a gap
DSA_HD int64_t spread_rank_at(const Spread& s, int64_t p) {
    if (s.e <= 0) return p;
    int64_t u = spread_gaps_upto(s, p);
    if (u > 0 && spread_G(s, u) == p + 1) return -1;
    return p - u;
}

// number of elements stored at offsets in [a, b)
DSA_HD int64_t spread_count_range(const Spread& s, int64_t a, int64_t b) {
    int64_t ga = a > 0 ? spread_gaps_upto(s, a - 1) : 0;
    int64_t gb = b > 0 ? spread_gaps_upto(s, b - 1) : 0;
    return (b - a) - (gb - ga);
}

}  // namespace dsa
