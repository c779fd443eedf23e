// TEST INFRASTRUCTURE — NOT PRODUCT CODE.  See dsa_oracle.hpp for the contract.
// Literal restatement of /root/reference/src/*.jl (file:line cited per function).
#include "dsa_oracle.hpp"
#include <numeric>
#include <cassert>

namespace orc {

static inline int64_t ilog2_exact(int64_t x) {  // Int(log2(x)) for a power of two
    int64_t r = 0;
    while ((int64_t(1) << r) < x) ++r;
    return r;
}

// ============================ utils.jl =========================================
// utils.jl:3-10
int64_t nextemptypos(const Elements& a, int64_t from) {
    int64_t pos = from + 1;
    const int64_t len = a.length();
    while (pos <= len) {
        if (a.empty_at(pos)) return pos;
        pos += 1;
    }
    return 0;
}
// utils.jl:12-19 on a semaphore / col_keys vector (0 / !live encodes nothing)
static int64_t nextnonemptypos_sem(const Semaphores& s, int64_t from) {
    int64_t pos = from + 1;
    while (pos <= (int64_t)s.size()) {
        if (s[pos - 1] != 0) return pos;
        pos += 1;
    }
    return 0;
}
// utils.jl:21-28
int64_t previousemptypos(const Elements& a, int64_t from) {
    int64_t pos = from - 1;
    while (pos >= 1) {
        if (a.empty_at(pos)) return pos;
        pos -= 1;
    }
    return 0;
}
// utils.jl:48-58 (from included, to excluded)
int64_t nbcells(const Elements& a, int64_t from, int64_t to) {
    if (!(1 <= from && to <= a.length() + 1)) throw Error{ERR_ASSERT, "_nbcells: 1 <= from && to <= length+1"};
    if (from >= to) return 0;
    int64_t n = 0;
    for (int64_t pos = from; pos <= to - 1; ++pos)
        if (!a.empty_at(pos)) n += 1;
    return n;
}

// ============================ finds.jl =========================================
// finds.jl:29-57.  Returns the position only (the element is a.at(pos)); 0 <=> (0, nothing).
int64_t find(const Elements& a, int64_t key, int64_t from, int64_t to) {
    while (from <= to) {
        int64_t mid = (from + to) / 2;   // positions are >= 1 so ÷ == /
        int64_t i = mid;
        while (i >= from && a.empty_at(i)) i -= 1;
        if (i < from) {
            from = mid + 1;
        } else {
            int64_t curkey = a.at(i).key;
            if (curkey > key) {
                to = i - 1;
            } else if (curkey < key) {
                from = mid + 1;
            } else {
                return i;
            }
        }
    }
    int64_t i = to;
    while (i > 0 && a.empty_at(i)) i -= 1;
    if (i > 0) return i;
    return 0;
}

// Same algorithm over col_keys::Vector{Union{Nothing,L}} (generic _getkey utils.jl:30-35; 2-arg find finds.jl:59)
int64_t colkeys_find(const ColKeys& ck, int64_t col, bool* exact) {
    int64_t from = 1, to = ck.length();
    *exact = false;
    while (from <= to) {
        int64_t mid = (from + to) / 2;
        int64_t i = mid;
        while (i >= from && !ck.live[i - 1]) i -= 1;
        if (i < from) {
            from = mid + 1;
        } else {
            int64_t curkey = ck.key[i - 1];
            if (curkey > col) to = i - 1;
            else if (curkey < col) from = mid + 1;
            else { *exact = true; return i; }
        }
    }
    int64_t i = to;
    while (i > 0 && !ck.live[i - 1]) i -= 1;
    if (i > 0) return i;
    return 0;
}

// ============================ moves.jl =========================================
static inline void sem_track(const Elements& a, int64_t i, int64_t newpos, Semaphores* sem) {
    // moves.jl:31-37 / 74-80 / 160-166 : semaphores[Int(val)] = newpos for cells whose key is the semaphore key
    if (sem && !a.empty_at(i) && a.at(i).key == 0) {
        int64_t id = (int64_t)a.at(i).val;
        (*sem)[id - 1] = newpos;
    }
}
// moves.jl:7-42
void movecellstoright(Elements& a, int64_t from, int64_t to, Semaphores* sem) {
    const int64_t len = a.length();
    // Julia evaluates array[to] first: out-of-range `to` raises BoundsError from the indexing itself.
    if (!(1 <= to && to <= len)) throw Error{ERR_BOUNDS, "cannot access array at index [to]"};
    if (!a.empty_at(to)) throw Error{ERR_ARGUMENT, "The cell erased by the movement must contain nothing."};
    if (!(1 <= from && from <= len)) throw Error{ERR_BOUNDS, "cannot access array at index [from]"};
    int64_t i = to;
    while (i > from) {
        i -= 1;
        sem_track(a, i, i + 1, sem);
        a.copy_cell(i + 1, i);
    }
    a.set_nothing(i);
}
// moves.jl:50-85
void movecellstoleft(Elements& a, int64_t from, int64_t to, Semaphores* sem) {
    const int64_t len = a.length();
    if (!(1 <= to && to <= len)) throw Error{ERR_BOUNDS, "cannot access array at index [to]"};
    if (!a.empty_at(to)) throw Error{ERR_ARGUMENT, "The cell erased by the movement must contain nothing."};
    if (!(1 <= from && from <= len)) throw Error{ERR_BOUNDS, "cannot access array at index [from]"};
    int64_t i = to;
    while (i < from) {
        i += 1;
        sem_track(a, i, i - 1, sem);
        a.copy_cell(i - 1, i);
    }
    a.set_nothing(i);
}
// moves.jl:94-110  (the scan is NOT bounded by window_end)
void pack(Elements& a, int64_t window_start, int64_t /*window_end*/, int64_t m) {
    int64_t i = window_start;
    int64_t j = window_start;
    while (i < window_start + m) {
        if (a.empty_at(j)) { j += 1; continue; }
        if (i < j) {
            a.copy_cell(i, j);
            a.set_nothing(j);
        }
        i += 1;
        j += 1;
    }
}
// moves.jl:120-140.  Float64 arithmetic exactly as written: capacity / nb_empty_cells,
// window_start + floor(nb_empty_cells * freq) - 1 (all Float64; NaN when nb_empty == 0).
void spread4(Elements& a, int64_t window_start, int64_t window_end, int64_t m) {
    int64_t capacity = window_end - window_start + 1;
    int64_t nb_empty_cells = capacity - m;
    double empty_cell_freq = (double)capacity / (double)nb_empty_cells;
    double next_empty_cell = (double)window_start + std::floor((double)nb_empty_cells * empty_cell_freq) - 1.0;
    int64_t i = window_start + m - 1;
    int64_t j = window_end;
    while (i != j && i >= window_start) {
        if ((double)j == next_empty_cell) {
            nb_empty_cells -= 1;
            next_empty_cell = (double)window_start + std::floor((double)nb_empty_cells * empty_cell_freq) - 1.0;
            j -= 1;
        } else {
            a.copy_cell(j, i);
            a.set_nothing(i);
            i -= 1;
            j -= 1;
        }
    }
}
// moves.jl:142-172
void spread5(Elements& a, int64_t window_start, int64_t window_end, int64_t m, Semaphores* sem) {
    int64_t capacity = window_end - window_start + 1;
    int64_t nb_empty_cells = capacity - m;
    double empty_cell_freq = (double)capacity / (double)nb_empty_cells;
    double next_empty_cell = (double)window_start + std::floor((double)nb_empty_cells * empty_cell_freq) - 1.0;
    int64_t i = window_start + m - 1;
    int64_t j = window_end;
    while (i >= window_start) {
        if ((double)j == next_empty_cell) {
            nb_empty_cells -= 1;
            next_empty_cell = (double)window_start + std::floor((double)nb_empty_cells * empty_cell_freq) - 1.0;
            j -= 1;
        } else {
            if (i != j) {
                a.copy_cell(j, i);
                a.set_nothing(i);
            }
            sem_track(a, j, j, sem);
            i -= 1;
            j -= 1;
        }
    }
}

// ============================ writes.jl ========================================
// writes.jl:26-43
PosFlag insert_after(Elements& a, int64_t key, double value, int64_t pos, Semaphores* sem) {
    int64_t insertion_pos = pos;
    int64_t next_empty_pos = nextemptypos(a, pos);
    if (next_empty_pos != 0) {
        movecellstoright(a, pos + 1, next_empty_pos, sem);
        a.set(pos + 1, key, value);
        insertion_pos += 1;
    } else {
        int64_t previous_empty_pos = previousemptypos(a, pos);
        if (previous_empty_pos != 0) {
            movecellstoleft(a, pos, previous_empty_pos, sem);
            a.set(pos, key, value);
        } else {
            throw Error{ERR_ERROR, "No empty cell to insert a new element."};
        }
    }
    return PosFlag{insertion_pos, true};
}
// writes.jl:14-23
PosFlag insert(Elements& a, int64_t key, double value, int64_t from, int64_t to, Semaphores* sem) {
    int64_t pos = find(a, key, from, to);
    if (pos != 0 && a.at(pos).key == key && from <= pos && pos <= to) {
        a.set(pos, key, value);
        return PosFlag{pos, false};
    }
    return insert_after(a, key, value, pos, sem);
}
// writes.jl:57-68
PosFlag del(Elements& a, int64_t key, int64_t from, int64_t to) {
    int64_t pos = find(a, key, from, to);
    if (pos != 0 && a.at(pos).key == key) {
        a.set_nothing(pos);
        return PosFlag{pos, true};
    }
    return PosFlag{0, false};
}
// writes.jl:80-92
PurgeRes purge(Elements& a, int64_t from, int64_t to) {
    if (to < from) return PurgeRes{0, 0};
    int64_t nb = 0;
    for (int64_t pos = from; pos <= to; ++pos) {
        if (!a.empty_at(pos)) {
            a.set_nothing(pos);
            nb += 1;
        }
    }
    int64_t mid = from + (to - from) / 2;
    return PurgeRes{mid, nb};
}

// pma.jl:236-260
bool arrays_equal(const Elements& a1, const Elements& a2) {
    int64_t i = 1, j = 1;
    const int64_t l1 = a1.length(), l2 = a2.length();
    while (i <= l1 || j <= l2) {
        while ((i <= l1 && a1.empty_at(i)) || (j <= l2 && a2.empty_at(j))) {
            if (i <= l1 && a1.empty_at(i)) i += 1;
            if (j <= l2 && a2.empty_at(j)) j += 1;
            if (i == l1 + 1 && j == l2 + 1) break;
        }
        if (i > l1 && j <= l2) return false;
        if (i <= l1 && j > l2) return false;
        if (i <= l1 && j <= l2) {
            const KV& x = a1.at(i);
            const KV& y = a2.at(j);
            if (!(x.key == y.key && x.val == y.val)) return false;
        }
        i += 1;
        j += 1;
    }
    return true;
}

// ============================ pma.jl ===========================================
// capacity = 2^ceil(Int, log2(ceil(n / t_h)))    (pma.jl:64,81,88)
static int64_t capacity_for(int64_t n, double t_h) {
    double c = std::ceil((double)n / t_h);
    int64_t e = (int64_t)std::ceil(std::log2(c));
    return int64_t(1) << e;
}
// pma.jl:94-103
void pma_even_rebalance(Pma& p, int64_t ws, int64_t we, int64_t m) {
    int64_t capacity = we - ws + 1;
    if (capacity == p.segment_capacity) return;
    pack(p.array, ws, we, m);
    spread4(p.array, ws, we, m);
}
// pma.jl:42-55
void pma_init(Pma& p, int64_t nb_elements) {
    int64_t capacity = p.array.length();
    double lc = std::log2((double)capacity);
    int64_t nb_segs = int64_t(1) << (int64_t)std::ceil(std::log2((double)capacity / lc));
    int64_t seg_capacity = capacity / nb_segs;
    int64_t height = ilog2_exact(nb_segs);
    p.capacity = capacity;
    p.segment_capacity = seg_capacity;
    p.nb_segments = nb_segs;
    p.nb_elements = nb_elements;
    p.height = height;
    p.t_h = 0.7; p.t_0 = 0.92; p.p_h = 0.3; p.p_0 = 0.08;
    p.t_d = (p.t_h - p.t_0) / (double)height;
    p.p_d = (p.p_h - p.p_0) / (double)height;
    pma_even_rebalance(p, 1, capacity, nb_elements);
}
// pma.jl:69-84 with sort=false + _array pma.jl:33-40
void pma_build_sorted(Pma& p, const int64_t* keys, const double* vals, int64_t n) {
    if (n == 0) { pma_empty(p); return; }
    int64_t capacity = capacity_for(n, 0.7);
    p.array = Elements((size_t)capacity);
    for (int64_t i = 1; i <= n; ++i) p.array.set(i, keys[i - 1], vals[i - 1]);
    pma_init(p, n);
}
// pma.jl:86-91
void pma_empty(Pma& p, int64_t expected_nb_elems) {
    int64_t capacity = capacity_for(expected_nb_elems, 0.7);
    p.array = Elements((size_t)capacity);
    pma_init(p, 0);
}
// pma.jl:143-151
static void pma_extend(Pma& p) {
    p.capacity *= 2;
    p.nb_segments *= 2;
    p.height += 1;
    p.t_d = (p.t_h - p.t_0) / (double)p.height;
    p.p_d = (p.p_h - p.p_0) / (double)p.height;
    p.array.resize(p.capacity);
}
// pma.jl:153-161
static void pma_shrink(Pma& p) {
    p.capacity /= 2;
    p.nb_segments /= 2;
    p.height -= 1;
    p.t_d = (p.t_h - p.t_0) / (double)p.height;
    p.p_d = (p.p_h - p.p_0) / (double)p.height;
    p.array.resize(p.capacity);
}
// pma.jl:105-141
Window look_for_rebalance(Pma& pma, int64_t pos) {
    double p = 0.0, t = 0.0, density = 0.0;
    int64_t height = 0;
    int64_t prev_win_start = pos;
    int64_t prev_win_end = pos - 1;
    int64_t nb_cells_left = 0, nb_cells_right = 0;
    while (height <= pma.height) {
        int64_t window_capacity = (int64_t(1) << height) * pma.segment_capacity;
        int64_t win_start = ((pos - 1) / window_capacity) * window_capacity + 1;
        int64_t win_end = win_start + window_capacity - 1;
        nb_cells_left += nbcells(pma.array, win_start, prev_win_start);
        nb_cells_right += nbcells(pma.array, prev_win_end + 1, win_end + 1);
        density = (double)(nb_cells_left + nb_cells_right) / (double)window_capacity;
        p = pma.p_0 + pma.p_d * (double)height;
        t = pma.t_0 + pma.t_d * (double)height;
        if (p <= density && density <= t) {
            return Window{win_start, win_end, nb_cells_left + nb_cells_right};
        }
        prev_win_start = win_start;
        prev_win_end = win_end;
        height += 1;
    }
    int64_t nb_cells = nb_cells_left + nb_cells_right;
    if (density > t) pma_extend(pma);
    if (density < p && pma.height > 1) {
        pack(pma.array, 1, pma.array.length() / 2, nb_cells);
        pma_shrink(pma);
    }
    return Window{1, pma.capacity, nb_cells};
}
// pma.jl:189-193
double pma_get(const Pma& p, int64_t key) {
    int64_t pos = find(p.array, key, 1, p.array.length());
    if (pos != 0 && p.array.at(pos).key == key) return p.array.at(pos).val;
    return 0.0;
}
// pma.jl:196-213
void pma_set(Pma& p, double value, int64_t key) {
    if (value != 0.0) {
        PosFlag r = insert(p.array, key, value, 1, p.array.length(), nullptr);
        if (r.flag) {
            p.nb_elements += 1;
            Window w = look_for_rebalance(p, r.pos);
            pma_even_rebalance(p, w.start, w.end, w.nbcells);
        }
    } else {
        PosFlag r = del(p.array, key, 1, p.array.length());
        if (r.flag) {
            p.nb_elements -= 1;
            Window w = look_for_rebalance(p, r.pos);
            pma_even_rebalance(p, w.start, w.end, w.nbcells);
        }
    }
}

// ============================ vector.jl ========================================
// vector.jl:10-36   (sortperm = stable; left fold in input order)
void prepare_keys_vals(std::vector<int64_t>& keys, std::vector<double>& vals, int combine) {
    if (keys.size() != vals.size()) throw Error{ERR_ASSERT, "length(keys) == length(values)"};
    const size_t n = keys.size();
    if (n == 0) return;
    std::vector<size_t> p(n);
    std::iota(p.begin(), p.end(), size_t(0));
    std::stable_sort(p.begin(), p.end(), [&](size_t a, size_t b) { return keys[a] < keys[b]; });
    std::vector<int64_t> k2(n);
    std::vector<double> v2(n);
    for (size_t i = 0; i < n; ++i) { k2[i] = keys[p[i]]; v2[i] = vals[p[i]]; }
    keys.swap(k2);
    vals.swap(v2);
    size_t write_pos = 1, read_pos = 1;
    int64_t prev_id = keys[read_pos - 1];
    while (read_pos < n) {
        read_pos += 1;
        int64_t cur_id = keys[read_pos - 1];
        if (prev_id == cur_id) {
            vals[write_pos - 1] = combine_apply(combine, vals[write_pos - 1], vals[read_pos - 1]);
        } else {
            write_pos += 1;
            if (write_pos < read_pos) {
                keys[write_pos - 1] = cur_id;
                vals[write_pos - 1] = vals[read_pos - 1];
            }
        }
        prev_id = cur_id;
    }
    keys.resize(write_pos);
    vals.resize(write_pos);
}
// vector.jl:38-62  (_guess_length vector.jl:6 = maximum(keys; init = 0))
void vec_build(Vec& v, std::vector<int64_t> I, std::vector<double> V, int combine, int64_t n, bool n_given) {
    if (I.size() != V.size()) throw Error{ERR_ARGUMENT, "keys & nonzeros vectors must have same length."};
    if (!n_given) {
        n = 0;
        for (int64_t k : I) n = std::max(n, k);
    }
    prepare_keys_vals(I, V, combine);
    // PackedMemoryArray(keys, values) with sort=true (pma.jl:69-84): already sorted, sortperm is the identity
    pma_build_sorted(v.pma, I.data(), V.data(), (int64_t)I.size());
    v.n = n;
}
// vector.jl:76-81
void vec_set(Vec& v, double value, int64_t key) {
    if (value != 0.0) v.n = std::max(v.n, key);
    pma_set(v.pma, value, key);
}
double vec_get(const Vec& v, int64_t key) { return pma_get(v.pma, key); }

// ============================ pcsr.jl ==========================================
// pcsr.jl:88-97
static void pcsc_even_rebalance(Pcsc& m, int64_t ws, int64_t we, int64_t nbcells_) {
    int64_t capacity = we - ws + 1;
    if (capacity == m.pma.segment_capacity) return;
    pack(m.pma.array, ws, we, nbcells_);
    spread5(m.pma.array, ws, we, nbcells_, &m.semaphores);
}
// pcsr.jl:65-68
void pcsc_empty(Pcsc& m) {
    pma_empty(m.pma);
    m.nb_partitions = 0;
    m.semaphores.clear();
}
// pcsr.jl:26-63
void pcsc_build(Pcsc& m, const std::vector<std::vector<int64_t>>& row_keys,
                const std::vector<std::vector<double>>& values, int combine) {
    const int64_t nb_semaphores = (int64_t)row_keys.size();
    if (nb_semaphores != (int64_t)values.size()) throw Error{ERR_ASSERT, "nb_semaphores == length(values)"};
    std::vector<int64_t> pk;
    std::vector<double> pv;
    for (int64_t sid = 1; sid <= nb_semaphores; ++sid) {
        pk.push_back(0);               // semaphore_key(L) = zero(L)   pcsr.jl:23,39
        pv.push_back((double)sid);     // T(semaphore_id)              pcsr.jl:40
        std::vector<int64_t> nk = row_keys[sid - 1];
        std::vector<double> nv = values[sid - 1];
        prepare_keys_vals(nk, nv, combine);
        for (size_t j = 0; j < nk.size(); ++j) { pk.push_back(nk[j]); pv.push_back(nv[j]); }
    }
    pma_build_sorted(m.pma, pk.data(), pv.data(), (int64_t)pk.size());   // sort = false
    m.semaphores.assign((size_t)nb_semaphores, 0);
    for (int64_t pos = 1; pos <= m.pma.array.length(); ++pos) {          // pcsr.jl:55-61
        if (!m.pma.array.empty_at(pos) && m.pma.array.at(pos).key == 0) {
            int64_t id = (int64_t)m.pma.array.at(pos).val;
            m.semaphores[id - 1] = pos;
        }
    }
    m.nb_partitions = nb_semaphores;
}
// pcsr.jl:99-112
void pcsc_addpartition_end(Pcsc& m) {
    int64_t sem_pos = m.pma.array.length();
    m.nb_partitions += 1;
    m.semaphores.push_back(sem_pos);
    double sem_val = (double)m.semaphores.size();
    PosFlag r = insert_after(m.pma.array, 0, sem_val, sem_pos, &m.semaphores);
    if (r.flag) {
        m.pma.nb_elements += 1;
        Window w = look_for_rebalance(m.pma, r.pos);
        pcsc_even_rebalance(m, w.start, w.end, w.nbcells);
    }
}
// pcsr.jl:114-146
void pcsc_addpartition_after(Pcsc& m, int64_t prev_sem_id) {
    Semaphores& semaphores = m.semaphores;
    int64_t nb_semaphores = (int64_t)semaphores.size();
    int64_t sem_pos = 0;
    if (!(prev_sem_id + 1 >= 1 && prev_sem_id + 1 <= nb_semaphores)) throw Error{ERR_BOUNDS, "semaphores[prev_sem_id + 1]"};
    int64_t semaphore_target = semaphores[prev_sem_id];   // semaphores[prev_sem_id + 1]
    if (semaphore_target == 0) {
        int64_t next_sem_id = nextnonemptypos_sem(semaphores, prev_sem_id + 1);
        if (next_sem_id == 0) throw Error{ERR_BOUNDS, "semaphores[0] (reference bug: reuse of a trailing deleted slot, pcsr.jl:121-125)"};
        int64_t next_semaphore = semaphores[next_sem_id - 1];
        sem_pos = next_semaphore - 1;
    } else {
        sem_pos = semaphore_target - 1;
        semaphores.resize((size_t)nb_semaphores + 1, 0);
        for (int64_t i = nb_semaphores; i >= prev_sem_id + 1; --i) {
            int64_t moved_sem_pos = semaphores[i - 1];
            semaphores[i] = semaphores[i - 1];
            if (moved_sem_pos == 0) throw Error{ERR_ASSERT, "!isnothing(moved_sem_pos) (reference bug: mid-insert with a deleted partition to the right, pcsr.jl:129-133)"};
            m.pma.array.set(moved_sem_pos, 0, (double)(i + 1));
        }
    }
    m.nb_partitions += 1;
    double sem_val = (double)(prev_sem_id + 1);
    PosFlag r = insert_after(m.pma.array, 0, sem_val, sem_pos, &m.semaphores);
    semaphores[prev_sem_id] = r.pos;
    if (r.flag) {
        m.pma.nb_elements += 1;
        Window w = look_for_rebalance(m.pma, r.pos);
        pcsc_even_rebalance(m, w.start, w.end, w.nbcells);
    }
}
// pcsr.jl:171-175
int64_t pos_of_partition_start(const Pcsc& m, int64_t partition) {
    int64_t s = m.semaphores[partition - 1];
    if (s == 0) throw Error{ERR_ASSERT, "!isnothing(partition_start_pos)"};
    return s;
}
// pcsr.jl:177-186
int64_t pos_of_partition_end(const Pcsc& m, int64_t partition) {
    int64_t pos = m.pma.array.length();
    int64_t next_partition = nextnonemptypos_sem(m.semaphores, partition);
    if (next_partition != 0) pos = m.semaphores[next_partition - 1] - 1;
    return pos;
}
// pcsr.jl:188-204
void pcsc_deletepartition(Pcsc& m, int64_t partition) {
    int64_t len = (int64_t)m.semaphores.size();
    if (!(1 <= partition && partition <= len)) throw Error{ERR_BOUNDS, "cannot access partition"};
    m.nb_partitions -= 1;
    int64_t sem_pos = pos_of_partition_start(m, partition);
    int64_t partition_end_pos = pos_of_partition_end(m, partition);
    PurgeRes pr = purge(m.pma.array, sem_pos, partition_end_pos);
    if (pr.nb > 0) {
        m.pma.nb_elements -= pr.nb;
        Window w = look_for_rebalance(m.pma, pr.mid);
        pcsc_even_rebalance(m, w.start, w.end, w.nbcells);
    }
    m.semaphores[partition - 1] = 0;
}
// pcsr.jl:222-232
double pcsc_get(const Pcsc& m, int64_t key, int64_t partition) {
    int64_t from = pos_of_partition_start(m, partition);
    int64_t to = pos_of_partition_end(m, partition);
    int64_t pos = find(m.pma.array, key, from, to);
    if (pos != 0 && m.pma.array.at(pos).key == key) return m.pma.array.at(pos).val;
    return 0.0;
}
// pcsr.jl:321-339
static void pcsc_insert(Pcsc& m, double value, int64_t key, int64_t from, int64_t to) {
    PosFlag r = insert(m.pma.array, key, value, from, to, &m.semaphores);
    if (r.flag) {
        m.pma.nb_elements += 1;
        Window w = look_for_rebalance(m.pma, r.pos);
        pcsc_even_rebalance(m, w.start, w.end, w.nbcells);
    }
}
static void pcsc_delete(Pcsc& m, int64_t key, int64_t from, int64_t to) {
    PosFlag r = del(m.pma.array, key, from, to);
    if (r.flag) {
        m.pma.nb_elements -= 1;
        Window w = look_for_rebalance(m.pma, r.pos);
        pcsc_even_rebalance(m, w.start, w.end, w.nbcells);
    }
}
// pcsr.jl:294-319
void pcsc_set(Pcsc& m, double value, int64_t key, int64_t partition) {
    if (partition > (int64_t)m.semaphores.size()) {
        int64_t p = (int64_t)m.semaphores.size() + 1;   // _add_partitions!
        while (p <= partition) { pcsc_addpartition_end(m); p += 1; }
    }
    if (partition < 1) throw Error{ERR_BOUNDS, "semaphores[partition]"};
    int64_t from = m.semaphores[partition - 1];
    if (from == 0) throw Error{ERR_ERROR, "The partition has been deleted."};
    int64_t to = pos_of_partition_end(m, partition);
    if (value != 0.0) pcsc_insert(m, value, key, from + 1, to);
    else pcsc_delete(m, key, from, to);
}
// pcsr.jl:148-169
int64_t mpcsc_addcolumn(Mpcsc& m, int64_t col, int64_t prev_col_pos) {
    int64_t col_pos = 0;
    ColKeys& ck = m.col_keys;
    if (prev_col_pos == ck.length()) {
        ck.key.push_back(col); ck.live.push_back(1);
        pcsc_addpartition_end(m.pcsc);
        col_pos = ck.length();
    } else {
        if (!ck.live[prev_col_pos]) {   // col_keys[prev_col_pos+1] === nothing
            ck.key[prev_col_pos] = col; ck.live[prev_col_pos] = 1;
        } else {
            int64_t nbcolkeys = ck.length();
            ck.key.resize((size_t)nbcolkeys + 1); ck.live.resize((size_t)nbcolkeys + 1);
            for (int64_t i = nbcolkeys; i >= prev_col_pos + 1; --i) {
                ck.key[i] = ck.key[i - 1]; ck.live[i] = ck.live[i - 1];
            }
            ck.key[prev_col_pos] = col; ck.live[prev_col_pos] = 1;
        }
        pcsc_addpartition_after(m.pcsc, prev_col_pos);
        col_pos = prev_col_pos + 1;
    }
    return col_pos;
}
// pcsr.jl:206-212
void mpcsc_deletecolumn(Mpcsc& m, int64_t col) {
    bool exact;
    int64_t col_pos = colkeys_find(m.col_keys, col, &exact);
    if (!exact) throw Error{ERR_ARGUMENT, "column does not exist."};
    m.col_keys.live[col_pos - 1] = 0;
    pcsc_deletepartition(m.pcsc, col_pos);
}
// pcsr.jl:261-267
double mpcsc_get(const Mpcsc& m, int64_t row, int64_t col) {
    bool exact;
    int64_t col_pos = colkeys_find(m.col_keys, col, &exact);
    if (!exact) return 0.0;
    return pcsc_get(m.pcsc, row, col_pos);
}
// pcsr.jl:341-347
void mpcsc_set(Mpcsc& m, double value, int64_t row, int64_t col) {
    bool exact;
    int64_t col_pos = colkeys_find(m.col_keys, col, &exact);
    if (!exact) col_pos = mpcsc_addcolumn(m, col, col_pos);
    pcsc_set(m.pcsc, value, row, col_pos);
}
// pcsr.jl:354-431 + 433-449
void mpcsc_build_coo(Mpcsc& m, std::vector<int64_t> I, std::vector<int64_t> J, std::vector<double> V, int combine) {
    if (!(I.size() == J.size() && J.size() == V.size()))
        throw Error{ERR_ARGUMENT, "rows, columns, and nonzeros do not have same length."};
    const size_t n = I.size();
    if (n == 0) {   // pcsr.jl:443 -> MappedPackedCSC(K,L,T)
        pcsc_empty(m.pcsc);
        m.col_keys = ColKeys();
        return;
    }
    std::vector<size_t> p(n);
    std::iota(p.begin(), p.end(), size_t(0));
    std::stable_sort(p.begin(), p.end(), [&](size_t a, size_t b) {   // sortperm(zip(J,I)): columns first, ties by index
        if (J[a] != J[b]) return J[a] < J[b];
        return I[a] < I[b];
    });
    {
        std::vector<int64_t> I2(n), J2(n);
        std::vector<double> V2(n);
        for (size_t i = 0; i < n; ++i) { I2[i] = I[p[i]]; J2[i] = J[p[i]]; V2[i] = V[p[i]]; }
        I.swap(I2); J.swap(J2); V.swap(V2);
    }
    int64_t nb_cols = 1;
    std::vector<int64_t> nb_rows_in_col;
    nb_rows_in_col.push_back(1);
    size_t write_pos = 1, read_pos = 1;
    int64_t prev_i = I[0], prev_j = J[0];
    while (read_pos < n) {
        read_pos += 1;
        int64_t cur_i = I[read_pos - 1], cur_j = J[read_pos - 1];
        if (prev_i == cur_i && prev_j == cur_j) {
            V[write_pos - 1] = combine_apply(combine, V[write_pos - 1], V[read_pos - 1]);
        } else {
            write_pos += 1;
            if (write_pos < read_pos) {
                I[write_pos - 1] = cur_i; J[write_pos - 1] = cur_j; V[write_pos - 1] = V[read_pos - 1];
            }
            if (cur_j != prev_j) { nb_cols += 1; nb_rows_in_col.push_back(1); }
            else if (cur_i != prev_i) { nb_rows_in_col.back() += 1; }
            prev_i = cur_i; prev_j = cur_j;
        }
    }
    I.resize(write_pos); J.resize(write_pos); V.resize(write_pos);

    std::vector<int64_t> col_keys((size_t)nb_cols);
    std::vector<std::vector<int64_t>> row_keys((size_t)nb_cols);
    std::vector<std::vector<double>> values((size_t)nb_cols);
    size_t i = 1;
    int64_t prev_col = J[0];
    int64_t col_pos = 0, row_pos = 0;
    while (i <= I.size()) {
        int64_t cur_col = J[i - 1];
        if (prev_col != cur_col || i == 1) {
            col_pos += 1;
            row_pos = 1;
            col_keys[col_pos - 1] = cur_col;
            row_keys[col_pos - 1].assign((size_t)nb_rows_in_col[col_pos - 1], 0);
            values[col_pos - 1].assign((size_t)nb_rows_in_col[col_pos - 1], 0.0);
        }
        row_keys[col_pos - 1][row_pos - 1] = I[i - 1];
        values[col_pos - 1][row_pos - 1] = V[i - 1];
        prev_col = cur_col;
        row_pos += 1;
        i += 1;
    }
    // MappedPackedCSC(row_keys, col_keys, values, combine)   pcsr.jl:73-80
    pcsc_build(m.pcsc, row_keys, values, combine);
    m.col_keys.key = col_keys;
    m.col_keys.live.assign(col_keys.size(), 1);
}
// pcsr.jl:285-291 -> 248-258 ; views.jl:15-35 (column gather = compaction of the span)
void mpcsc_column(const Mpcsc& m, int64_t col, std::vector<int64_t>& keys, std::vector<double>& vals) {
    keys.clear(); vals.clear();
    bool exact;
    int64_t col_pos = colkeys_find(m.col_keys, col, &exact);
    if (!exact) return;
    int64_t from = pos_of_partition_start(m.pcsc, col_pos) + 1;
    int64_t to = pos_of_partition_end(m.pcsc, col_pos);
    for (int64_t pos = from; pos <= to; ++pos) {
        if (!m.pcsc.pma.array.empty_at(pos)) {
            keys.push_back(m.pcsc.pma.array.at(pos).key);
            vals.push_back(m.pcsc.pma.array.at(pos).val);
        }
    }
}
// pcsr.jl:269-283 (full scan tracking the current partition); result keyed by col_keys[partition], sorted by key
void mpcsc_row(const Mpcsc& m, int64_t row, std::vector<int64_t>& keys, std::vector<double>& vals) {
    keys.clear(); vals.clear();
    int64_t partition_id = 0;
    std::vector<std::pair<int64_t, double>> el;
    const Elements& a = m.pcsc.pma.array;
    for (int64_t pos = 1; pos <= a.length(); ++pos) {
        if (a.empty_at(pos)) continue;
        int64_t k = a.at(pos).key;
        double v = a.at(pos).val;
        if (k == 0) partition_id = (int64_t)v;
        if (k == row) el.push_back({m.col_keys.key[partition_id - 1], v});
    }
    std::stable_sort(el.begin(), el.end(), [](auto& x, auto& y) { return x.first < y.first; });  // PackedMemoryArray(elements) sorts
    for (auto& e : el) { keys.push_back(e.first); vals.push_back(e.second); }
}

// ============================ operations.jl ====================================
// operations.jl:62-105 + 120-135.  x given as ascending (key, value) pairs (iteration order of x.pma / rowvals).
void mpcsc_mul(const Mpcsc& mat, const int64_t* xk, const double* xv, int64_t nx,
               std::vector<int64_t>& yk, std::vector<double>& yv) {
    std::unordered_map<int64_t, double> result;
    const ColKeys& ck = mat.col_keys;
    const Elements& arr = mat.pcsc.pma.array;
    int64_t col_key_pos = 1;
    for (int64_t t = 0; t < nx; ++t) {
        int64_t vec_row_id = xk[t];
        double vec_val = xv[t];
        // _mul_dyn_mat_col_loop!
        while (col_key_pos <= ck.length()) {
            if (ck.live[col_key_pos - 1] && ck.key[col_key_pos - 1] >= vec_row_id) break;
            col_key_pos += 1;
        }
        if (col_key_pos > ck.length()) break;   // stop
        if (!ck.live[col_key_pos - 1] || ck.key[col_key_pos - 1] != vec_row_id) continue;
        int64_t next_col_key_pos = col_key_pos + 1;
        while (next_col_key_pos <= ck.length() && !ck.live[next_col_key_pos - 1]) next_col_key_pos += 1;
        int64_t cur_semaphore = mat.pcsc.semaphores[col_key_pos - 1];
        if (cur_semaphore == 0) throw Error{ERR_ASSERT, "!isnothing(cur_semaphore)"};
        int64_t mat_row_start = cur_semaphore + 1;
        int64_t mat_row_end = arr.length();
        if (next_col_key_pos <= ck.length()) {
            int64_t next_semaphore = mat.pcsc.semaphores[next_col_key_pos - 1];
            if (next_semaphore == 0) throw Error{ERR_ASSERT, "!isnothing(next_semaphore)"};
            mat_row_end = next_semaphore - 1;
        }
        for (int64_t pos = mat_row_start; pos <= mat_row_end; ++pos) {
            if (!arr.empty_at(pos)) {
                int64_t mat_row_id = arr.at(pos).key;
                double coeff = arr.at(pos).val;
                auto it = result.find(mat_row_id);
                double cur = (it == result.end()) ? 0.0 : it->second;
                double prod = vec_val * coeff;           // separate mul and add (no FMA)
                result[mat_row_id] = cur + prod;
            }
        }
        col_key_pos = next_col_key_pos;
    }
    // _mul_output -> sparsevec(result, n): indices sorted ascending, stored zeros kept
    std::vector<std::pair<int64_t, double>> out(result.begin(), result.end());
    std::sort(out.begin(), out.end(), [](auto& a, auto& b) { return a.first < b.first; });
    yk.clear(); yv.clear();
    for (auto& e : out) { yk.push_back(e.first); yv.push_back(e.second); }
}

// ============================ buffer.jl / matrix.jl ============================
// matrix.jl:15-19
void matrix_build(Matrix& A, const std::vector<int64_t>& I, const std::vector<int64_t>& J,
                  const std::vector<double>& V, int64_t m, int64_t n, bool dims_given, int combine) {
    if (!dims_given) {
        m = 0; n = 0;
        for (int64_t k : I) m = std::max(m, k);
        for (int64_t k : J) n = std::max(n, k);
    }
    A.m = m; A.n = n; A.fillmode = false;
    A.buffer = Buffer();
    mpcsc_build_coo(A.colmajor, I, J, V, combine);
    mpcsc_build_coo(A.rowmajor, J, I, V, combine);
}
// matrix.jl:31-41
void matrix_empty(Matrix& A, bool fill_mode) {
    A.m = 0; A.n = 0; A.fillmode = fill_mode;
    A.buffer = Buffer();
    if (!fill_mode) {
        mpcsc_build_coo(A.colmajor, {}, {}, {}, COMB_ADD);
        mpcsc_build_coo(A.rowmajor, {}, {}, {}, COMB_ADD);
    }
}
// buffer.jl:20-31
static void buffer_addelem(Buffer& b, int64_t rowid, int64_t colid, double val) {
    auto it = b.index.find(rowid);
    size_t idx;
    if (it == b.index.end()) {
        idx = b.rowid.size();
        b.index[rowid] = idx;
        b.rowid.push_back(rowid);
        b.colids.emplace_back();
        b.vals.emplace_back();
    } else idx = it->second;
    b.colids[idx].push_back(colid);
    b.vals[idx].push_back(val);
    b.length += 1;
}
// buffer.jl:10-18
static void buffer_addrow(Buffer& b, int64_t rowid, const std::vector<int64_t>& colids, const std::vector<double>& vals) {
    if (b.index.count(rowid)) throw Error{ERR_ERROR, "Row already written in dynamic sparse matrix buffer."};
    std::vector<size_t> p(colids.size());
    std::iota(p.begin(), p.end(), size_t(0));
    std::stable_sort(p.begin(), p.end(), [&](size_t a, size_t c) { return colids[a] < colids[c]; });
    size_t idx = b.rowid.size();
    b.index[rowid] = idx;
    b.rowid.push_back(rowid);
    b.colids.emplace_back(); b.vals.emplace_back();
    for (size_t k : p) { b.colids[idx].push_back(colids[k]); b.vals[idx].push_back(vals[k]); }
    b.length += (int64_t)vals.size();
}
// matrix.jl:43-62
void matrix_set(Matrix& A, double val, int64_t row, int64_t col) {
    if (val != 0.0) {
        A.m = std::max(A.m, row);
        A.n = std::max(A.n, col);
    }
    if (A.fillmode) {
        buffer_addelem(A.buffer, row, col, val);
    } else {
        mpcsc_set(A.colmajor, val, row, col);
        mpcsc_set(A.rowmajor, val, col, row);
    }
}
// matrix.jl:64-68 (non fill mode only)
double matrix_get(const Matrix& A, int64_t row, int64_t col) {
    if (A.fillmode) throw Error{ERR_ERROR, "getindex(row, col) not available in fill mode"};
    return mpcsc_get(A.colmajor, row, col);
}
// matrix.jl:95-102
void matrix_deletecolumn(Matrix& A, int64_t col) {
    if (A.fillmode) throw Error{ERR_ERROR, "Cannot delete a column in fill mode"};
    std::vector<int64_t> rows; std::vector<double> vals;
    mpcsc_column(A.colmajor, col, rows, vals);      // @view matrix[:, col]
    for (int64_t r : rows) mpcsc_set(A.rowmajor, 0.0, col, r);
    mpcsc_deletecolumn(A.colmajor, col);
}
// matrix.jl:104-111
void matrix_deleterow(Matrix& A, int64_t row) {
    if (A.fillmode) throw Error{ERR_ERROR, "Cannot delete a row in fill mode"};
    std::vector<int64_t> cols; std::vector<double> vals;
    mpcsc_column(A.rowmajor, row, cols, vals);      // @view matrix[row, :]
    for (int64_t c : cols) mpcsc_set(A.colmajor, 0.0, row, c);
    mpcsc_deletecolumn(A.rowmajor, row);
}
// matrix.jl:113-124
void matrix_addrow(Matrix& A, int64_t row, const std::vector<int64_t>& colids, const std::vector<double>& vals) {
    if (A.fillmode) {
        buffer_addrow(A.buffer, row, colids, vals);
    } else {
        for (size_t j = 0; j < colids.size(); ++j) matrix_set(A, vals[j], row, colids[j]);
    }
}
// matrix.jl:126-134 + buffer.jl:33-50
void matrix_closefillmode(Matrix& A) {
    if (!A.fillmode) throw Error{ERR_ERROR, "Cannot close fill mode because matrix is not in fill mode."};
    std::vector<int64_t> I, J; std::vector<double> V;
    for (size_t r = 0; r < A.buffer.rowid.size(); ++r) {
        for (size_t k = 0; k < A.buffer.vals[r].size(); ++k) {
            I.push_back(A.buffer.rowid[r]);
            J.push_back(A.buffer.colids[r][k]);
            V.push_back(A.buffer.vals[r][k]);
        }
    }
    A.fillmode = false;
    A.buffer = Buffer();
    mpcsc_build_coo(A.colmajor, I, J, V, COMB_ADD);
    mpcsc_build_coo(A.rowmajor, J, I, V, COMB_ADD);
}

}  // namespace orc
