// TEST INFRASTRUCTURE — NOT PRODUCT CODE.  extern "C" surface of the CPU oracle
// (ctypes-loaded by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline only).
#include "dsa_oracle.hpp"
#include "batch_policy.hpp"
#include <cstring>
#include <chrono>

using namespace orc;

static thread_local std::string g_err;

#define ORC_TRY try {
#define ORC_CATCH                                             \
    }                                                         \
    catch (const Error& e) { g_err = e.msg; return e.code; }  \
    catch (const std::exception& e) { g_err = e.what(); return 99; }

static Elements make_elements(const uint8_t* tag, const int64_t* key, const double* val, int64_t n) {
    Elements a((size_t)n);
    for (int64_t i = 0; i < n; ++i) {
        a.tag[i] = tag[i];
        a.kv[i] = KV{key[i], val[i]};
    }
    return a;
}
static void store_elements(const Elements& a, uint8_t* tag, int64_t* key, double* val) {
    for (int64_t i = 0; i < a.length(); ++i) {
        tag[i] = a.tag[i];
        key[i] = a.tag[i] ? a.kv[i].key : 0;
        val[i] = a.tag[i] ? a.kv[i].val : 0.0;
    }
}

extern "C" {

const char* orc_last_error() { return g_err.c_str(); }

// ---- raw gapped-array primitives (finds.jl / writes.jl / moves.jl) -------------------
int orc_find(const uint8_t* tag, const int64_t* key, const double* val, int64_t n, int64_t k,
             int64_t from, int64_t to, int64_t* pos_out) {
    ORC_TRY
    Elements a = make_elements(tag, key, val, n);
    *pos_out = find(a, k, from, to);
    return 0;
    ORC_CATCH
}
int orc_insert(uint8_t* tag, int64_t* key, double* val, int64_t n, int64_t k, double v, int64_t from, int64_t to,
               int64_t* sem, int64_t nsem, int64_t* pos_out, int* isnew_out) {
    ORC_TRY
    Elements a = make_elements(tag, key, val, n);
    Semaphores s;
    if (sem) s.assign(sem, sem + nsem);
    PosFlag r = insert(a, k, v, from, to, sem ? &s : nullptr);
    store_elements(a, tag, key, val);
    if (sem) std::copy(s.begin(), s.end(), sem);
    *pos_out = r.pos; *isnew_out = r.flag;
    return 0;
    ORC_CATCH
}
int orc_delete(uint8_t* tag, int64_t* key, double* val, int64_t n, int64_t k, int64_t from, int64_t to,
               int64_t* pos_out, int* deleted_out) {
    ORC_TRY
    Elements a = make_elements(tag, key, val, n);
    PosFlag r = del(a, k, from, to);
    store_elements(a, tag, key, val);
    *pos_out = r.pos; *deleted_out = r.flag;
    return 0;
    ORC_CATCH
}
int orc_purge(uint8_t* tag, int64_t* key, double* val, int64_t n, int64_t from, int64_t to, int64_t* mid_out, int64_t* nb_out) {
    ORC_TRY
    Elements a = make_elements(tag, key, val, n);
    PurgeRes r = purge(a, from, to);
    store_elements(a, tag, key, val);
    *mid_out = r.mid; *nb_out = r.nb;
    return 0;
    ORC_CATCH
}
int orc_move(uint8_t* tag, int64_t* key, double* val, int64_t n, int right, int64_t from, int64_t to, int64_t* sem, int64_t nsem) {
    ORC_TRY
    Elements a = make_elements(tag, key, val, n);
    Semaphores s;
    if (sem) s.assign(sem, sem + nsem);
    if (right) movecellstoright(a, from, to, sem ? &s : nullptr);
    else movecellstoleft(a, from, to, sem ? &s : nullptr);
    store_elements(a, tag, key, val);
    if (sem) std::copy(s.begin(), s.end(), sem);
    return 0;
    ORC_CATCH
}
int orc_pack(uint8_t* tag, int64_t* key, double* val, int64_t n, int64_t ws, int64_t we, int64_t m) {
    ORC_TRY
    Elements a = make_elements(tag, key, val, n);
    pack(a, ws, we, m);
    store_elements(a, tag, key, val);
    return 0;
    ORC_CATCH
}
int orc_spread(uint8_t* tag, int64_t* key, double* val, int64_t n, int64_t ws, int64_t we, int64_t m,
               int five_arg, int64_t* sem, int64_t nsem) {
    ORC_TRY
    Elements a = make_elements(tag, key, val, n);
    Semaphores s;
    if (sem) s.assign(sem, sem + nsem);
    if (five_arg) spread5(a, ws, we, m, sem ? &s : nullptr);
    else spread4(a, ws, we, m);
    store_elements(a, tag, key, val);
    if (sem) std::copy(s.begin(), s.end(), sem);
    return 0;
    ORC_CATCH
}
int orc_arrays_equal(const uint8_t* t1, const int64_t* k1, const double* v1, int64_t n1,
                     const uint8_t* t2, const int64_t* k2, const double* v2, int64_t n2) {
    Elements a = make_elements(t1, k1, v1, n1), b = make_elements(t2, k2, v2, n2);
    return arrays_equal(a, b) ? 1 : 0;
}
// geometry of a bulk-built / empty PMA: out = {capacity, segment_capacity, nb_segments, height}
int orc_geometry(int64_t n, int64_t* out) {
    ORC_TRY
    Pma p;
    if (n == 0) pma_empty(p);
    else {
        std::vector<int64_t> k((size_t)n);
        std::vector<double> v((size_t)n, 1.0);
        for (int64_t i = 0; i < n; ++i) k[i] = i + 1;
        pma_build_sorted(p, k.data(), v.data(), n);
    }
    out[0] = p.capacity; out[1] = p.segment_capacity; out[2] = p.nb_segments; out[3] = p.height;
    return 0;
    ORC_CATCH
}

// ---- PMA export helper -------------------------------------------------------------
static void export_pma(const Pma& p, uint8_t* tag, int64_t* key, double* val) { store_elements(p.array, tag, key, val); }
static void pma_info(const Pma& p, int64_t* out) {
    out[0] = p.capacity; out[1] = p.segment_capacity; out[2] = p.nb_segments; out[3] = p.nb_elements; out[4] = p.height;
}

// ---- DynamicSparseVector -----------------------------------------------------------
void* orc_vec_build(const int64_t* I, const double* V, int64_t n, int combine, int64_t len, int len_given, int* err) {
    *err = 0;
    Vec* v = new Vec();
    try {
        vec_build(*v, std::vector<int64_t>(I, I + n), std::vector<double>(V, V + n), combine, len, len_given != 0);
    } catch (const Error& e) { g_err = e.msg; *err = e.code; delete v; return nullptr; }
    return v;
}
void orc_vec_free(void* h) { delete (Vec*)h; }
void* orc_vec_clone(void* h) { return new Vec(*(Vec*)h); }
int orc_vec_set(void* h, int64_t key, double val) {
    ORC_TRY
    vec_set(*(Vec*)h, val, key);
    return 0;
    ORC_CATCH
}
int orc_vec_set_many(void* h, const int64_t* keys, const double* vals, int64_t n) {   // loop of setindex! (matrix.jl:119-121 style)
    ORC_TRY
    Vec& v = *(Vec*)h;
    for (int64_t i = 0; i < n; ++i) vec_set(v, vals[i], keys[i]);
    return 0;
    ORC_CATCH
}
double orc_vec_get(void* h, int64_t key) { return vec_get(*(Vec*)h, key); }
void orc_vec_get_many(void* h, const int64_t* keys, int64_t n, double* out) {
    Vec& v = *(Vec*)h;
    for (int64_t i = 0; i < n; ++i) out[i] = vec_get(v, keys[i]);
}
// out = {capacity, segment_capacity, nb_segments, nb_elements, height, n}
void orc_vec_info(void* h, int64_t* out) { pma_info(((Vec*)h)->pma, out); out[5] = ((Vec*)h)->n; }
void orc_vec_export(void* h, uint8_t* tag, int64_t* key, double* val) { export_pma(((Vec*)h)->pma, tag, key, val); }
int64_t orc_vec_shrink_size(void* h) {   // vector.jl:64
    Vec& v = *(Vec*)h;
    int64_t mx = 0;
    for (int64_t pos = 1; pos <= v.pma.array.length(); ++pos)
        if (!v.pma.array.empty_at(pos)) mx = std::max(mx, v.pma.array.at(pos).key);
    v.n = mx;
    return mx;
}
int orc_vec_equal(void* h1, void* h2) {   // vector.jl:85 + pma.jl:262
    Vec& a = *(Vec*)h1; Vec& b = *(Vec*)h2;
    if (a.n != b.n) return 0;
    if (a.pma.nb_elements != b.pma.nb_elements) return 0;
    return arrays_equal(a.pma.array, b.pma.array) ? 1 : 0;
}
// batch policy (CPU statement of the GPU batch algorithm; layout target for the CUDA path)
int orc_vec_set_batch_policy(void* h, const int64_t* keys, const double* vals, int64_t n) {
    ORC_TRY
    Vec& v = *(Vec*)h;
    for (int64_t i = 0; i < n; ++i) if (vals[i] != 0.0) v.n = std::max(v.n, keys[i]);
    policy::pma_set_batch(v.pma, keys, vals, n);
    return 0;
    ORC_CATCH
}

// ---- raw PackedCSC (partitions addressed by integer id) -----------------------------
void* orc_pcsc_build(const int64_t* keys, const double* vals, const int64_t* offsets, int64_t nparts, int combine, int* err) {
    *err = 0;
    Pcsc* m = new Pcsc();
    try {
        std::vector<std::vector<int64_t>> rk((size_t)nparts);
        std::vector<std::vector<double>> rv((size_t)nparts);
        for (int64_t p = 0; p < nparts; ++p) {
            rk[p].assign(keys + offsets[p], keys + offsets[p + 1]);
            rv[p].assign(vals + offsets[p], vals + offsets[p + 1]);
        }
        pcsc_build(*m, rk, rv, combine);
    } catch (const Error& e) { g_err = e.msg; *err = e.code; delete m; return nullptr; }
    return m;
}
void orc_pcsc_free(void* h) { delete (Pcsc*)h; }
void* orc_pcsc_clone(void* h) { return new Pcsc(*(Pcsc*)h); }
int orc_pcsc_set(void* h, int64_t key, int64_t partition, double val) {
    ORC_TRY
    pcsc_set(*(Pcsc*)h, val, key, partition);
    return 0;
    ORC_CATCH
}
int orc_pcsc_get(void* h, int64_t key, int64_t partition, double* out) {
    ORC_TRY
    *out = pcsc_get(*(Pcsc*)h, key, partition);
    return 0;
    ORC_CATCH
}
int orc_pcsc_deletepartition(void* h, int64_t partition) {
    ORC_TRY
    pcsc_deletepartition(*(Pcsc*)h, partition);
    return 0;
    ORC_CATCH
}
// out = {capacity, seg, nsegs, nb_elements, height, nb_partitions, len(semaphores)}
void orc_pcsc_info(void* h, int64_t* out) {
    Pcsc& m = *(Pcsc*)h;
    pma_info(m.pma, out);
    out[5] = m.nb_partitions; out[6] = (int64_t)m.semaphores.size();
}
void orc_pcsc_export(void* h, uint8_t* tag, int64_t* key, double* val, int64_t* sem) {
    Pcsc& m = *(Pcsc*)h;
    export_pma(m.pma, tag, key, val);
    std::copy(m.semaphores.begin(), m.semaphores.end(), sem);
}

// ---- DynamicSparseMatrix -------------------------------------------------------------
void* orc_mat_build(const int64_t* I, const int64_t* J, const double* V, int64_t n, int64_t m_, int64_t n_, int dims_given,
                    int combine, int* err) {
    *err = 0;
    Matrix* A = new Matrix();
    try {
        matrix_build(*A, std::vector<int64_t>(I, I + n), std::vector<int64_t>(J, J + n), std::vector<double>(V, V + n),
                     m_, n_, dims_given != 0, combine);
    } catch (const Error& e) { g_err = e.msg; *err = e.code; delete A; return nullptr; }
    return A;
}
void* orc_mat_empty(int fill_mode) {
    Matrix* A = new Matrix();
    matrix_empty(*A, fill_mode != 0);
    return A;
}
void orc_mat_free(void* h) { delete (Matrix*)h; }
void* orc_mat_clone(void* h) { return new Matrix(*(Matrix*)h); }
int orc_mat_set(void* h, int64_t row, int64_t col, double val) {
    ORC_TRY
    matrix_set(*(Matrix*)h, val, row, col);
    return 0;
    ORC_CATCH
}
int orc_mat_set_many(void* h, const int64_t* rows, const int64_t* cols, const double* vals, int64_t n) {
    ORC_TRY
    Matrix& A = *(Matrix*)h;
    for (int64_t i = 0; i < n; ++i) matrix_set(A, vals[i], rows[i], cols[i]);
    return 0;
    ORC_CATCH
}
int orc_mat_get(void* h, int64_t row, int64_t col, double* out) {
    ORC_TRY
    *out = matrix_get(*(Matrix*)h, row, col);
    return 0;
    ORC_CATCH
}
int orc_mat_get_many(void* h, int which, const int64_t* rows, const int64_t* cols, int64_t n, double* out) {
    ORC_TRY
    Matrix& A = *(Matrix*)h;
    for (int64_t i = 0; i < n; ++i)
        out[i] = which == 0 ? mpcsc_get(A.colmajor, rows[i], cols[i]) : mpcsc_get(A.rowmajor, cols[i], rows[i]);
    return 0;
    ORC_CATCH
}
int orc_mat_deletecolumn(void* h, int64_t col) {
    ORC_TRY
    matrix_deletecolumn(*(Matrix*)h, col);
    return 0;
    ORC_CATCH
}
int orc_mat_deleterow(void* h, int64_t row) {
    ORC_TRY
    matrix_deleterow(*(Matrix*)h, row);
    return 0;
    ORC_CATCH
}
int orc_mat_addrow(void* h, int64_t row, const int64_t* colids, const double* vals, int64_t n) {
    ORC_TRY
    matrix_addrow(*(Matrix*)h, row, std::vector<int64_t>(colids, colids + n), std::vector<double>(vals, vals + n));
    return 0;
    ORC_CATCH
}
int orc_mat_closefillmode(void* h) {
    ORC_TRY
    matrix_closefillmode(*(Matrix*)h);
    return 0;
    ORC_CATCH
}
// which: 0 = colmajor, 1 = rowmajor.  out = {capacity, seg, nsegs, nb_elements, height, nb_partitions, len(semaphores), m, n, fillmode}
void orc_mat_info(void* h, int which, int64_t* out) {
    Matrix& A = *(Matrix*)h;
    Mpcsc& M = which == 0 ? A.colmajor : A.rowmajor;
    pma_info(M.pcsc.pma, out);
    out[5] = M.pcsc.nb_partitions; out[6] = (int64_t)M.pcsc.semaphores.size();
    out[7] = A.m; out[8] = A.n; out[9] = A.fillmode;
}
void orc_mat_export(void* h, int which, uint8_t* tag, int64_t* key, double* val, int64_t* sem, int64_t* colkeys, uint8_t* collive) {
    Matrix& A = *(Matrix*)h;
    Mpcsc& M = which == 0 ? A.colmajor : A.rowmajor;
    export_pma(M.pcsc.pma, tag, key, val);
    std::copy(M.pcsc.semaphores.begin(), M.pcsc.semaphores.end(), sem);
    std::copy(M.col_keys.key.begin(), M.col_keys.key.end(), colkeys);
    std::copy(M.col_keys.live.begin(), M.col_keys.live.end(), collive);
}
// column / row gathers; two-call size query: returns count, fills up to cap entries
int64_t orc_mat_column(void* h, int which, int64_t col, int64_t* keys, double* vals, int64_t cap) {
    Matrix& A = *(Matrix*)h;
    std::vector<int64_t> k; std::vector<double> v;
    mpcsc_column(which == 0 ? A.colmajor : A.rowmajor, col, k, v);
    for (int64_t i = 0; i < (int64_t)k.size() && i < cap; ++i) { keys[i] = k[i]; vals[i] = v[i]; }
    return (int64_t)k.size();
}
int64_t orc_mat_row_scan(void* h, int which, int64_t row, int64_t* keys, double* vals, int64_t cap) {
    Matrix& A = *(Matrix*)h;
    std::vector<int64_t> k; std::vector<double> v;
    mpcsc_row(which == 0 ? A.colmajor : A.rowmajor, row, k, v);
    for (int64_t i = 0; i < (int64_t)k.size() && i < cap; ++i) { keys[i] = k[i]; vals[i] = v[i]; }
    return (int64_t)k.size();
}
// SpMSpV. trans = 0: mat * x (colmajor, operations.jl:14-18); trans = 1: transpose(mat) * x (rowmajor, operations.jl:26-30)
int64_t orc_mat_mul(void* h, int trans, const int64_t* xk, const double* xv, int64_t nx, int64_t* yk, double* yv, int64_t cap, int* err) {
    *err = 0;
    try {
        Matrix& A = *(Matrix*)h;
        std::vector<int64_t> k; std::vector<double> v;
        mpcsc_mul(trans == 0 ? A.colmajor : A.rowmajor, xk, xv, nx, k, v);
        for (int64_t i = 0; i < (int64_t)k.size() && i < cap; ++i) { yk[i] = k[i]; yv[i] = v[i]; }
        return (int64_t)k.size();
    } catch (const Error& e) { g_err = e.msg; *err = e.code; return -1; }
}
// dense-x convenience used by the CPU baseline: x_j = xd[j-1] for j in 1..nx (all stored); y dense of length ny
int orc_mat_mul_dense(void* h, int trans, const double* xd, int64_t nx, double* yd, int64_t ny) {
    ORC_TRY
    Matrix& A = *(Matrix*)h;
    std::vector<int64_t> xk((size_t)nx);
    for (int64_t j = 0; j < nx; ++j) xk[j] = j + 1;
    std::vector<int64_t> k; std::vector<double> v;
    mpcsc_mul(trans == 0 ? A.colmajor : A.rowmajor, xk.data(), xd, nx, k, v);
    std::fill(yd, yd + ny, 0.0);
    for (size_t i = 0; i < k.size(); ++i) if (k[i] >= 1 && k[i] <= ny) yd[k[i] - 1] = v[i];
    return 0;
    ORC_CATCH
}
// batch policy at matrix level (both orientations)
int orc_mat_set_batch_policy(void* h, const int64_t* rows, const int64_t* cols, const double* vals, int64_t n) {
    ORC_TRY
    Matrix& A = *(Matrix*)h;
    if (A.fillmode) throw Error{ERR_ERROR, "set_batch in fill mode"};
    for (int64_t i = 0; i < n; ++i) if (vals[i] != 0.0) { A.m = std::max(A.m, rows[i]); A.n = std::max(A.n, cols[i]); }
    policy::mpcsc_set_batch(A.colmajor, rows, cols, vals, n);
    policy::mpcsc_set_batch(A.rowmajor, cols, rows, vals, n);
    return 0;
    ORC_CATCH
}
int orc_mat_delete_columns_policy(void* h, int rows_instead, const int64_t* ids, int64_t n) {
    ORC_TRY
    Matrix& A = *(Matrix*)h;
    if (A.fillmode) throw Error{ERR_ERROR, "Cannot delete a column in fill mode"};
    if (rows_instead) policy::matrix_delete_partitions(A.rowmajor, A.colmajor, ids, n);
    else policy::matrix_delete_partitions(A.colmajor, A.rowmajor, ids, n);
    return 0;
    ORC_CATCH
}

}  // extern "C"
