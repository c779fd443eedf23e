// TEST INFRASTRUCTURE — NOT PRODUCT CODE.  See batch_policy.hpp.
#include "batch_policy.hpp"
#include <numeric>

namespace orc {
namespace policy {

// accept iff p <= cnt/wc <= t with p = p_0 + p_d*h, t = t_0 + t_d*h   (pma.jl:119-123)
void level_bounds(int64_t S, int64_t H, double t_d, double p_d, std::vector<int64_t>& mn, std::vector<int64_t>& mx) {
    const double t_0 = 0.92, p_0 = 0.08;
    mn.assign((size_t)H + 1, 0);
    mx.assign((size_t)H + 1, 0);
    for (int64_t h = 0; h <= H; ++h) {
        const int64_t wc = (int64_t(1) << h) * S;
        const double p = p_0 + p_d * (double)h;
        const double t = t_0 + t_d * (double)h;
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

static void set_geometry(Pma& p, int64_t capacity, int64_t height) {
    p.capacity = capacity;
    p.nb_segments = capacity / p.segment_capacity;
    p.height = height;
    p.t_d = (p.t_h - p.t_0) / (double)height;   // pma.jl:147-148,157-158
    p.p_d = (p.p_h - p.p_0) / (double)height;
}

// merged sequence (survivors + inserts) of the cells [ws, we]; inserts are emitted right after their predecessor position
static void gather_merged(const Pma& p, const std::vector<const Op*>& ins, size_t& ins_cursor, int64_t ws, int64_t we,
                          std::vector<KV>& out) {
    // inserts are sorted by (pid,key) => non-decreasing pos; cursor points at the first insert with pos >= ws-ish
    if (ws == 1) {
        while (ins_cursor < ins.size() && ins[ins_cursor]->pos == 0) {
            out.push_back(KV{ins[ins_cursor]->key, ins[ins_cursor]->val});
            ++ins_cursor;
        }
    }
    for (int64_t pos = ws; pos <= we; ++pos) {
        if (!p.array.empty_at(pos)) out.push_back(p.array.at(pos));
        while (ins_cursor < ins.size() && ins[ins_cursor]->pos == pos) {
            out.push_back(KV{ins[ins_cursor]->key, ins[ins_cursor]->val});
            ++ins_cursor;
        }
    }
}

static void relayout(Pma& p, Semaphores* sem, int64_t ws, int64_t we, const std::vector<KV>& items) {
    for (int64_t pos = ws; pos <= we; ++pos) p.array.set_nothing(pos);
    int64_t pos = ws;
    for (const KV& e : items) { p.array.set(pos, e.key, e.val); ++pos; }
    const int64_t m = (int64_t)items.size();
    if (sem) spread5(p.array, ws, we, m, sem);      // moves.jl:142 (refreshes semaphores[] for every element)
    else spread4(p.array, ws, we, m);               // moves.jl:120
}

void apply_located(Pma& p, Semaphores* sem, std::vector<Op>& ops,
                   const std::vector<std::pair<int64_t, int64_t>>& purge_ranges) {
    const int64_t S = p.segment_capacity;
    const int64_t nsegs = p.nb_segments;
    const int64_t H = p.height;
    std::vector<int64_t> leafcnt((size_t)nsegs, 0), inscnt((size_t)nsegs, 0);
    std::vector<uint8_t> touched((size_t)nsegs, 0);
    for (int64_t pos = 1; pos <= p.capacity; ++pos)
        if (!p.array.empty_at(pos)) leafcnt[(pos - 1) / S] += 1;
    // purge (writes.jl:80-92)
    for (auto& r : purge_ranges) {
        for (int64_t pos = r.first; pos <= r.second; ++pos) {
            if (!p.array.empty_at(pos)) {
                p.array.set_nothing(pos);
                leafcnt[(pos - 1) / S] -= 1;
                touched[(pos - 1) / S] = 1;
            }
        }
    }
    // hits: overwrite (writes.jl:16-19) / blank (writes.jl:65-68)
    std::vector<const Op*> ins;
    for (Op& op : ops) {
        if (op.hit) {
            if (op.kind == OP_SET) p.array.set(op.pos, op.key, op.val);
            else if (op.kind == OP_DEL) {
                p.array.set_nothing(op.pos);
                leafcnt[(op.pos - 1) / S] -= 1;
                touched[(op.pos - 1) / S] = 1;
            }
        } else if (op.kind == OP_SET || op.kind == OP_SEM) {
            int64_t leaf = (std::max<int64_t>(op.pos, 1) - 1) / S;
            inscnt[leaf] += 1;
            touched[leaf] = 1;
            ins.push_back(&op);
        }
    }
    for (size_t i = 1; i < ins.size(); ++i)
        if (ins[i]->pos < ins[i - 1]->pos) throw Error{ERR_ASSERT, "policy: insert predecessors not monotone"};
    // implicit tree of post-batch counts
    std::vector<std::vector<int64_t>> post((size_t)H + 1);
    post[0].resize((size_t)nsegs);
    for (int64_t l = 0; l < nsegs; ++l) post[0][l] = leafcnt[l] + inscnt[l];
    for (int64_t h = 1; h <= H; ++h) {
        post[h].resize((size_t)(nsegs >> h));
        for (int64_t w = 0; w < (nsegs >> h); ++w) post[h][w] = post[h - 1][2 * w] + post[h - 1][2 * w + 1];
    }
    std::vector<int64_t> mn, mx;
    level_bounds(S, H, p.t_d, p.p_d, mn, mx);
    std::vector<std::vector<uint8_t>> mark((size_t)H + 1);
    for (int64_t h = 0; h <= H; ++h) mark[h].assign((size_t)(nsegs >> h), 0);
    bool root_fail = false;
    for (int64_t l = 0; l < nsegs; ++l) {
        if (!touched[l]) continue;
        int64_t h = 0;
        for (; h <= H; ++h) {
            int64_t c = post[h][l >> h];
            if (mn[h] <= c && c <= mx[h]) break;
        }
        if (h > H) root_fail = true;
        else mark[h][l >> h] = 1;
    }
    const int64_t N = post[H][0];
    if (root_fail) {
        std::vector<KV> items;
        size_t cur = 0;
        gather_merged(p, ins, cur, 1, p.capacity, items);
        int64_t cap = p.capacity, hh = H;
        std::vector<int64_t> mn2, mx2;
        if (N > mx[H]) {
            do {   // _extend! pma.jl:143-151, repeated until the root accepts
                cap *= 2; hh += 1;
                level_bounds(S, hh, (p.t_h - p.t_0) / (double)hh, (p.p_h - p.p_0) / (double)hh, mn2, mx2);
            } while (N > mx2[hh]);
        } else {
            while (hh > 1) {   // _shrink! pma.jl:153-161 (only while height > 1, pma.jl:135)
                level_bounds(S, hh, (p.t_h - p.t_0) / (double)hh, (p.p_h - p.p_0) / (double)hh, mn2, mx2);
                if (N >= mn2[hh]) break;
                cap /= 2; hh -= 1;
            }
        }
        p.array = Elements((size_t)cap);
        set_geometry(p, cap, hh);
        relayout(p, sem, 1, cap, items);
    } else {
        size_t cur = 0;
        int64_t l = 0;
        while (l < nsegs) {
            int64_t fh = -1;
            for (int64_t h = H; h >= 0; --h)
                if (mark[h][l >> h]) { fh = h; break; }
            if (fh < 0) { l += 1; continue; }
            const int64_t w = l >> fh;
            const int64_t first_leaf = w << fh;
            const int64_t nleaves = int64_t(1) << fh;
            const int64_t ws = first_leaf * S + 1, we = (first_leaf + nleaves) * S;
            // position the insert cursor at the first insert of this window
            while (cur < ins.size() && std::max<int64_t>(ins[cur]->pos, 1) < ws) ++cur;   // (cannot skip: untouched windows have no inserts)
            if (fh == 0 && inscnt[l] == 0) { l += 1; continue; }   // leaf accepted, deletes only: nothing moves (pma.jl:96-99)
            std::vector<KV> items;
            gather_merged(p, ins, cur, ws, we, items);
            if ((int64_t)items.size() != post[fh][w]) throw Error{ERR_ASSERT, "policy: window count mismatch"};
            relayout(p, sem, ws, we, items);
            l = first_leaf + nleaves;
        }
    }
    p.nb_elements = N;
}

// ---- plain PMA -------------------------------------------------------------------------
void pma_set_batch(Pma& p, const int64_t* keys, const double* vals, int64_t n) {
    if (n == 1) {   // a batch of ONE op is the reference's own setindex! (pma.jl:196-213): shift to the next gap (writes.jl:26-43),
        pma_set(p, vals[0], keys[0]);   // one leaf -> root walk, at most one _extend! / _shrink!
        return;
    }
    std::vector<int64_t> idx((size_t)n);
    std::iota(idx.begin(), idx.end(), int64_t(0));
    std::stable_sort(idx.begin(), idx.end(), [&](int64_t a, int64_t b) { return keys[a] < keys[b]; });
    std::vector<Op> ops;
    for (int64_t i = 0; i < n; ++i) {
        if (i + 1 < n && keys[idx[i + 1]] == keys[idx[i]]) continue;   // last writer wins
        Op op;
        op.pid = 0; op.key = keys[idx[i]]; op.val = vals[idx[i]];
        op.kind = op.val != 0.0 ? OP_SET : OP_DEL;                    // pma.jl:197
        op.pos = find(p.array, op.key, 1, p.array.length());
        op.hit = op.pos != 0 && p.array.at(op.pos).key == op.key;
        ops.push_back(op);
    }
    apply_located(p, nullptr, ops, {});
}

// ---- MappedPackedCSC -------------------------------------------------------------------
static int64_t next_live_sem_pos(const Semaphores& s, int64_t pid, int64_t array_len) {   // pcsr.jl:177-186 -> position after the span
    for (int64_t q = pid + 1; q <= (int64_t)s.size(); ++q)
        if (s[q - 1] != 0) return s[q - 1];
    return array_len + 1;
}

void mpcsc_set_batch(Mpcsc& M, const int64_t* inkeys, const int64_t* partkeys, const double* vals, int64_t n) {
    for (int64_t i = 0; i < n; ++i)
        if (inkeys[i] < 1) throw Error{ERR_ARGUMENT, "in-array keys must be >= 1 (key 0 is the semaphore key, pcsr.jl:23)"};
    if (n == 1) {   // a batch of ONE op on an existing partition is the reference's own setindex! (pcsr.jl:341-347 -> 294-339)
        bool exact;
        colkeys_find(M.col_keys, partkeys[0], &exact);
        if (exact) {
            mpcsc_set(M, vals[0], inkeys[0], partkeys[0]);
            return;
        }   // a single op that creates its partition follows the batch policy below (semaphore + element laid by one spread!)
    }
    // 1. column map: replay addcolumn!'s slot logic (pcsr.jl:148-169) in arrival order, without touching the array
    ColKeys ck = M.col_keys;
    std::vector<int64_t> old_id((size_t)ck.length());
    for (int64_t s = 0; s < ck.length(); ++s) old_id[s] = ck.live[s] ? s + 1 : 0;
    int64_t nb_new = 0;
    for (int64_t i = 0; i < n; ++i) {
        bool exact;
        int64_t pos = colkeys_find(ck, partkeys[i], &exact);
        if (exact) continue;
        nb_new += 1;
        if (pos == ck.length()) {
            ck.key.push_back(partkeys[i]); ck.live.push_back(1); old_id.push_back(0);
        } else if (!ck.live[pos]) {
            ck.key[pos] = partkeys[i]; ck.live[pos] = 1; old_id[pos] = 0;
        } else {
            ck.key.insert(ck.key.begin() + pos, partkeys[i]);
            ck.live.insert(ck.live.begin() + pos, 1);
            old_id.insert(old_id.begin() + pos, 0);
        }
    }
    // 2. renumber: new partition id = slot index; rewrite moved semaphore cells (pcsr.jl:128-134)
    Semaphores nsem((size_t)ck.length(), 0);
    std::vector<uint8_t> is_new((size_t)ck.length(), 0);
    Pcsc& P = M.pcsc;
    for (int64_t s = 0; s < ck.length(); ++s) {
        if (!ck.live[s]) continue;
        if (old_id[s] == 0) { is_new[s] = 1; continue; }
        int64_t sp = P.semaphores[old_id[s] - 1];
        nsem[s] = sp;
        if (old_id[s] != s + 1) P.pma.array.set(sp, 0, (double)(s + 1));
    }
    P.semaphores = nsem;
    M.col_keys = ck;
    // 3. ops: last-writer-wins per (pid, key), sorted; new partitions contribute a semaphore insert
    std::vector<int64_t> pid((size_t)n);
    for (int64_t i = 0; i < n; ++i) {
        bool exact;
        pid[i] = colkeys_find(ck, partkeys[i], &exact);
        if (!exact) throw Error{ERR_ASSERT, "policy: column map"};
    }
    std::vector<int64_t> idx((size_t)n);
    std::iota(idx.begin(), idx.end(), int64_t(0));
    std::stable_sort(idx.begin(), idx.end(), [&](int64_t a, int64_t b) {
        if (pid[a] != pid[b]) return pid[a] < pid[b];
        return inkeys[a] < inkeys[b];
    });
    std::vector<Op> ops;
    const int64_t len = P.pma.array.length();
    size_t k = 0;
    std::vector<int64_t> new_slots;
    for (int64_t s = 0; s < ck.length(); ++s) if (is_new[s]) new_slots.push_back(s + 1);
    size_t ns = 0;
    auto emit_sems_upto = [&](int64_t upto_pid) {   // semaphores of new partitions with id <= upto_pid, in id order
        while (ns < new_slots.size() && new_slots[ns] <= upto_pid) {
            Op op;
            op.pid = new_slots[ns]; op.key = 0; op.val = (double)new_slots[ns]; op.kind = OP_SEM;
            op.pos = next_live_sem_pos(P.semaphores, op.pid, len) - 1;
            op.hit = false;
            ops.push_back(op);
            ++ns;
        }
    };
    while (k < (size_t)n) {
        size_t e = k;
        while (e + 1 < (size_t)n && pid[idx[e + 1]] == pid[idx[k]] && inkeys[idx[e + 1]] == inkeys[idx[k]]) ++e;
        const int64_t i = idx[e];   // last writer
        Op op;
        op.pid = pid[i]; op.key = inkeys[i]; op.val = vals[i];
        op.kind = op.val != 0.0 ? OP_SET : OP_DEL;   // pcsr.jl:301
        emit_sems_upto(op.pid);
        if (is_new[op.pid - 1]) {
            op.pos = next_live_sem_pos(P.semaphores, op.pid, len) - 1;
            op.hit = false;
        } else {
            int64_t from = P.semaphores[op.pid - 1];
            int64_t to = next_live_sem_pos(P.semaphores, op.pid, len) - 1;
            if (op.kind == OP_SET) {   // pcsr.jl:305 + writes.jl:14-19
                op.pos = find(P.pma.array, op.key, from + 1, to);
                op.hit = op.pos != 0 && P.pma.array.at(op.pos).key == op.key && from + 1 <= op.pos && op.pos <= to;
            } else {                   // pcsr.jl:307 + writes.jl:57-63
                op.pos = find(P.pma.array, op.key, from, to);
                op.hit = op.pos != 0 && P.pma.array.at(op.pos).key == op.key;
            }
        }
        if (!(op.kind == OP_DEL && !op.hit)) ops.push_back(op);
        k = e + 1;
    }
    emit_sems_upto((int64_t)ck.length());
    apply_located(P.pma, &P.semaphores, ops, {});
    P.nb_partitions += nb_new;
}

void matrix_delete_partitions(Mpcsc& primary, Mpcsc& twin, const int64_t* ids, int64_t n) {
    // validate before mutate: every id live and distinct (pcsr.jl:208 ArgumentError)
    std::vector<int64_t> slots;
    for (int64_t i = 0; i < n; ++i) {
        bool exact;
        int64_t s = colkeys_find(primary.col_keys, ids[i], &exact);
        if (!exact) throw Error{ERR_ARGUMENT, "column does not exist."};
        for (int64_t q : slots) if (q == s) throw Error{ERR_ARGUMENT, "column listed twice."};
        slots.push_back(s);
    }
    // entries to delete from the twin: (partition = in-array key of the primary, key = id)   matrix.jl:97-99
    std::vector<int64_t> tk, tp;
    std::vector<double> tv;
    std::vector<std::pair<int64_t, int64_t>> ranges;
    Pcsc& P = primary.pcsc;
    const int64_t len = P.pma.array.length();
    for (int64_t i = 0; i < n; ++i) {
        int64_t s = slots[i];
        int64_t from = P.semaphores[s - 1];
        int64_t to = next_live_sem_pos(P.semaphores, s, len) - 1;
        for (int64_t pos = from + 1; pos <= to; ++pos) {
            if (!P.pma.array.empty_at(pos)) {
                tk.push_back(ids[i]);
                tp.push_back(P.pma.array.at(pos).key);
                tv.push_back(0.0);
            }
        }
        ranges.push_back({from, to});
    }
    mpcsc_set_batch(twin, tk.data(), tp.data(), tv.data(), (int64_t)tk.size());
    std::vector<Op> none;
    apply_located(P.pma, &P.semaphores, none, ranges);
    for (int64_t s : slots) {
        P.semaphores[s - 1] = 0;              // pcsr.jl:202
        primary.col_keys.live[s - 1] = 0;     // pcsr.jl:209
    }
    P.nb_partitions -= n;                      // pcsr.jl:191
}

}  // namespace policy
}  // namespace orc
