// ORACLE — TEST INFRASTRUCTURE ONLY (see fields.hpp). CPU restatement of the reference's Lasso node,
// memory-checking prover/verifier, sumcheck and transcript. Every function cites the reference file:line it
// follows; anything that lives in the un-vendored `gkr` / `plonkish_backend` crates is restated from its published
// algorithm and guarded by a named switch in `Assumptions`. PARITY UNPINNED for those.
#pragma once
#include <algorithm>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "fields.hpp"

namespace hgo {

// ---------------------------------------------------------------- Appendix-B switches
struct Assumptions {
    // A3: wire format of a sumcheck round message. 0 = coefficients c0,c2..cd (c1 omitted, derived from the claim);
    //     1 = evaluations h(0),h(2)..h(d) (h(1) omitted).
    int a3_wire = 0;
    // A3': how the prover obtains h(1). 0 = claim - h(0) (what a prover that trusts the claim does);
    //      1 = computed from the tables. They differ because the lasso sumchecks are not eq-weighted.
    int a3_h1 = 0;
    // A5: distribute_powers(exprs, b) = sum_i b^i expr_i (1) or Horner with the first expression getting the
    //     highest power (0).
    int a5_ascending = 1;
};
inline Assumptions& assumptions() { static Assumptions a; return a; }

struct OracleError : std::runtime_error { using std::runtime_error::runtime_error; };

// ---------------------------------------------------------------- transcript (bfv-gkr/src/transcript.rs:117-203)
// Event recorder for the interchange dump (scripts/hg_dump.py, patches/hyper-greco-dump.diff): when enabled, every base-field
// squeeze and every base-field write of every Transcript is appended as one kind byte ('S' / 'W') followed by the element in
// proof encoding (to_repr reversed = big-endian, transcript.rs:183-188). Single-threaded use (the protocol driver is).
struct EventLog {
    bool on = false;
    std::vector<uint8_t> bytes;
};
inline EventLog& event_log() { static EventLog e; return e; }

template <class F> struct Transcript {
    typedef typename ExtOf<F>::type E;
    static void log_event(uint8_t kind, const F& f) {
        EventLog& e = event_log();
        if (!e.on) return;
        uint8_t b[F::REPR_BYTES]; f.to_repr_le(b);
        std::reverse(b, b + F::REPR_BYTES);
        e.bytes.push_back(kind);
        e.bytes.insert(e.bytes.end(), b, b + F::REPR_BYTES);
    }
    std::vector<uint8_t> pending;  // bytes absorbed into the hasher since the last reset
    std::vector<uint8_t> stream;   // proof bytes (write mode)
    const uint8_t* rd = nullptr; size_t rd_len = 0, rd_pos = 0;  // read mode
    size_t n_base_squeezed = 0;
    Transcript() {}
    Transcript(const uint8_t* proof, size_t len) : rd(proof), rd_len(len) {}
    // transcript.rs:199-203: hash = finalize_fixed_reset(); update(hash); fe_mod_from_le_bytes(hash)
    F squeeze_base() {
        uint8_t h[32];
        keccak256(pending.data(), pending.size(), h);
        pending.assign(h, h + 32);
        n_base_squeezed++;
        const F c = F::from_le_bytes_mod(h, 32);
        log_event('S', c);
        return c;
    }
    // transcript.rs:149-154
    E squeeze() {
        F b[2];
        for (int i = 0; i < E::DEGREE; i++) b[i] = squeeze_base();
        return E::from_bases(b);
    }
    std::vector<E> squeeze_n(size_t n) { std::vector<E> r(n); for (auto& x : r) x = squeeze(); return r; }
    // transcript.rs:156 common_felt is a no-op
    void common(const E&) {}
    // transcript.rs:183-196: to_repr() reversed (big-endian); Ext = its bases in order
    void write_base(const F& f) {
        uint8_t b[F::REPR_BYTES]; f.to_repr_le(b);
        std::reverse(b, b + F::REPR_BYTES);
        stream.insert(stream.end(), b, b + F::REPR_BYTES);
        log_event('W', f);
    }
    void write(const E& e) {
        F b[2]; e.as_bases(b);
        for (int i = 0; i < E::DEGREE; i++) write_base(b[i]);
    }
    void write_n(const std::vector<E>& v) { for (auto& e : v) write(e); }
    // transcript.rs:162-177
    F read_base() {
        if (rd_pos + F::REPR_BYTES > rd_len) throw OracleError("Transcript: unexpected end of proof");
        uint8_t b[F::REPR_BYTES]; memcpy(b, rd + rd_pos, F::REPR_BYTES); rd_pos += F::REPR_BYTES;
        std::reverse(b, b + F::REPR_BYTES);
        F f;
        if (!F::from_repr_le(b, &f)) throw OracleError("Transcript: Invalid field element read from stream");
        return f;
    }
    E read() {
        F b[2];
        for (int i = 0; i < E::DEGREE; i++) b[i] = read_base();
        return E::from_bases(b);
    }
    std::vector<E> read_n(size_t n) { std::vector<E> r(n); for (auto& x : r) x = read(); return r; }
};

// ---------------------------------------------------------------- range lookups and subtables (lasso/src/table/range.rs)
static inline int ilog2_u64(uint64_t x) { return 63 - __builtin_clzll(x); }

struct Subtable {
    bool full; uint64_t bound;
    std::string id() const { return full ? "full" : "bound_" + std::to_string(bound); }  // range.rs:40-42,163-165
    // range.rs:58-62 (Q5): cutoff = 2^(ilog2(B) % log2M) + B % M
    uint64_t cutoff(int log2M) const {
        if (full) return 1ULL << log2M;
        int bb = ilog2_u64(bound);
        return (1ULL << (bb % log2M)) + bound % (1ULL << log2M);
    }
    // range.rs:15-17, 58-72
    template <class F> std::vector<F> materialize(int log2M) const {
        uint64_t M = 1ULL << log2M, c = cutoff(log2M);
        std::vector<F> t(M);
        for (uint64_t i = 0; i < M; i++) t[i] = (full || i < c) ? F::from_u64(i) : F::zero();
        return t;
    }
    // range.rs:19-26, 74-112
    template <class F, class E> E evaluate_mle(const std::vector<E>& point, int log2M) const {
        size_t b = point.size();
        E result = E::zero();
        if (full) {
            for (size_t i = 0; i < b; i++) result += point[i] * F::from_u64(1ULL << i);
            return result;
        }
        uint64_t cut = cutoff(log2M);
        size_t cutoff_log2 = ilog2_u64(cut);
        uint64_t g_base = 1ULL << cutoff_log2, num_extra = cut - g_base;
        for (size_t i = 0; i < b; i++) {
            if (i < cutoff_log2) {
                result += point[i] * F::from_u64(1ULL << i);
            } else {
                E g_value = E::zero();
                if (i == cutoff_log2) {
                    for (uint64_t k = 0; k < num_extra; k++) {
                        E term = E::from_base(F::from_u64(g_base + k));
                        for (size_t j = 0; j < cutoff_log2; j++)
                            term *= ((k >> j) & 1) ? point[j] : (E::one() - point[j]);
                        g_value += term;
                    }
                }
                result = (E::one() - point[i]) * result + point[i] * g_value;
            }
        }
        return result;
    }
};

struct RangeLookup {
    uint64_t bound;
    std::string id() const { return "range_" + std::to_string(bound); }  // range.rs:256-258
    // range.rs:207-228: (subtable, dimension indices)
    std::vector<std::pair<Subtable, std::vector<int>>> subtables(int log2M) const {
        uint64_t M = 1ULL << log2M;
        int bound_bits = ilog2_u64(bound), num_chunks = bound_bits / log2M;
        std::vector<int> fullr; for (int i = 0; i < num_chunks; i++) fullr.push_back(i);
        Subtable full{true, 0}, rem{false, bound};
        if (bound % M == 0) return {{full, fullr}};
        if (bound < M) return {{rem, {0}}};
        return {{full, fullr}, {rem, {num_chunks}}};
    }
    // range.rs:234-250
    std::vector<int> chunk_bits(int log2M) const {
        uint64_t M = 1ULL << log2M;
        int bound_bits = ilog2_u64(bound);
        std::vector<int> r(bound_bits / log2M, log2M);
        if (bound % M != 0) {
            uint64_t cut = (1ULL << (bound_bits % log2M)) + bound % M;
            r.push_back(ilog2_u64(cut));
        }
        return r;
    }
};

// ---------------------------------------------------------------- preprocessing (lasso/src/lasso.rs:527-627)
struct Preprocessing {
    int C = 4, log2M = 16;
    std::vector<RangeLookup> lookups;                 // BTreeMap order: sorted by id STRING, deduplicated (F7)
    std::map<std::string, int> lookup_id_to_index;
    std::vector<Subtable> subtables;                  // unique_by id, first appearance
    std::vector<std::vector<int>> subtable_to_memory_indices, lookup_to_memory_indices;
    std::vector<int> memory_to_subtable_index, memory_to_dimension_index;
    int num_memories = 0;

    static Preprocessing preprocess(const std::vector<uint64_t>& bounds, int C, int log2M) {
        Preprocessing pp; pp.C = C; pp.log2M = log2M;
        std::map<std::string, RangeLookup> m;  // lasso.rs:530-534
        for (uint64_t b : bounds) m[RangeLookup{b}.id()] = RangeLookup{b};
        for (auto& kv : m) { pp.lookup_id_to_index[kv.first] = (int)pp.lookups.size(); pp.lookups.push_back(kv.second); }
        std::map<std::string, int> sid;  // lasso.rs:543-552
        for (auto& l : pp.lookups)
            for (auto& st : l.subtables(log2M))
                if (!sid.count(st.first.id())) { sid[st.first.id()] = (int)pp.subtables.size(); pp.subtables.push_back(st.first); }
        std::vector<std::vector<bool>> dims(pp.subtables.size(), std::vector<bool>(64, false));  // lasso.rs:555-572
        for (auto& l : pp.lookups)
            for (auto& st : l.subtables(log2M))
                for (int d : st.second) dims[sid[st.first.id()]][d] = true;
        for (size_t s = 0; s < pp.subtables.size(); s++) {  // lasso.rs:578-586
            std::vector<int> mem;
            for (int d = 0; d < 64; d++) if (dims[s][d]) {
                mem.push_back(pp.num_memories++);
                pp.memory_to_subtable_index.push_back((int)s);
                pp.memory_to_dimension_index.push_back(d);
            }
            pp.subtable_to_memory_indices.push_back(mem);
        }
        pp.lookup_to_memory_indices.resize(pp.lookups.size());  // lasso.rs:590-602
        for (size_t li = 0; li < pp.lookups.size(); li++)
            for (auto& st : pp.lookups[li].subtables(log2M))
                for (int mi : pp.subtable_to_memory_indices[sid[st.first.id()]])
                    if (std::find(st.second.begin(), st.second.end(), pp.memory_to_dimension_index[mi]) != st.second.end())
                        pp.lookup_to_memory_indices[li].push_back(mi);
        return pp;
    }
};

template <class E, class T> struct Lift { static inline E f(const T& x) { return E::from_base(x); } };
template <class E> struct Lift<E, E> { static inline E f(const E& x) { return x; } };
template <class E, class T> static inline E lift(const T& x) { return Lift<E, T>::f(x); }

// ---------------------------------------------------------------- multilinear helpers [UPSTREAM gkr::poly, A4: LSB first]
template <class E, class T> E mle_evaluate(const std::vector<T>& evals, const std::vector<E>& point) {
    // evaluate(point): point[0] binds the lowest index bit
    size_t n = evals.size();
    if (n != (size_t)1 << point.size()) throw OracleError("mle_evaluate: size mismatch");
    if (point.empty()) return lift<E, T>(evals[0]);
    std::vector<E> cur(n >> 1);
    {
        const E r = point[0]; size_t h = n >> 1;
#pragma omp parallel for schedule(static) if (h >= 4096)
        for (size_t i = 0; i < h; i++) { E lo = lift<E, T>(evals[2 * i]); cur[i] = lo + r * (lift<E, T>(evals[2 * i + 1]) - lo); }
    }
    for (size_t v = 1; v < point.size(); v++) {
        size_t h = cur.size() >> 1; const E r = point[v];
        std::vector<E> nx(h);
#pragma omp parallel for schedule(static) if (h >= 4096)
        for (size_t i = 0; i < h; i++) nx[i] = cur[2 * i] + r * (cur[2 * i + 1] - cur[2 * i]);
        cur.swap(nx);
    }
    return cur[0];
}
template <class E> E mle_evaluate_ext(const std::vector<E>& evals, const std::vector<E>& point) {
    std::vector<E> cur = evals;
    for (size_t v = 0; v < point.size(); v++) {
        size_t h = cur.size() >> 1; const E r = point[v];
        std::vector<E> nx(h);
#pragma omp parallel for schedule(static) if (h >= 4096)
        for (size_t i = 0; i < h; i++) nx[i] = cur[2 * i] + r * (cur[2 * i + 1] - cur[2 * i]);
        cur.swap(nx);
    }
    return cur[0];
}
// plonkish MultilinearPolynomial::eq_xy (A10): eq[k] = prod_i (k_i ? r_i : 1 - r_i), k_0 = LSB
template <class E> std::vector<E> eq_xy(const std::vector<E>& r) {
    std::vector<E> eq(1, E::one());
    for (size_t i = 0; i < r.size(); i++) {
        size_t n = eq.size();
        eq.resize(2 * n);
        for (size_t k = 0; k < n; k++) { E hi = eq[k] * r[i]; eq[k + n] = hi; eq[k] = eq[k] - hi; }
    }
    return eq;
}

// ---------------------------------------------------------------- sumcheck [UPSTREAM gkr::sum_check, A3/A4/A5]
// The only function shape the lasso crate builds (lasso.rs:457-475, prover.rs:268-279):
//     g = poly(0) * sum_i coeffs[i] * prod_{k<arity} poly(arity*i + k)
// arity 1 = collation (degree 2), arity 2 = grand product (degree 3). poly(0) is the first DATA table (F4).
template <class E> struct SumcheckFn {
    int num_vars, arity;
    std::vector<E> coeffs;
    int degree() const { return arity + 1; }
    size_t num_polys() const { return coeffs.size() * arity; }
};
template <class E> std::vector<E> distribute_powers_coeffs(size_t n, E base) {
    std::vector<E> c(n); E p = E::one();
    for (size_t i = 0; i < n; i++) { c[i] = p; p *= base; }
    if (!assumptions().a5_ascending) std::reverse(c.begin(), c.end());
    return c;
}
// coefficients of the polynomial through (0,y0),(1,y1)..(d,yd)
template <class E, class F> std::vector<E> interpolate_coeffs(const std::vector<E>& y) {
    int n = (int)y.size();
    std::vector<std::vector<E>> a(n, std::vector<E>(n + 1));
    for (int i = 0; i < n; i++) {
        E p = E::one(), x = E::from_base(F::from_u64(i));
        for (int j = 0; j < n; j++) { a[i][j] = p; p *= x; }
        a[i][n] = y[i];
    }
    for (int c = 0; c < n; c++) {
        int piv = c; while (a[piv][c] == E::zero()) piv++;
        std::swap(a[piv], a[c]);
        E iv = a[c][c].inv();
        for (int j = c; j <= n; j++) a[c][j] *= iv;
        for (int r = 0; r < n; r++) if (r != c && a[r][c] != E::zero()) {
            E f = a[r][c];
            for (int j = c; j <= n; j++) a[r][j] -= f * a[c][j];
        }
    }
    std::vector<E> co(n); for (int i = 0; i < n; i++) co[i] = a[i][n];
    return co;
}
template <class E> E horner(const std::vector<E>& c, E x) {
    E r = E::zero(); for (size_t i = c.size(); i-- > 0;) r = r * x + c[i]; return r;
}
template <class E, class F> E lagrange_eval(const std::vector<E>& y, E x) {
    return horner(interpolate_coeffs<E, F>(y), x);
}

// evaluations of the true round polynomial at X = 0..degree for tables of current length 2h
template <class F, class E, class T> std::vector<E> round_evals(const SumcheckFn<E>& g, const std::vector<std::vector<T>>& t) {
    const int d = g.degree(), ar = g.arity; const size_t h = t[0].size() >> 1, nterm = g.coeffs.size();
    std::vector<E> acc(d + 1, E::zero());
#pragma omp parallel if (h >= 1024)
    {
        std::vector<E> loc(d + 1, E::zero());
#pragma omp for schedule(static) nowait
        for (size_t b = 0; b < h; b++) {
            E inner[4] = {E::zero(), E::zero(), E::zero(), E::zero()};
            for (size_t i = 0; i < nterm; i++) {
                E prod[4] = {g.coeffs[i], g.coeffs[i], g.coeffs[i], g.coeffs[i]};
                for (int k = 0; k < ar; k++) {
                    const auto& tb = t[ar * i + k];
                    E lo = lift<E, T>(tb[2 * b]), df = lift<E, T>(tb[2 * b + 1]) - lo, v = lo;
                    for (int x = 0; x <= d; x++) { prod[x] *= v; v += df; }
                }
                for (int x = 0; x <= d; x++) inner[x] += prod[x];
            }
            E lo = lift<E, T>(t[0][2 * b]), df = lift<E, T>(t[0][2 * b + 1]) - lo, v = lo;
            for (int x = 0; x <= d; x++) { loc[x] += v * inner[x]; v += df; }
        }
#pragma omp critical
        for (int x = 0; x <= d; x++) acc[x] += loc[x];
    }
    return acc;
}
template <class E, class T> std::vector<std::vector<E>> fold_tables(const std::vector<std::vector<T>>& t, E r) {
    std::vector<std::vector<E>> out(t.size());
    for (size_t k = 0; k < t.size(); k++) {
        size_t h = t[k].size() >> 1; out[k].resize(h);
        const auto& tb = t[k]; auto& o = out[k];
#pragma omp parallel for schedule(static) if (h >= 4096)
        for (size_t b = 0; b < h; b++) { E lo = lift<E, T>(tb[2 * b]); o[b] = lo + r * (lift<E, T>(tb[2 * b + 1]) - lo); }
    }
    return out;
}
// prove_sum_check(&g, claim, polys, transcript) -> (claim, r, evals)   [call sites lasso.rs:278-279, prover.rs:251]
template <class F, class E, class T>
void prove_sum_check(const SumcheckFn<E>& g, E claim, const std::vector<std::vector<T>>& polys, Transcript<F>& tr,
                     E* out_claim, std::vector<E>* out_r, std::vector<E>* out_evals,
                     std::vector<std::vector<E>>* true_evals_trace = nullptr) {
    const int d = g.degree();
    std::vector<std::vector<E>> cur;
    out_r->clear();
    for (int round = 0; round < g.num_vars; round++) {
        std::vector<E> ev = round == 0 ? round_evals<F, E, T>(g, polys) : round_evals<F, E, E>(g, cur);
        if (true_evals_trace) true_evals_trace->push_back(ev);
        if (assumptions().a3_h1 == 0) ev[1] = claim - ev[0];
        E r_i;
        if (assumptions().a3_wire == 0) {
            std::vector<E> co = interpolate_coeffs<E, F>(ev);
            tr.write(co[0]); for (int i = 2; i <= d; i++) tr.write(co[i]);
            r_i = tr.squeeze();
            claim = horner(co, r_i);
        } else {
            tr.write(ev[0]); for (int i = 2; i <= d; i++) tr.write(ev[i]);
            r_i = tr.squeeze();
            claim = lagrange_eval<E, F>(ev, r_i);
        }
        cur = round == 0 ? fold_tables<E, T>(polys, r_i) : fold_tables<E, E>(cur, r_i);
        out_r->push_back(r_i);
    }
    out_evals->resize(cur.size());
    for (size_t k = 0; k < cur.size(); k++) (*out_evals)[k] = cur[k][0];
    *out_claim = claim;
}
// verify_sum_check(&g, claim, transcript) -> (claim, r)   [call sites lasso.rs:130, verifier.rs:220]
template <class F, class E> void verify_sum_check(const SumcheckFn<E>& g, E claim, Transcript<F>& tr, E* out_claim, std::vector<E>* out_r) {
    const int d = g.degree();
    out_r->clear();
    for (int round = 0; round < g.num_vars; round++) {
        std::vector<E> m(d + 1);
        m[0] = tr.read(); for (int i = 2; i <= d; i++) m[i] = tr.read();
        E r_i = tr.squeeze();
        if (assumptions().a3_wire == 0) {
            E s = m[0].dbl(); for (int i = 2; i <= d; i++) s += m[i];
            m[1] = claim - s;
            claim = horner(m, r_i);
        } else {
            m[1] = claim - m[0];
            claim = lagrange_eval<E, F>(m, r_i);
        }
        out_r->push_back(r_i);
    }
    *out_claim = claim;
}

// ---------------------------------------------------------------- Lasso node (lasso/src/lasso.rs)
template <class F> struct LassoPolys {  // lasso.rs:479-510
    std::vector<std::vector<uint64_t>> dims;       // C x num_reads (usize)
    std::vector<std::vector<uint64_t>> read_cts;   // num_memories x num_reads
    std::vector<std::vector<uint64_t>> final_cts;  // num_memories x M
    std::vector<std::vector<F>> e_polys;           // num_memories x num_reads
    std::vector<int> row_lookup;                   // flag polynomials as one lookup index per row (-1 = padding)
    std::vector<F> lookup_outputs;
};

template <class F> struct LassoNode {
    typedef typename ExtOf<F>::type E;
    Preprocessing pp;
    int num_vars;
    std::vector<int> lookups;  // per row: index into pp.lookups (Vec<LookupId> of lasso.rs:35)
    std::vector<std::vector<F>> materialized;  // lasso.rs:604-609

    LassoNode(const Preprocessing& p, int nv, const std::vector<int>& rows) : pp(p), num_vars(nv), lookups(rows) {
        for (auto& s : pp.subtables) materialized.push_back(s.materialize<F>(pp.log2M));
    }

    // lasso.rs:381-414 + fe_to_bits_le :654-669 + range.rs:234-254 (Q8: silent truncation to sum(chunk_bits))
    std::vector<std::vector<uint64_t>> subtable_lookup_indices(const std::vector<F>& inputs) const {
        size_t rows = std::min(inputs.size(), lookups.size());
        std::vector<std::vector<uint64_t>> idx(pp.C, std::vector<uint64_t>(rows, 0));
        std::vector<int> total_bits(pp.lookups.size());
        for (size_t l = 0; l < pp.lookups.size(); l++) { int s = 0; for (int b : pp.lookups[l].chunk_bits(pp.log2M)) s += b; total_bits[l] = s; }
#pragma omp parallel for schedule(static)
        for (size_t i = 0; i < rows; i++) {
            uint8_t repr[F::REPR_BYTES]; inputs[i].to_repr_le(repr);
            int tb = total_bits[lookups[i]];
            for (int c = 0; c < pp.C; c++) {
                uint64_t v = 0;
                for (int k = 0; k < pp.log2M; k++) {
                    int bit = c * pp.log2M + k;
                    if (bit < tb && bit < 8 * F::REPR_BYTES && ((repr[bit >> 3] >> (bit & 7)) & 1)) v |= 1ULL << k;
                }
                idx[c][i] = v;
            }
        }
        return idx;
    }

    // lasso.rs:157-250
    LassoPolys<F> polynomialize(const std::vector<F>& inputs) const {
        LassoPolys<F> P;
        size_t num_reads = 1; while (num_reads < inputs.size()) num_reads <<= 1;
        size_t M = (size_t)1 << pp.log2M;
        auto sli = subtable_lookup_indices(inputs);
        size_t rows = std::min(inputs.size(), lookups.size());
        P.read_cts.resize(pp.num_memories); P.final_cts.resize(pp.num_memories); P.e_polys.resize(pp.num_memories);
#pragma omp parallel for schedule(dynamic)
        for (int mi = 0; mi < pp.num_memories; mi++) {  // lasso.rs:170-204
            int dim = pp.memory_to_dimension_index[mi], st = pp.memory_to_subtable_index[mi];
            std::vector<uint8_t> used(pp.lookups.size(), 0);
            for (size_t l = 0; l < pp.lookups.size(); l++)
                for (int x : pp.lookup_to_memory_indices[l]) if (x == mi) used[l] = 1;
            std::vector<uint64_t> fin(M, 0), rd(num_reads, 0);
            std::vector<F> e(num_reads, F::zero());
            for (size_t j = 0; j < rows; j++) if (used[lookups[j]]) {
                uint64_t a = sli[dim][j];
                uint64_t c = fin[a]; rd[j] = c; fin[a] = c + 1;
                e[j] = materialized[st][a];
            }
            P.read_cts[mi].swap(rd); P.final_cts[mi].swap(fin); P.e_polys[mi].swap(e);
        }
        P.dims.resize(pp.C);  // lasso.rs:216-223
        for (int c = 0; c < pp.C; c++) { P.dims[c] = sli[c]; P.dims[c].resize(num_reads, 0); }
        P.row_lookup.assign(num_reads, -1);  // lasso.rs:225-235
        for (size_t j = 0; j < rows; j++) P.row_lookup[j] = lookups[j];
        P.lookup_outputs.assign(num_reads, F::zero());  // lasso.rs:237-239, range.rs:230-232
        for (size_t j = 0; j < rows; j++) P.lookup_outputs[j] = inputs[j];
        return P;
    }

    // lasso.rs:422-454 with range.rs:184-195
    E sum_check_claim(const std::vector<E>& r, const LassoPolys<F>& P) const {
        std::vector<E> eq = eq_xy(r);
        size_t n = eq.size();
        F weight = F::from_u64(1ULL << pp.log2M);
        E claim = E::zero();
#pragma omp parallel
        {
            E loc = E::zero();
#pragma omp for schedule(static) nowait
            for (size_t k = 0; k < n; k++) {
                int l = P.row_lookup[k];
                if (l < 0) continue;
                F comb = F::zero(), w = F::one();
                for (int mi : pp.lookup_to_memory_indices[l]) { comb += P.e_polys[mi][k] * w; w *= weight; }
                loc += eq[k] * comb;
            }
#pragma omp critical
            claim += loc;
        }
        return claim;
    }

    // lasso.rs:457-475 (Q1)
    SumcheckFn<E> collation_sum_check_function() const {
        SumcheckFn<E> g; g.num_vars = num_vars; g.arity = 1;
        g.coeffs = distribute_powers_coeffs<E>(pp.num_memories, E::from_base(F::from_u64(1ULL << pp.log2M)));
        return g;
    }

    struct Chunk { int index; std::vector<int> memories; };
    std::vector<Chunk> chunks() const {  // lasso.rs:303-336
        std::map<int, Chunk> m;
        for (int mi = 0; mi < pp.num_memories; mi++) {
            int d = pp.memory_to_dimension_index[mi];
            m[d].index = d; m[d].memories.push_back(mi);
        }
        std::vector<Chunk> r; for (auto& kv : m) r.push_back(kv.second);
        return r;
    }

    static inline F hash(F a, F v, F t, F gamma, F gamma2, F tau) { return a + v * gamma + t * gamma2 - tau; }  // prover.rs:44

    // prover.rs:183-266
    void prove_grand_product(std::vector<std::vector<F>>& vs, Transcript<F>& tr, std::vector<E>* out_x) const {
        size_t nb = vs.size();
        // layers[0] = bottom (prover.rs:310-315); up = elementwise product of the halves (prover.rs:332-354)
        std::vector<std::vector<std::vector<F>>> layers;  // layers[k][i] = vector i at level k (halves = v_l, v_r)
        layers.push_back(std::move(vs));
        while (layers.back()[0].size() > 2) {
            auto& lo = layers.back();
            std::vector<std::vector<F>> up(nb);
            for (size_t i = 0; i < nb; i++) {
                size_t h = lo[i].size() >> 1; up[i].resize(h);
                const auto& v = lo[i]; auto& u = up[i];
#pragma omp parallel for schedule(static) if (h >= 4096)
                for (size_t k = 0; k < h; k++) u[k] = v[k] * v[k + h];
            }
            layers.push_back(std::move(up));
        }
        std::vector<E> claimed(nb);  // prover.rs:197-221
        for (size_t i = 0; i < nb; i++) { claimed[i] = E::from_base(layers.back()[i][0] * layers.back()[i][1]); tr.write(claimed[i]); }
        std::vector<E> x;
        for (size_t li = layers.size(); li-- > 0;) {  // prover.rs:223-265
            auto& L = layers[li];
            size_t half = L[0].size() >> 1;
            int nv = 0; while (((size_t)1 << nv) < half) nv++;
            std::vector<E> evals;
            if (nv == 0) {
                x.clear();
                for (size_t i = 0; i < nb; i++) { evals.push_back(E::from_base(L[i][0])); evals.push_back(E::from_base(L[i][1])); }
            } else {
                E gamma = tr.squeeze();
                SumcheckFn<E> g; g.num_vars = nv; g.arity = 2; g.coeffs = distribute_powers_coeffs<E>(nb, gamma);
                E claim = E::zero();  // prover.rs:281-286
                { E p = E::one(); for (size_t i = 0; i < nb; i++) { claim += claimed[i] * p; p *= gamma; } }
                std::vector<std::vector<F>> polys(2 * nb);
                for (size_t i = 0; i < nb; i++) {
                    polys[2 * i].assign(L[i].begin(), L[i].begin() + half);
                    polys[2 * i + 1].assign(L[i].begin() + half, L[i].end());
                }
                E c2; prove_sum_check<F, E, F>(g, claim, polys, tr, &c2, &x, &evals);
            }
            tr.write_n(evals);
            E mu = tr.squeeze();
            for (size_t i = 0; i < nb; i++) claimed[i] = evals[2 * i] + mu * (evals[2 * i + 1] - evals[2 * i]);  // prover.rs:288-294
            x.push_back(mu);
            std::vector<std::vector<F>>().swap(L);
        }
        *out_x = x;
    }

    // lasso.rs:57-114 -> claims for input 0: (r, claimed_sum)
    void prove_claim_reduction(const std::vector<F>& inputs, Transcript<F>& tr, std::vector<E>* out_r, E* out_claimed_sum) const {
        LassoPolys<F> P = polynomialize(inputs);
        if ((size_t)1 << num_vars != P.lookup_outputs.size()) throw OracleError("assert_eq!(num_vars, self.num_vars)");
        std::vector<E> r = tr.squeeze_n(num_vars);  // lasso.rs:85
        E claimed_sum = sum_check_claim(r, P);     // lasso.rs:264
        tr.write(claimed_sum);                     // lasso.rs:269
        {
            SumcheckFn<E> g = collation_sum_check_function();
            E c2; std::vector<E> rx, ev;
            prove_sum_check<F, E, F>(g, claimed_sum, P.e_polys, tr, &c2, &rx, &ev);  // lasso.rs:278-279
        }
        E gamma_e = tr.squeeze(), tau_e = tr.squeeze();  // lasso.rs:99
        F gamma = gamma_e.base0(), tau = tau_e.base0(), gamma2 = gamma.square();  // prover.rs:38-39 (Q4)
        auto ch = chunks();
        size_t R = P.lookup_outputs.size(), M = (size_t)1 << pp.log2M;
        std::vector<int> order;  // chunk-major memory order (Q10)
        for (auto& c : ch) for (int mi : c.memories) order.push_back(mi);
        size_t m = order.size();
        std::vector<E> x, y;
        {
            std::vector<std::vector<F>> vs(2 * m);  // reads || writes, prover.rs:161-165
            size_t pos = 0;
            for (auto& c : ch) for (int mi : c.memories) {
                const auto& dim = P.dims[c.index]; const auto& ts = P.read_cts[c.index];  // F6: lasso.rs:318
                const auto& e = P.e_polys[mi];
                auto& rd = vs[pos]; auto& wr = vs[m + pos]; rd.resize(R); wr.resize(R);
#pragma omp parallel for schedule(static)
                for (size_t j = 0; j < R; j++) {  // prover.rs:70-73
                    F a = F::from_u64(dim[j]), t = F::from_u64(ts[j]);
                    rd[j] = hash(a, e[j], t, gamma, gamma2, tau);
                    wr[j] = hash(a, e[j], t + F::one(), gamma, gamma2, tau);
                }
                pos++;
            }
            prove_grand_product(vs, tr, &x);
        }
        {
            std::vector<std::vector<F>> vs(2 * m);  // inits || final_reads, prover.rs:167-171
            size_t pos = 0;
            for (auto& c : ch) for (int mi : c.memories) {
                const auto& T = materialized[pp.memory_to_subtable_index[mi]]; const auto& fc = P.final_cts[c.index];  // F6: lasso.rs:319
                auto& in = vs[pos]; auto& fr = vs[m + pos]; in.resize(M); fr.resize(M);
                for (size_t i = 0; i < M; i++) {  // prover.rs:62-69
                    in[i] = hash(F::from_u64(i), T[i], F::zero(), gamma, gamma2, tau);
                    fr[i] = hash(F::from_u64(i), T[i], F::from_u64(fc[i]), gamma, gamma2, tau);
                }
                pos++;
            }
            prove_grand_product(vs, tr, &y);
        }
        for (auto& c : ch) {  // prover.rs:173-178, mod.rs:80-93
            tr.write(mle_evaluate<E, F>(wrap(P.dims[c.index]), x));
            tr.write(mle_evaluate<E, F>(wrap(P.read_cts[c.index]), x));
            tr.write(mle_evaluate<E, F>(wrap(P.final_cts[c.index]), y));
            for (int mi : c.memories) tr.write(mle_evaluate<E, F>(P.e_polys[mi], x));
        }
        *out_r = r; *out_claimed_sum = claimed_sum;
    }

    // helpers to evaluate usize tables as field tables (DensePolynomial::from_usize, lasso.rs:199-201,221)
    static std::vector<F> wrap(const std::vector<uint64_t>& v) {
        std::vector<F> r(v.size());
#pragma omp parallel for schedule(static) if (v.size() >= 4096)
        for (size_t i = 0; i < v.size(); i++) r[i] = F::from_u64(v[i]);
        return r;
    }

    // ------------------------------------------------------------ verifier
    // verifier.rs:178-235
    void verify_grand_product(int num_vars_gp, size_t nb, Transcript<F>& tr, std::vector<E>* out_claims, std::vector<E>* out_x) const {
        std::vector<E> claimed = tr.read_n(nb);
        std::vector<E> x;
        for (int nv = 0; nv < num_vars_gp; nv++) {
            std::vector<E> evals;
            if (nv == 0) {
                evals = tr.read_n(2 * nb);
                for (size_t i = 0; i < nb; i++)
                    if (claimed[i] != evals[2 * i] * evals[2 * i + 1]) throw OracleError("InvalidSumCheck: unmatched sum check output");
                x.clear();
            } else {
                E gamma = tr.squeeze();
                SumcheckFn<E> g; g.num_vars = nv; g.arity = 2; g.coeffs = distribute_powers_coeffs<E>(nb, gamma);
                E claim = E::zero();
                { E p = E::one(); for (size_t i = 0; i < nb; i++) { claim += claimed[i] * p; p *= gamma; } }
                E c2; verify_sum_check<F, E>(g, claim, tr, &c2, &x);  // final claim discarded: verifier.rs:218-221
                evals = tr.read_n(2 * nb);
            }
            E mu = tr.squeeze();
            for (size_t i = 0; i < nb; i++) claimed[i] = evals[2 * i] + mu * (evals[2 * i + 1] - evals[2 * i]);
            x.push_back(mu);
        }
        *out_claims = claimed; *out_x = x;
    }

    // lasso.rs:116-139 + :342-378 + verifier.rs:130-176 + :61-95
    void verify_claim_reduction(Transcript<F>& tr, std::vector<E>* out_r, E* out_claimed_sum) const {
        std::vector<E> r = tr.squeeze_n(num_vars);
        SumcheckFn<E> g = collation_sum_check_function();
        E claimed_sum = tr.read();
        { E c2; std::vector<E> rx; verify_sum_check<F, E>(g, claimed_sum, tr, &c2, &rx); }  // result discarded: lasso.rs:129-130
        E gamma_e = tr.squeeze(), tau_e = tr.squeeze();
        F gamma = gamma_e.base0(), tau = tau_e.base0(), gamma2 = gamma.square();
        auto ch = chunks();
        size_t m = pp.num_memories;
        std::vector<E> rw, x, ify, y;
        verify_grand_product(num_vars, 2 * m, tr, &rw, &x);
        verify_grand_product(pp.log2M, 2 * m, tr, &ify, &y);
        auto hashE = [&](E a, E v, E t) { return a + v * gamma + t * gamma2 - E::from_base(tau); };  // verifier.rs:158
        size_t off = 0;
        for (auto& c : ch) {  // verifier.rs:61-95
            E dim_x = tr.read(), ts_x = tr.read(), fc_y = tr.read();
            std::vector<E> e_x = tr.read_n(c.memories.size());
            E id_y = E::zero(); { E p2 = E::one(); for (auto& yi : y) { id_y += p2 * yi; p2 = p2.dbl(); } }  // verifier.rs:74-78
            for (size_t i = 0; i < c.memories.size(); i++) {
                if (rw[off + i] != hashE(dim_x, e_x[i], ts_x)) throw OracleError("verify_memories: read_xs mismatch");
                if (rw[m + off + i] != hashE(dim_x, e_x[i], ts_x + E::one())) throw OracleError("verify_memories: write_xs mismatch");
                E st_y = pp.subtables[pp.memory_to_subtable_index[c.memories[i]]].template evaluate_mle<F, E>(y, pp.log2M);
                if (ify[off + i] != hashE(id_y, st_y, E::zero())) throw OracleError("verify_memories: init_ys mismatch");
                if (ify[m + off + i] != hashE(id_y, st_y, fc_y)) throw OracleError("verify_memories: final_read_ys mismatch");
            }
            off += c.memories.size();
        }
        *out_r = r; *out_claimed_sum = claimed_sum;
    }
};

}  // namespace hgo

// ---------------------------------------------------------------- NTT + forward circuit evaluation
// [UPSTREAM gkr FftNode / VanillaNode::evaluate, assumption A9]: radix-2 NTT over w = ROOT_OF_UNITY^(2^(S - log2 n)),
// natural order in and out, inverse scaled by n^{-1}. Call sites: sk_encryption_circuit.rs:224,249,251.
namespace hgo {
template <class F> void ntt_naive_order(std::vector<F>& a, bool inverse) {
    size_t n = a.size();
    int lg = 0; while (((size_t)1 << lg) < n) lg++;
    F w = F::root_of_unity();
    for (int i = lg; i < F::TWO_ADICITY; i++) w = w.square();
    if (inverse) w = w.inv();
    // bit reversal + iterative Cooley-Tukey
    for (size_t i = 1, j = 0; i < n; i++) {
        size_t bit = n >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) std::swap(a[i], a[j]);
    }
    for (size_t len = 2; len <= n; len <<= 1) {
        F wl = w;
        for (size_t k = len; k < n; k <<= 1) wl = wl.square();
        for (size_t i = 0; i < n; i += len) {
            F x = F::one();
            for (size_t j = 0; j < len / 2; j++) {
                F u = a[i + j], v = a[i + j + len / 2] * x;
                a[i + j] = u + v; a[i + j + len / 2] = u - v;
                x *= wl;
            }
        }
    }
    if (inverse) { F ni = F::from_u64(n).inv(); for (auto& x : a) x *= ni; }
}

// Values of the circuit's layers that matter downstream (sk_encryption_circuit.rs:86-293 evaluated on get_inputs :365-415):
//   lasso_inputs = output of `lasso_inputs_batched` (:163-181), sum = output of `sum` (:280-285) which must equal ct0is.
template <class F> struct BfvParams { int log2_size, K; std::vector<F> qis, k0is; std::vector<uint64_t> r1_bounds, r2_bounds; uint64_t s_bound, e_bound, k1_bound; };
template <class F>
void bfv_evaluate(const BfvParams<F>& P, const std::vector<F>& s, const std::vector<F>& e, const std::vector<F>& k1,
                  const std::vector<std::vector<F>>& ais, const std::vector<std::vector<F>>& r1is, const std::vector<F>& r2is,
                  std::vector<F>* lasso_inputs, std::vector<F>* sum) {
    const size_t N2 = (size_t)1 << P.log2_size, n = N2 / 2;
    const int K = P.K;
    lasso_inputs->clear();
    for (int i = 0; i < K; i++) for (size_t j = 0; j < N2; j++) lasso_inputs->push_back(r1is[i][j] + F::from_u64(P.r1_bounds[i]));
    for (size_t c = 0; c * N2 < r2is.size() || c == 0; c++) {  // r2is_chunks (:150-161); Q7: every chunk shifted by R2_BOUNDS[0]
        for (size_t j = 0; j < N2; j++) {
            size_t q = c * N2 + j;
            F v = q < r2is.size() ? r2is[q] : F::zero();
            lasso_inputs->push_back(v + F::from_u64(P.r2_bounds[0]));
        }
        if ((c + 1) * N2 >= r2is.size()) break;
    }
    for (size_t j = 0; j < N2; j++) lasso_inputs->push_back(s[j] + F::from_u64(P.s_bound));
    for (size_t j = 0; j < N2; j++) lasso_inputs->push_back(e[j] + F::from_u64(P.e_bound));
    for (size_t j = 0; j < N2; j++) lasso_inputs->push_back(k1[j] + F::from_u64(P.k1_bound));
    std::vector<F> s_eval = s;
    ntt_naive_order(s_eval, false);
    sum->assign((size_t)K * N2, F::zero());
    for (int i = 0; i < K; i++) {
        std::vector<F> a = ais[i];
        ntt_naive_order(a, false);
        for (size_t j = 0; j < N2; j++) a[j] *= s_eval[j];
        ntt_naive_order(a, true);
        for (size_t j = 0; j < N2; j++) {
            F v = a[j] + e[j] + k1[j] * P.k0is[i] + r1is[i][j] * P.qis[i];
            // r2i_cyclo (:262-278): [r2i[0..n-2], 0, r2i[0..n-2], 0]
            size_t jj = j % n;
            if (jj < n - 1) v += r2is[(size_t)i * n + jj];
            (*sum)[(size_t)i * N2 + j] = v;
        }
    }
}
}  // namespace hgo
