// ORACLE — TEST INFRASTRUCTURE ONLY. C entry points so tests/, __graft_entry__.smoke() and bench.py's cpu_baseline
// leg can drive the CPU restatement through ctypes. Not part of the product; the product never links this.
//
// Encoding at this boundary: field id 0 = Goldilocks/GoldilocksExt2, 1 = BN254 Fr (E = F).
// Base element = LIMBS u64 (canonical integer, little-endian limbs); E element = DEGREE base elements.
#include <omp.h>

#include <chrono>
#include <cstdio>

#include "gkr.hpp"
#include "protocol.hpp"

using namespace hgo;

namespace {
thread_local std::string g_err;
template <class F> std::vector<F> load_base(const uint64_t* p, size_t n) {
    std::vector<F> v(n);
#pragma omp parallel for schedule(static) if (n >= 65536)
    for (size_t i = 0; i < n; i++) v[i] = F::from_limbs(p + i * F::LIMBS);
    return v;
}
template <class F, class E> E load_ext(const uint64_t* p) {
    F b[2];
    for (int i = 0; i < E::DEGREE; i++) b[i] = F::from_limbs(p + i * F::LIMBS);
    return E::from_bases(b);
}
template <class F, class E> void store_ext(const E& e, uint64_t* p) {
    F b[2]; e.as_bases(b);
    for (int i = 0; i < E::DEGREE; i++) b[i].to_limbs(p + i * F::LIMBS);
}
template <class F> void skip(Transcript<F>& tr, size_t n) { for (size_t i = 0; i < n; i++) tr.squeeze_base(); }

template <class F>
int lasso_prove_t(const Preprocessing& pp, int num_vars, const int32_t* rows, size_t n_rows, const uint64_t* inputs, size_t n_inputs,
                  size_t skip_n, uint8_t* proof, size_t cap, size_t* len, uint64_t* r_out, uint64_t* sum_out, size_t* n_squeezed) {
    typedef typename ExtOf<F>::type E;
    LassoNode<F> node(pp, num_vars, std::vector<int>(rows, rows + n_rows));
    std::vector<F> in = load_base<F>(inputs, n_inputs);
    Transcript<F> tr; skip(tr, skip_n);
    std::vector<E> r; E s;
    node.prove_claim_reduction(in, tr, &r, &s);
    *len = tr.stream.size();
    if (n_squeezed) *n_squeezed = tr.n_base_squeezed;
    if (tr.stream.size() > cap) { g_err = "proof buffer too small"; return 2; }
    memcpy(proof, tr.stream.data(), tr.stream.size());
    if (r_out) for (size_t i = 0; i < r.size(); i++) store_ext<F, E>(r[i], r_out + i * E::DEGREE * F::LIMBS);
    if (sum_out) store_ext<F, E>(s, sum_out);
    return 0;
}
template <class F>
int lasso_verify_t(const Preprocessing& pp, int num_vars, const uint8_t* proof, size_t len, size_t skip_n, uint64_t* r_out, uint64_t* sum_out,
                   size_t* consumed) {
    typedef typename ExtOf<F>::type E;
    LassoNode<F> node(pp, num_vars, {});
    Transcript<F> tr(proof, len); skip(tr, skip_n);
    std::vector<E> r; E s;
    node.verify_claim_reduction(tr, &r, &s);
    if (consumed) *consumed = tr.rd_pos;
    if (r_out) for (size_t i = 0; i < r.size(); i++) store_ext<F, E>(r[i], r_out + i * E::DEGREE * F::LIMBS);
    if (sum_out) store_ext<F, E>(s, sum_out);
    return 0;
}
template <class F>
int polynomialize_t(const Preprocessing& pp, int num_vars, const int32_t* rows, size_t n_rows, const uint64_t* inputs, size_t n_inputs,
                    uint64_t* dims, uint64_t* read_cts, uint64_t* final_cts, uint64_t* e_polys) {
    LassoNode<F> node(pp, num_vars, std::vector<int>(rows, rows + n_rows));
    std::vector<F> in = load_base<F>(inputs, n_inputs);
    LassoPolys<F> P = node.polynomialize(in);
    size_t R = P.lookup_outputs.size(), M = (size_t)1 << pp.log2M;
    for (int c = 0; c < pp.C; c++) memcpy(dims + c * R, P.dims[c].data(), R * 8);
    for (int m = 0; m < pp.num_memories; m++) {
        memcpy(read_cts + m * R, P.read_cts[m].data(), R * 8);
        memcpy(final_cts + m * M, P.final_cts[m].data(), M * 8);
        for (size_t j = 0; j < R; j++) P.e_polys[m][j].to_limbs(e_polys + (m * R + j) * F::LIMBS);
    }
    return 0;
}
template <class F>
int sumcheck_t(int arity, size_t nterms, int num_vars, const uint64_t* coeffs, const uint64_t* tables, const uint64_t* claim, size_t skip_n,
               uint8_t* proof, size_t cap, size_t* len, uint64_t* true_evals, uint64_t* r_out, uint64_t* final_evals) {
    typedef typename ExtOf<F>::type E;
    const size_t EL = E::DEGREE * F::LIMBS;
    SumcheckFn<E> g; g.num_vars = num_vars; g.arity = arity; g.coeffs.resize(nterms);
    for (size_t i = 0; i < nterms; i++) g.coeffs[i] = load_ext<F, E>(coeffs + i * EL);
    size_t n = (size_t)1 << num_vars, nt = nterms * arity;
    std::vector<std::vector<F>> polys(nt);
    for (size_t t = 0; t < nt; t++) polys[t] = load_base<F>(tables + t * n * F::LIMBS, n);
    Transcript<F> tr; skip(tr, skip_n);
    E c2; std::vector<E> r, ev; std::vector<std::vector<E>> trace;
    prove_sum_check<F, E, F>(g, load_ext<F, E>(claim), polys, tr, &c2, &r, &ev, &trace);
    *len = tr.stream.size();
    if (tr.stream.size() > cap) { g_err = "proof buffer too small"; return 2; }
    memcpy(proof, tr.stream.data(), tr.stream.size());
    int d = g.degree();
    if (true_evals) for (int rd = 0; rd < num_vars; rd++) for (int x = 0; x <= d; x++) store_ext<F, E>(trace[rd][x], true_evals + (rd * (d + 1) + x) * EL);
    if (r_out) for (int i = 0; i < num_vars; i++) store_ext<F, E>(r[i], r_out + i * EL);
    if (final_evals) for (size_t t = 0; t < nt; t++) store_ext<F, E>(ev[t], final_evals + t * EL);
    return 0;
}
template <class F> int mle_eval_t(const uint64_t* table, int num_vars, const uint64_t* point, uint64_t* out) {
    typedef typename ExtOf<F>::type E;
    const size_t EL = E::DEGREE * F::LIMBS;
    std::vector<F> t = load_base<F>(table, (size_t)1 << num_vars);
    std::vector<E> pt(num_vars);
    for (int i = 0; i < num_vars; i++) pt[i] = load_ext<F, E>(point + i * EL);
    store_ext<F, E>(mle_evaluate<E, F>(t, pt), out);
    return 0;
}
template <class F> int subtable_t(int full, uint64_t bound, int log2M, const uint64_t* point, uint64_t* table_out, uint64_t* mle_out) {
    typedef typename ExtOf<F>::type E;
    const size_t EL = E::DEGREE * F::LIMBS;
    Subtable s{full != 0, bound};
    if (table_out) { auto t = s.materialize<F>(log2M); for (size_t i = 0; i < t.size(); i++) t[i].to_limbs(table_out + i * F::LIMBS); }
    if (mle_out) {
        std::vector<E> pt(log2M);
        for (int i = 0; i < log2M; i++) pt[i] = load_ext<F, E>(point + i * EL);
        store_ext<F, E>(s.evaluate_mle<F, E>(pt, log2M), mle_out);
    }
    return 0;
}
template <class F> int ntt_t(uint64_t* data, int log_n, int inverse, size_t batch) {
    size_t n = (size_t)1 << log_n;
    for (size_t b = 0; b < batch; b++) {
        std::vector<F> a = load_base<F>(data + b * n * F::LIMBS, n);
        ntt_naive_order(a, inverse != 0);
        for (size_t i = 0; i < n; i++) a[i].to_limbs(data + (b * n + i) * F::LIMBS);
    }
    return 0;
}
template <class F>
int bfv_eval_t(int log2_size, int K, const uint64_t* qis, const uint64_t* k0is, const uint64_t* r1b, const uint64_t* r2b, uint64_t sb, uint64_t eb,
               uint64_t k1b, const uint64_t* s, const uint64_t* e, const uint64_t* k1, const uint64_t* ais, const uint64_t* r1is, const uint64_t* r2is,
               uint64_t* lasso_out, size_t* n_lasso, uint64_t* sum_out) {
    BfvParams<F> P; P.log2_size = log2_size; P.K = K; P.s_bound = sb; P.e_bound = eb; P.k1_bound = k1b;
    P.qis = load_base<F>(qis, K); P.k0is = load_base<F>(k0is, K);
    P.r1_bounds.assign(r1b, r1b + K); P.r2_bounds.assign(r2b, r2b + K);
    size_t N2 = (size_t)1 << log2_size;
    std::vector<std::vector<F>> A(K), R1(K);
    for (int i = 0; i < K; i++) { A[i] = load_base<F>(ais + i * N2 * F::LIMBS, N2); R1[i] = load_base<F>(r1is + i * N2 * F::LIMBS, N2); }
    std::vector<F> li, sum;
    bfv_evaluate<F>(P, load_base<F>(s, N2), load_base<F>(e, N2), load_base<F>(k1, N2), A, R1, load_base<F>(r2is, (size_t)K * N2 / 2), &li, &sum);
    *n_lasso = li.size();
    for (size_t i = 0; i < li.size(); i++) li[i].to_limbs(lasso_out + i * F::LIMBS);
    for (size_t i = 0; i < sum.size(); i++) sum[i].to_limbs(sum_out + i * F::LIMBS);
    return 0;
}
template <class F>
BfvParams<F> make_params(int log2_size, int K, const uint64_t* qis, const uint64_t* k0is, const uint64_t* r1b, const uint64_t* r2b, uint64_t sb,
                         uint64_t eb, uint64_t k1b) {
    BfvParams<F> P; P.log2_size = log2_size; P.K = K; P.s_bound = sb; P.e_bound = eb; P.k1_bound = k1b;
    P.qis = load_base<F>(qis, K); P.k0is = load_base<F>(k0is, K);
    P.r1_bounds.assign(r1b, r1b + K); P.r2_bounds.assign(r2b, r2b + K);
    return P;
}
template <class F>
std::vector<std::vector<F>> load_inputs(int log2_size, int K, const uint64_t* s, const uint64_t* e, const uint64_t* k1, const uint64_t* ais,
                                        const uint64_t* r1is, const uint64_t* r2is) {
    size_t N2 = (size_t)1 << log2_size;
    std::vector<std::vector<F>> in;
    in.push_back(load_base<F>(s, N2)); in.push_back(load_base<F>(e, N2)); in.push_back(load_base<F>(k1, N2));
    for (int i = 0; i < K; i++) in.push_back(load_base<F>(ais + i * N2 * F::LIMBS, N2));
    for (int i = 0; i < K; i++) in.push_back(load_base<F>(r1is + i * N2 * F::LIMBS, N2));
    in.push_back(load_base<F>(r2is, (size_t)K * N2 / 2));
    return in;
}
// full BfvEncrypt::prove / ::verify (sk_encryption_circuit.rs:417-517) with the restated engine of gkr.hpp
template <class F>
int bfv_prove_t(int log2_size, int K, const uint64_t* qis, const uint64_t* k0is, const uint64_t* r1b, const uint64_t* r2b, uint64_t sb, uint64_t eb,
                uint64_t k1b, const uint64_t* s, const uint64_t* e, const uint64_t* k1, const uint64_t* ais, const uint64_t* r1is, const uint64_t* r2is,
                const uint64_t* ct0is, uint8_t* proof, size_t cap, size_t* len, int verify_only) {
    BfvEncrypt<F> bfv(make_params<F>(log2_size, K, qis, k0is, r1b, r2b, sb, eb, k1b));
    auto in = load_inputs<F>(log2_size, K, s, e, k1, ais, r1is, r2is);
    auto ct = load_base<F>(ct0is, (size_t)K << log2_size);
    if (verify_only) { bfv.verify(in, ct, proof, *len); return 0; }
    auto pr = bfv.prove(in, ct);
    *len = pr.size();
    if (pr.size() > cap) { g_err = "proof buffer too small"; return 2; }
    memcpy(proof, pr.data(), pr.size());
    return 0;
}
template <class F> struct BfvSession {
    BfvEncrypt<F> bfv;
    std::vector<std::vector<F>> inputs, values;
    std::vector<F> ct0is;
    BfvSession(const BfvParams<F>& P) : bfv(P) {}
};
template <class F> int field_op_t(int op, const uint64_t* a, const uint64_t* b, uint64_t* out) {
    typedef typename ExtOf<F>::type E;
    E x = load_ext<F, E>(a), y = load_ext<F, E>(b), z;
    switch (op) { case 0: z = x + y; break; case 1: z = x - y; break; case 2: z = x * y; break; case 3: z = x.inv(); break; default: return 1; }
    store_ext<F, E>(z, out);
    return 0;
}
}  // namespace

#define GUARD(...) try { __VA_ARGS__ } catch (const std::exception& e) { g_err = e.what(); return 1; }

extern "C" {
const char* hgo_last_error() { return g_err.c_str(); }
int hgo_num_threads() { return omp_get_max_threads(); }
void hgo_set_num_threads(int n) { omp_set_num_threads(n); }
void hgo_set_assumption(int which, int value) {
    if (which == 3) assumptions().a3_wire = value;
    if (which == 31) assumptions().a3_h1 = value;
    if (which == 5) assumptions().a5_ascending = value;
}
void hgo_keccak256(const uint8_t* in, size_t n, uint8_t* out) { keccak256(in, n, out); }
// interchange dump: record every base-field squeeze / write of the transcripts used from now on (protocol.hpp EventLog)
void hgo_events_begin() { event_log().bytes.clear(); event_log().on = true; }
size_t hgo_events_end(uint8_t* out, size_t cap) {  // stops recording; returns the number of bytes (copied when they fit)
    event_log().on = false;
    const size_t n = event_log().bytes.size();
    if (out && n <= cap) memcpy(out, event_log().bytes.data(), n);
    return n;
}
// base-field challenge chain c_i (transcript.rs:199-203)
void hgo_challenges(int field, size_t n, uint64_t* out) {
    if (field == 0) { Transcript<Gl> t; for (size_t i = 0; i < n; i++) t.squeeze_base().to_limbs(out + i); }
    else { Transcript<Fr> t; for (size_t i = 0; i < n; i++) t.squeeze_base().to_limbs(out + 4 * i); }
}
int hgo_field_op(int field, int op, const uint64_t* a, const uint64_t* b, uint64_t* out) {
    return field == 0 ? field_op_t<Gl>(op, a, b, out) : field_op_t<Fr>(op, a, b, out);
}

void* hgo_pp_new(const uint64_t* bounds, size_t n, int C, int log2M) {
    return new Preprocessing(Preprocessing::preprocess(std::vector<uint64_t>(bounds, bounds + n), C, log2M));
}
void hgo_pp_free(void* pp) { delete (Preprocessing*)pp; }
// info[0..3) = num_lookups, num_subtables, num_memories
void hgo_pp_info(void* h, int* info) {
    auto* pp = (Preprocessing*)h;
    info[0] = (int)pp->lookups.size(); info[1] = (int)pp->subtables.size(); info[2] = pp->num_memories;
}
void hgo_pp_maps(void* h, uint64_t* lookup_bounds, int* subtable_full, uint64_t* subtable_bound, int* mem_to_sub, int* mem_to_dim) {
    auto* pp = (Preprocessing*)h;
    for (size_t i = 0; i < pp->lookups.size(); i++) lookup_bounds[i] = pp->lookups[i].bound;
    for (size_t i = 0; i < pp->subtables.size(); i++) { subtable_full[i] = pp->subtables[i].full; subtable_bound[i] = pp->subtables[i].bound; }
    for (int i = 0; i < pp->num_memories; i++) { mem_to_sub[i] = pp->memory_to_subtable_index[i]; mem_to_dim[i] = pp->memory_to_dimension_index[i]; }
}
int hgo_pp_lookup_memories(void* h, int lookup, int* out) {
    auto* pp = (Preprocessing*)h;
    auto& v = pp->lookup_to_memory_indices[lookup];
    for (size_t i = 0; i < v.size(); i++) out[i] = v[i];
    return (int)v.size();
}
int hgo_pp_chunk_bits(void* h, int lookup, int* out) {
    auto* pp = (Preprocessing*)h;
    auto v = pp->lookups[lookup].chunk_bits(pp->log2M);
    for (size_t i = 0; i < v.size(); i++) out[i] = v[i];
    return (int)v.size();
}

int hgo_lasso_prove(int field, void* pp, int num_vars, const int32_t* rows, size_t n_rows, const uint64_t* inputs, size_t n_inputs, size_t skip_n,
                    uint8_t* proof, size_t cap, size_t* len, uint64_t* r_out, uint64_t* sum_out, size_t* n_squeezed) {
    GUARD(return field == 0 ? lasso_prove_t<Gl>(*(Preprocessing*)pp, num_vars, rows, n_rows, inputs, n_inputs, skip_n, proof, cap, len, r_out, sum_out, n_squeezed)
                            : lasso_prove_t<Fr>(*(Preprocessing*)pp, num_vars, rows, n_rows, inputs, n_inputs, skip_n, proof, cap, len, r_out, sum_out, n_squeezed);)
}
int hgo_lasso_verify(int field, void* pp, int num_vars, const uint8_t* proof, size_t len, size_t skip_n, uint64_t* r_out, uint64_t* sum_out,
                     size_t* consumed) {
    GUARD(return field == 0 ? lasso_verify_t<Gl>(*(Preprocessing*)pp, num_vars, proof, len, skip_n, r_out, sum_out, consumed)
                            : lasso_verify_t<Fr>(*(Preprocessing*)pp, num_vars, proof, len, skip_n, r_out, sum_out, consumed);)
}
int hgo_lasso_polynomialize(int field, void* pp, int num_vars, const int32_t* rows, size_t n_rows, const uint64_t* inputs, size_t n_inputs,
                            uint64_t* dims, uint64_t* read_cts, uint64_t* final_cts, uint64_t* e_polys) {
    GUARD(return field == 0 ? polynomialize_t<Gl>(*(Preprocessing*)pp, num_vars, rows, n_rows, inputs, n_inputs, dims, read_cts, final_cts, e_polys)
                            : polynomialize_t<Fr>(*(Preprocessing*)pp, num_vars, rows, n_rows, inputs, n_inputs, dims, read_cts, final_cts, e_polys);)
}
// generic sumcheck over g = poly(0) * sum_i coeffs[i] * prod_{k<arity} poly(arity*i+k); tables = (nterms*arity) x 2^num_vars base elements
int hgo_sumcheck_prove(int field, int arity, size_t nterms, int num_vars, const uint64_t* coeffs, const uint64_t* tables, const uint64_t* claim,
                       size_t skip_n, uint8_t* proof, size_t cap, size_t* len, uint64_t* true_evals, uint64_t* r_out, uint64_t* final_evals) {
    GUARD(return field == 0 ? sumcheck_t<Gl>(arity, nterms, num_vars, coeffs, tables, claim, skip_n, proof, cap, len, true_evals, r_out, final_evals)
                            : sumcheck_t<Fr>(arity, nterms, num_vars, coeffs, tables, claim, skip_n, proof, cap, len, true_evals, r_out, final_evals);)
}
int hgo_mle_eval(int field, const uint64_t* table, int num_vars, const uint64_t* point, uint64_t* out) {
    GUARD(return field == 0 ? mle_eval_t<Gl>(table, num_vars, point, out) : mle_eval_t<Fr>(table, num_vars, point, out);)
}
int hgo_ntt(int field, uint64_t* data, int log_n, int inverse, size_t batch) {
    GUARD(return field == 0 ? ntt_t<Gl>(data, log_n, inverse, batch) : ntt_t<Fr>(data, log_n, inverse, batch);)
}
// forward evaluation of the BFV circuit: lasso_inputs_batched output and `sum` output (sk_encryption_circuit.rs:163-181,280-285)
int hgo_bfv_eval(int field, int log2_size, int K, const uint64_t* qis, const uint64_t* k0is, const uint64_t* r1b, const uint64_t* r2b, uint64_t sb,
                 uint64_t eb, uint64_t k1b, const uint64_t* s, const uint64_t* e, const uint64_t* k1, const uint64_t* ais, const uint64_t* r1is,
                 const uint64_t* r2is, uint64_t* lasso_out, size_t* n_lasso, uint64_t* sum_out) {
    GUARD(return field == 0 ? bfv_eval_t<Gl>(log2_size, K, qis, k0is, r1b, r2b, sb, eb, k1b, s, e, k1, ais, r1is, r2is, lasso_out, n_lasso, sum_out)
                            : bfv_eval_t<Fr>(log2_size, K, qis, k0is, r1b, r2b, sb, eb, k1b, s, e, k1, ais, r1is, r2is, lasso_out, n_lasso, sum_out);)
}
int hgo_bfv_prove(int field, int log2_size, int K, const uint64_t* qis, const uint64_t* k0is, const uint64_t* r1b, const uint64_t* r2b, uint64_t sb,
                  uint64_t eb, uint64_t k1b, const uint64_t* s, const uint64_t* e, const uint64_t* k1, const uint64_t* ais, const uint64_t* r1is,
                  const uint64_t* r2is, const uint64_t* ct0is, uint8_t* proof, size_t cap, size_t* len, int verify_only) {
    GUARD(return field == 0 ? bfv_prove_t<Gl>(log2_size, K, qis, k0is, r1b, r2b, sb, eb, k1b, s, e, k1, ais, r1is, r2is, ct0is, proof, cap, len, verify_only)
                            : bfv_prove_t<Fr>(log2_size, K, qis, k0is, r1b, r2b, sb, eb, k1b, s, e, k1, ais, r1is, r2is, ct0is, proof, cap, len, verify_only);)
}
// session API for timing the reference's `GKR prove` span alone: circuit built and evaluated once (witness gen), then
// hgo_bfv_session_prove = squeeze output point + output MLE + prove_gkr (sk_encryption_circuit.rs:444-457)
void* hgo_bfv_session_new(int field, int log2_size, int K, const uint64_t* qis, const uint64_t* k0is, const uint64_t* r1b, const uint64_t* r2b, uint64_t sb,
                          uint64_t eb, uint64_t k1b, const uint64_t* s, const uint64_t* e, const uint64_t* k1, const uint64_t* ais, const uint64_t* r1is,
                          const uint64_t* r2is, const uint64_t* ct0is) {
    try {
        if (field != 0) { g_err = "session API: Goldilocks only"; return nullptr; }
        auto* S = new BfvSession<Gl>(make_params<Gl>(log2_size, K, qis, k0is, r1b, r2b, sb, eb, k1b));
        S->inputs = load_inputs<Gl>(log2_size, K, s, e, k1, ais, r1is, r2is);
        S->ct0is = load_base<Gl>(ct0is, (size_t)K << log2_size);
        S->values = S->bfv.circuit.evaluate(S->inputs);
        return S;
    } catch (const std::exception& ex) { g_err = ex.what(); return nullptr; }
}
void hgo_bfv_session_free(void* h) { delete (BfvSession<Gl>*)h; }
int hgo_bfv_session_prove(void* h, uint8_t* proof, size_t cap, size_t* len) {
    GUARD(
        auto* S = (BfvSession<Gl>*)h;
        Transcript<Gl> tr;
        std::vector<Gl2> point = tr.squeeze_n(S->bfv.ct0is_log2_size());
        Gl2 value = mle_evaluate<Gl2, Gl>(S->ct0is, point);
        std::vector<EvalClaim<Gl>> oc = {{{}, Gl2::zero()}, {point, value}};
        S->bfv.circuit.prove_gkr(S->values, oc, tr);
        *len = tr.stream.size();
        if (tr.stream.size() > cap) { g_err = "proof buffer too small"; return 2; }
        memcpy(proof, tr.stream.data(), tr.stream.size());
        return 0;)
}
int hgo_subtable(int field, int full, uint64_t bound, int log2M, const uint64_t* point, uint64_t* table_out, uint64_t* mle_out) {
    GUARD(return field == 0 ? subtable_t<Gl>(full, bound, log2M, point, table_out, mle_out) : subtable_t<Fr>(full, bound, log2M, point, table_out, mle_out);)
}
}
