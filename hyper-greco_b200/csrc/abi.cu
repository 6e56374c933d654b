// extern "C" boundary (include/hg_b200.h). Translates handles + limb arrays to the templated C++/CUDA implementation
// and turns every exception into an error code + message.
#include "../../include/hg_b200.h"

#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include <map>

#include "bfv.cuh"
#include "gkr_dev.cuh"
#include "prover.cuh"

using namespace hg;

namespace {
thread_local std::string g_err;
int fail(const std::string& m) { g_err = m; return 1; }
#define HG_TRY(...)                                                             \
    try { __VA_ARGS__; return 0; }                                              \
    catch (const std::exception& e) { cudaGetLastError(); return fail(e.what()); } \
    catch (...) { cudaGetLastError(); return fail("unknown error"); }

template <class FP> __global__ void k_selftest(int op, const typename FP::X* a, const typename FP::X* b, size_t n, typename FP::X* out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    typename FP::X x = a[i], y = b[i], z;
    if (op == 0) z = FP::x_add(x, y);
    else if (op == 1) z = FP::x_sub(x, y);
    else z = FP::x_mul(x, y);
    out[i] = z;
}
}  // namespace

struct hg_ctx {
    DeviceCtx dev;
    int field_id;
    WireOptions wire;
    std::unique_ptr<NttEngine<GlField>> ntt;
};
struct hg_circuit {
    hg_ctx* ctx;
    std::unique_ptr<GkrCircuitDev<GlField>> gl;
    std::vector<std::vector<GkrCircuitDev<GlField>::InputClaim>> input_claims;
};

namespace {
void ntt_run(hg_ctx* ctx, u64* d_data, int log_n, bool inverse, size_t batch) {
    if (!ctx->ntt) ctx->ntt.reset(new NttEngine<GlField>(&ctx->dev));
    ctx->ntt->run(d_data, log_n, inverse, batch);
}
}  // namespace
struct hg_buf {
    void* p = nullptr;
    size_t bytes = 0;
    int device = 0;
};
struct hg_transcript {
    int field_id;
    std::unique_ptr<Keccak256Transcript<GlField>> gl;
};
struct hg_lasso_pp {
    LassoPreprocessing pp;
    std::vector<uint64_t> lookup_bounds;  // in preprocessing order
};
struct hg_lasso_node {
    hg_ctx* ctx;
    std::unique_ptr<LassoNodeDev<GlField>> gl;
    DevBuf<u64> staging;  // device copy of host inputs
    size_t log2_input_size = 0;
};

static void need_gl(int field_id) {
    if (field_id != HG_FIELD_GOLDILOCKS) throw std::runtime_error("field not supported by this build (only HG_FIELD_GOLDILOCKS)");
}

extern "C" {

const char* hg_last_error(void) { return g_err.c_str(); }
int hg_version(void) { return 1; }

int hg_ctx_create(int device, int field_id, hg_ctx** out) {
    HG_TRY({
        need_gl(field_id);
        int n = 0;
        HG_CUDA(cudaGetDeviceCount(&n));
        if (device < 0 || device >= n) throw std::runtime_error("no such CUDA device");
        HG_CUDA(cudaSetDevice(device));
        std::unique_ptr<hg_ctx> c(new hg_ctx());
        c->dev.device = device;
        c->field_id = field_id;
        HG_CUDA(cudaStreamCreateWithFlags(&c->dev.stream, cudaStreamNonBlocking));
        HG_CUDA(cudaDeviceGetAttribute(&c->dev.sm_count, cudaDevAttrMultiProcessorCount, device));
        *out = c.release();
    })
}
void hg_ctx_destroy(hg_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->dev.device);
    if (ctx->dev.stream) cudaStreamDestroy(ctx->dev.stream);
    delete ctx;
}
int hg_ctx_set_option(hg_ctx* ctx, int option, int value) {
    HG_TRY({
        if (option == HG_OPT_A3_WIRE) ctx->wire.a3_wire = value;
        else if (option == HG_OPT_A3_H1) ctx->wire.a3_h1 = value;
        else if (option == HG_OPT_A5_ASCENDING) ctx->wire.a5_ascending = value;
        else throw std::runtime_error("unknown option");
    })
}
int hg_ctx_synchronize(hg_ctx* ctx) { HG_TRY({ HG_CUDA(cudaSetDevice(ctx->dev.device)); HG_CUDA(cudaStreamSynchronize(ctx->dev.stream)); }) }
uint64_t hg_ctx_launch_count(hg_ctx* ctx) { return ctx->dev.launches; }
int hg_ctx_profile(hg_ctx* ctx, int enable) {
    HG_TRY({
        HG_CUDA(cudaSetDevice(ctx->dev.device));
        ctx->dev.profile_collect();
        if (enable) ctx->dev.profile_reset();
        ctx->dev.profile = enable != 0;
    })
}
int hg_ctx_profile_read(hg_ctx* ctx, int kernel_class, uint64_t* launches, double* ms, uint64_t* algorithmic_bytes) {
    HG_TRY({
        if (kernel_class < 0 || kernel_class >= KC_COUNT) throw std::runtime_error("no such kernel class");
        HG_CUDA(cudaSetDevice(ctx->dev.device));
        ctx->dev.profile_collect();
        *launches = ctx->dev.cls_launches[kernel_class];
        *ms = ctx->dev.cls_ms[kernel_class];
        *algorithmic_bytes = ctx->dev.cls_bytes[kernel_class];
    })
}
int hg_kernel_class_count(void) { return KC_COUNT; }
const char* hg_kernel_class_name(int kernel_class) { return kernel_class_name(kernel_class); }
void* hg_ctx_stream(hg_ctx* ctx) { return (void*)ctx->dev.stream; }

int hg_buf_alloc(hg_ctx* ctx, size_t bytes, hg_buf** out) {
    HG_TRY({
        HG_CUDA(cudaSetDevice(ctx->dev.device));
        std::unique_ptr<hg_buf> b(new hg_buf());
        HG_CUDA(cudaMalloc(&b->p, bytes ? bytes : 1));
        b->bytes = bytes;
        b->device = ctx->dev.device;
        *out = b.release();
    })
}
int hg_buf_upload(hg_ctx* ctx, hg_buf* buf, size_t offset, const void* host, size_t bytes) {
    HG_TRY({
        if (offset + bytes > buf->bytes) throw std::runtime_error("hg_buf_upload: out of range");
        HG_CUDA(cudaSetDevice(ctx->dev.device));
        HG_CUDA(cudaMemcpyAsync((char*)buf->p + offset, host, bytes, cudaMemcpyHostToDevice, ctx->dev.stream));
        HG_CUDA(cudaStreamSynchronize(ctx->dev.stream));
    })
}
int hg_buf_download(hg_ctx* ctx, const hg_buf* buf, size_t offset, void* host, size_t bytes) {
    HG_TRY({
        if (offset + bytes > buf->bytes) throw std::runtime_error("hg_buf_download: out of range");
        HG_CUDA(cudaSetDevice(ctx->dev.device));
        HG_CUDA(cudaMemcpyAsync(host, (const char*)buf->p + offset, bytes, cudaMemcpyDeviceToHost, ctx->dev.stream));
        HG_CUDA(cudaStreamSynchronize(ctx->dev.stream));
    })
}
void* hg_buf_device_ptr(hg_buf* buf) { return buf->p; }
size_t hg_buf_size(const hg_buf* buf) { return buf->bytes; }
void hg_buf_free(hg_buf* buf) {
    if (!buf) return;
    cudaSetDevice(buf->device);
    cudaFree(buf->p);
    delete buf;
}

// ---- transcript
int hg_transcript_new(int field_id, hg_transcript** out) {
    HG_TRY({
        need_gl(field_id);
        std::unique_ptr<hg_transcript> t(new hg_transcript());
        t->field_id = field_id;
        t->gl.reset(new Keccak256Transcript<GlField>());
        *out = t.release();
    })
}
int hg_transcript_from_proof(int field_id, const uint8_t* proof, size_t len, hg_transcript** out) {
    HG_TRY({
        need_gl(field_id);
        std::unique_ptr<hg_transcript> t(new hg_transcript());
        t->field_id = field_id;
        t->gl.reset(new Keccak256Transcript<GlField>(proof, len));
        *out = t.release();
    })
}
void hg_transcript_free(hg_transcript* t) { delete t; }
int hg_transcript_squeeze_challenge(hg_transcript* t, uint64_t* out_ext) { HG_TRY({ GlField::x_to_limbs(t->gl->squeeze_challenge(), out_ext); }) }
int hg_transcript_write_felt_ext(hg_transcript* t, const uint64_t* ext) { HG_TRY({ t->gl->write_felt_ext(GlField::x_from_limbs(ext)); }) }
int hg_transcript_read_felt_ext(hg_transcript* t, uint64_t* out_ext) { HG_TRY({ GlField::x_to_limbs(t->gl->read_felt_ext(), out_ext); }) }
size_t hg_transcript_proof_len(const hg_transcript* t) { return t->gl->proof().size(); }
int hg_transcript_proof_copy(const hg_transcript* t, uint8_t* out, size_t cap) {
    HG_TRY({
        auto& p = t->gl->proof();
        if (p.size() > cap) throw std::runtime_error("hg_transcript_proof_copy: buffer too small");
        memcpy(out, p.data(), p.size());
    })
}
size_t hg_transcript_num_squeezed(const hg_transcript* t) { return t->gl->num_base_squeezed(); }

// ---- preprocessing
int hg_lasso_preprocess(const uint64_t* bounds, size_t n_bounds, size_t C, size_t M, hg_lasso_pp** out) {
    HG_TRY({
        if (M < 2 || (M & (M - 1))) throw std::runtime_error("M must be a power of two >= 2");
        std::vector<std::shared_ptr<LookupType>> lk;
        for (size_t i = 0; i < n_bounds; i++) lk.push_back(std::make_shared<RangeLookup>(bounds[i]));
        std::unique_ptr<hg_lasso_pp> h(new hg_lasso_pp());
        h->pp = LassoPreprocessing::preprocess(lk, C, M);
        for (auto& l : h->pp.lookups) h->lookup_bounds.push_back(static_cast<RangeLookup*>(l.get())->bound());
        *out = h.release();
    })
}
void hg_lasso_pp_free(hg_lasso_pp* pp) { delete pp; }
size_t hg_lasso_pp_num_lookups(const hg_lasso_pp* pp) { return pp->pp.lookups.size(); }
size_t hg_lasso_pp_num_subtables(const hg_lasso_pp* pp) { return pp->pp.subtables_by_idx.size(); }
size_t hg_lasso_pp_num_memories(const hg_lasso_pp* pp) { return pp->pp.num_memories; }
int hg_lasso_pp_lookup_index(const hg_lasso_pp* pp, uint64_t bound) {
    auto it = pp->pp.lookup_id_to_index.find(RangeLookup::id_for(bound));
    return it == pp->pp.lookup_id_to_index.end() ? -1 : (int)it->second;
}
int hg_lasso_pp_memory_maps(const hg_lasso_pp* pp, uint32_t* mem_to_subtable, uint32_t* mem_to_dimension) {
    HG_TRY({
        for (size_t i = 0; i < pp->pp.num_memories; i++) {
            if (mem_to_subtable) mem_to_subtable[i] = (uint32_t)pp->pp.memory_to_subtable_index[i];
            if (mem_to_dimension) mem_to_dimension[i] = (uint32_t)pp->pp.memory_to_dimension_index[i];
        }
    })
}
int hg_lasso_pp_subtable_id(const hg_lasso_pp* pp, size_t idx, char* out, size_t cap) {
    HG_TRY({
        if (idx >= pp->pp.subtables_by_idx.size()) throw std::runtime_error("subtable index out of range");
        std::string id = pp->pp.subtables_by_idx[idx]->subtable_id();
        if (id.size() + 1 > cap) throw std::runtime_error("buffer too small");
        memcpy(out, id.c_str(), id.size() + 1);
    })
}

// ---- node
int hg_lasso_node_new(hg_ctx* ctx, const hg_lasso_pp* pp, size_t num_vars, const uint64_t* seg_bounds, const uint64_t* seg_lens, size_t n_segs,
                      hg_lasso_node** out) {
    HG_TRY({
        HG_CUDA(cudaSetDevice(ctx->dev.device));
        if (num_vars < 1 || num_vars > 30) throw std::runtime_error("num_vars out of range");
        std::vector<uint8_t> rows;
        for (size_t s = 0; s < n_segs; s++) {
            int li = hg_lasso_pp_lookup_index(pp, seg_bounds[s]);
            if (li < 0) throw std::runtime_error("lookup id " + RangeLookup::id_for(seg_bounds[s]) + " not in preprocessing");
            rows.insert(rows.end(), seg_lens[s], (uint8_t)li);
        }
        std::unique_ptr<hg_lasso_node> n(new hg_lasso_node());
        n->ctx = ctx;
        n->gl.reset(new LassoNodeDev<GlField>(&ctx->dev, pp->pp, (int)num_vars, rows));
        n->log2_input_size = std::max<size_t>(num_vars, ilog2u(pp->pp.M));  // lasso.rs:45-47
        *out = n.release();
    })
}
void hg_lasso_node_free(hg_lasso_node* node) {
    if (!node) return;
    cudaSetDevice(node->ctx->dev.device);
    delete node;
}
size_t hg_lasso_node_log2_input_size(const hg_lasso_node* node) { return node->log2_input_size; }
size_t hg_lasso_node_device_bytes(const hg_lasso_node* node) { return node->gl->device_bytes(); }
size_t hg_lasso_node_num_chunks(const hg_lasso_node* node) { return node->gl->chunk_dims().size(); }
void hg_lasso_node_timing(const hg_lasso_node* node, double* out_us4) { for (int i = 0; i < 4; i++) out_us4[i] = node->gl->timing()[i]; }

int hg_lasso_node_prove(hg_lasso_node* node, const void* inputs, size_t n_inputs, int inputs_on_device, hg_transcript* t, int mode,
                        uint64_t* out_point, uint64_t* out_value) {
    HG_TRY({
        hg_ctx* ctx = node->ctx;
        HG_CUDA(cudaSetDevice(ctx->dev.device));
        const u64* d_in = (const u64*)inputs;
        if (!inputs_on_device) {
            if (node->staging.n < n_inputs) node->staging.alloc(n_inputs);
            HG_CUDA(cudaMemcpyAsync(node->staging.p, inputs, n_inputs * sizeof(u64), cudaMemcpyHostToDevice, ctx->dev.stream));
            d_in = node->staging.p;
        }
        std::vector<gl2> pt;
        gl2 val;
        node->gl->prove(d_in, n_inputs, *t->gl, mode == HG_MODE_INTERACTIVE ? kModeInteractive : kModePrefetch, ctx->wire, &pt, &val);
        if (out_point) for (size_t i = 0; i < pt.size(); i++) GlField::x_to_limbs(pt[i], out_point + 2 * i);
        if (out_value) GlField::x_to_limbs(val, out_value);
    })
}
int hg_lasso_node_download_polys(hg_lasso_node* node, uint16_t* dims, uint32_t* read_cts, uint32_t* final_cts, uint64_t* e_polys) {
    HG_TRY({
        HG_CUDA(cudaSetDevice(node->ctx->dev.device));
        std::vector<u16> d; std::vector<u32> r, f; std::vector<u64> e;
        node->gl->download_polys(dims ? &d : nullptr, read_cts ? &r : nullptr, final_cts ? &f : nullptr, e_polys ? &e : nullptr);
        if (dims) memcpy(dims, d.data(), d.size() * sizeof(u16));
        if (read_cts) memcpy(read_cts, r.data(), r.size() * sizeof(u32));
        if (final_cts) memcpy(final_cts, f.data(), f.size() * sizeof(u32));
        if (e_polys) memcpy(e_polys, e.data(), e.size() * sizeof(u64));
    })
}

// ---- generic sumcheck
int hg_sumcheck_prove(hg_ctx* ctx, int arity, size_t n_terms, size_t num_vars, const uint64_t* coeffs_ext, const void* d_tables,
                      const uint64_t* claim_ext, hg_transcript* t, int mode, uint64_t* out_point, uint64_t* out_evals) {
    HG_TRY({
        typedef GlField FP;
        HG_CUDA(cudaSetDevice(ctx->dev.device));
        if (arity != 1 && arity != 2) throw std::runtime_error("arity must be 1 or 2");
        if (num_vars < 1 || num_vars > 30) throw std::runtime_error("num_vars out of range");
        const size_t n = (size_t)1 << num_vars, ntab = n_terms * arity;
        DevBuf<gl2> coeffs, bufA, bufB, partials;
        DevBuf<unsigned> counters;
        coeffs.alloc(n_terms);
        std::vector<gl2> hc(n_terms);
        for (size_t i = 0; i < n_terms; i++) hc[i] = FP::x_from_limbs(coeffs_ext + 2 * i);
        HG_CUDA(cudaMemcpy(coeffs.p, hc.data(), n_terms * sizeof(gl2), cudaMemcpyHostToDevice));
        bufA.alloc(std::max<size_t>(ntab * (n / 2), ntab));
        bufB.alloc(std::max<size_t>(ntab * (n / 4), ntab));
        ScScratch sc;
        sc.max_blocks = ctx->dev.sm_count * 8;
        partials.alloc((size_t)sc.max_blocks * 4);
        counters.alloc(4);
        HG_CUDA(cudaMemset(counters.p, 0, counters.bytes()));
        sc.partials = partials.p; sc.counters = counters.p;
        Channel<FP> ch(&ctx->dev, num_vars + 1, 4 * num_vars + ntab + 4);
        ch.begin(t->gl.get(), mode == HG_MODE_INTERACTIVE ? kModeInteractive : kModePrefetch, num_vars);
        auto st = std::make_shared<ScHostState<FP>>();
        st->claim = FP::x_from_limbs(claim_ext);
        size_t first = 0, eo = 0;
        if (arity == 1) sumcheck_dev<FP, 1>(&ctx->dev, KC_SC_COLL, ch, ctx->wire, (const u64*)d_tables, n, (int)n_terms, coeffs.p, bufA.p, bufB.p, sc, st, &first, &eo);
        else sumcheck_dev<FP, 2>(&ctx->dev, KC_SC_GP, ch, ctx->wire, (const u64*)d_tables, n, (int)n_terms, coeffs.p, bufA.p, bufB.p, sc, st, &first, &eo);
        ch.flush();
        if (out_point) for (size_t i = 0; i < num_vars; i++) FP::x_to_limbs(ch.chal(first + i), out_point + 2 * i);
        if (out_evals) for (size_t i = 0; i < ntab; i++) FP::x_to_limbs(ch.msg(eo + i), out_evals + 2 * i);
    })
}

int hg_mle_eval_batch(hg_ctx* ctx, const void* d_tables, size_t n_tables, size_t stride, size_t num_vars, const uint64_t* point_ext,
                      uint64_t* out_ext) {
    HG_TRY({
        typedef GlField FP;
        HG_CUDA(cudaSetDevice(ctx->dev.device));
        const size_t n = (size_t)1 << num_vars;
        cudaStream_t s = ctx->dev.stream;
        DevBuf<gl2> pt, eq, partials, out;
        DevBuf<unsigned> counters;
        std::vector<gl2> hp(num_vars);
        for (size_t i = 0; i < num_vars; i++) hp[i] = FP::x_from_limbs(point_ext + 2 * i);
        pt.alloc(std::max<size_t>(num_vars, 1));
        HG_CUDA(cudaMemcpy(pt.p, hp.data(), num_vars * sizeof(gl2), cudaMemcpyHostToDevice));
        const int lo = num_vars < 12 ? (int)num_vars : 12;
        const size_t nlo = (size_t)1 << lo, nhi = n >> lo;
        eq.alloc(nlo + nhi);
        out.alloc(n_tables);
        int blocks = (int)std::min<size_t>(nhi, (size_t)ctx->dev.sm_count * 2);
        partials.alloc((size_t)blocks * n_tables);
        counters.alloc(n_tables);
        HG_CUDA(cudaMemsetAsync(counters.p, 0, counters.bytes(), s));
        HG_K(&ctx->dev, KC_EQ, (nlo + nhi) * sizeof(gl2), k_eq_split<FP><<<(unsigned)((nlo + nhi + HG_BLOCK - 1) / HG_BLOCK), HG_BLOCK, 0, s>>>(pt.p, (int)num_vars, lo, eq.p, eq.p + nlo));
        HG_K(&ctx->dev, KC_DOT, n_tables * n * 8,
             k_dot_eq<FP, u64><<<dim3(blocks, (unsigned)n_tables), HG_BLOCK, 0, s>>>((const u64*)d_tables, stride, n, lo, eq.p, eq.p + nlo, partials.p, counters.p, out.p));
        std::vector<gl2> ho(n_tables);
        HG_CUDA(cudaMemcpyAsync(ho.data(), out.p, n_tables * sizeof(gl2), cudaMemcpyDeviceToHost, s));
        HG_CUDA(cudaStreamSynchronize(s));
        for (size_t i = 0; i < n_tables; i++) FP::x_to_limbs(ho[i], out_ext + 2 * i);
    })
}

int hg_ntt(hg_ctx* ctx, void* d_data, size_t log_n, int inverse, size_t batch) {
    HG_TRY({
        HG_CUDA(cudaSetDevice(ctx->dev.device));
        ntt_run(ctx, (u64*)d_data, (int)log_n, inverse != 0, batch);
    })
}

int hg_bfv_evaluate(hg_ctx* ctx, size_t log2_size, size_t K, const uint64_t* qis, const uint64_t* k0is, const uint64_t* r1_bounds,
                    const uint64_t* r2_bounds, uint64_t s_bound, uint64_t e_bound, uint64_t k1_bound, const void* d_s, const void* d_e,
                    const void* d_k1, const void* d_ais, const void* d_r1is, const void* d_r2is, void* d_lasso_inputs, void* d_sum) {
    HG_TRY({
        typedef GlField FP;
        HG_CUDA(cudaSetDevice(ctx->dev.device));
        cudaStream_t s = ctx->dev.stream;
        const size_t N2 = (size_t)1 << log2_size, r2_len = K * (N2 / 2);
        BfvShape sh;
        sh.log2_size = (int)log2_size; sh.K = (int)K; sh.n_chunks = (int)std::max<size_t>(1, (r2_len + N2 - 1) / N2);
        // per-modulus constants as field elements (F::from_str_vartime(QIS/K0IS), sk_encryption_circuit.rs:109,135)
        std::vector<u64> consts(3 * K);
        for (size_t i = 0; i < K; i++) { consts[i] = gl_from_u64(qis[i]); consts[K + i] = gl_from_u64(k0is[i]); consts[2 * K + i] = gl_from_u64(r1_bounds[i]); }
        DevBuf<u64> d_consts, d_sai, d_seval;
        d_consts.alloc(3 * K);
        HG_CUDA(cudaMemcpyAsync(d_consts.p, consts.data(), consts.size() * 8, cudaMemcpyHostToDevice, s));
        const size_t total = (K + sh.n_chunks + 3) * N2;
        HG_K(&ctx->dev, KC_MISC, total * 16,
             k_bfv_lasso_inputs<FP><<<(unsigned)((total + 255) / 256), 256, 0, s>>>(sh, (const u64*)d_s, (const u64*)d_e, (const u64*)d_k1, (const u64*)d_r1is,
                                                                                  (const u64*)d_r2is, r2_len, d_consts.p + 2 * K, gl_from_u64(r2_bounds[0]),
                                                                                  gl_from_u64(s_bound), gl_from_u64(e_bound), gl_from_u64(k1_bound), (u64*)d_lasso_inputs));
        // s_eval = FFT(s); ai_eval = FFT(ai); sai = IFFT(s_eval . ai_eval)   (sk_encryption_circuit.rs:224-260)
        d_seval.alloc(N2);
        d_sai.alloc(K * N2);
        HG_CUDA(cudaMemcpyAsync(d_seval.p, d_s, N2 * 8, cudaMemcpyDeviceToDevice, s));
        HG_CUDA(cudaMemcpyAsync(d_sai.p, d_ais, K * N2 * 8, cudaMemcpyDeviceToDevice, s));
        ntt_run(ctx, d_seval.p, (int)log2_size, false, 1);
        ntt_run(ctx, d_sai.p, (int)log2_size, false, K);
        HG_K(&ctx->dev, KC_MISC, 3 * K * N2 * 8, k_pointwise_mul_bcast<FP><<<dim3((unsigned)((N2 + 255) / 256), (unsigned)K), 256, 0, s>>>(d_sai.p, d_seval.p, N2));
        ntt_run(ctx, d_sai.p, (int)log2_size, true, K);
        HG_K(&ctx->dev, KC_MISC, 4 * K * N2 * 8,
             k_bfv_sum<FP><<<dim3((unsigned)((N2 + 255) / 256), (unsigned)K), 256, 0, s>>>(sh, d_sai.p, (const u64*)d_e, (const u64*)d_k1, (const u64*)d_r1is,
                                                                                         (const u64*)d_r2is, d_consts.p, d_consts.p + K, (u64*)d_sum));
        HG_CUDA(cudaStreamSynchronize(s));  // temporaries are freed on return
    })
}

// ---- circuit-level API: Circuit::{insert, connect, evaluate} + prove_gkr
int hg_circuit_new(hg_ctx* ctx, hg_circuit** out) {
    HG_TRY({
        HG_CUDA(cudaSetDevice(ctx->dev.device));
        if (!ctx->ntt) ctx->ntt.reset(new NttEngine<GlField>(&ctx->dev));
        std::unique_ptr<hg_circuit> c(new hg_circuit());
        c->ctx = ctx;
        c->gl.reset(new GkrCircuitDev<GlField>(&ctx->dev, ctx->ntt.get()));
        *out = c.release();
    })
}
void hg_circuit_free(hg_circuit* c) {
    if (!c) return;
    cudaSetDevice(c->ctx->dev.device);
    delete c;
}
int hg_circuit_insert_input(hg_circuit* c, size_t log2_size, size_t num_reps, int* out_id) { HG_TRY({ *out_id = c->gl->insert_input(log2_size, num_reps); }) }
int hg_circuit_insert_fft(hg_circuit* c, size_t log2_size, int inverse, int* out_id) {
    HG_TRY({ HG_CUDA(cudaSetDevice(c->ctx->dev.device)); *out_id = c->gl->insert_fft(log2_size, inverse != 0); })
}
int hg_circuit_insert_lasso(hg_circuit* c, hg_lasso_node* node, int* out_id) { HG_TRY({ *out_id = c->gl->insert_lasso(node->gl.get()); }) }
int hg_circuit_insert_vanilla(hg_circuit* c, size_t input_arity, size_t log2_sub_input_size, size_t num_reps, size_t n_gates, const uint8_t* has_const,
                              const uint64_t* consts, const uint64_t* add_ptr, const uint64_t* add_coef, const uint32_t* add_input,
                              const uint64_t* add_wire, const uint64_t* mul_ptr, const uint64_t* mul_coef, const uint32_t* mul_in0, const uint64_t* mul_w0,
                              const uint32_t* mul_in1, const uint64_t* mul_w1, int* out_id) {
    HG_TRY({
        HG_CUDA(cudaSetDevice(c->ctx->dev.device));
        VanillaDesc d;
        d.arity = input_arity; d.log2_sub = log2_sub_input_size; d.num_reps = num_reps; d.n_gates = n_gates;
        d.has_const.assign(has_const, has_const + n_gates);
        d.consts.assign(consts, consts + n_gates);
        d.add_ptr.assign(add_ptr, add_ptr + n_gates + 1);
        const size_t na = d.add_ptr[n_gates];
        d.add_coef.assign(add_coef, add_coef + na); d.add_in.assign(add_input, add_input + na); d.add_wire.assign(add_wire, add_wire + na);
        d.mul_ptr.assign(mul_ptr, mul_ptr + n_gates + 1);
        const size_t nm = d.mul_ptr[n_gates];
        d.mul_coef.assign(mul_coef, mul_coef + nm); d.mul_in0.assign(mul_in0, mul_in0 + nm); d.mul_w0.assign(mul_w0, mul_w0 + nm);
        d.mul_in1.assign(mul_in1, mul_in1 + nm); d.mul_w1.assign(mul_w1, mul_w1 + nm);
        *out_id = c->gl->insert_vanilla(d);
    })
}
int hg_circuit_connect(hg_circuit* c, int from, int to) { HG_TRY({ c->gl->connect(from, to); }) }
int hg_circuit_evaluate(hg_circuit* c, const void* const* d_inputs, size_t n_inputs) {
    HG_TRY({
        HG_CUDA(cudaSetDevice(c->ctx->dev.device));
        std::vector<const u64*> in;
        for (size_t i = 0; i < n_inputs; i++) in.push_back((const u64*)d_inputs[i]);
        c->gl->evaluate(in);
    })
}
int hg_circuit_node_value(hg_circuit* c, int id, const void** d_ptr, size_t* len) {
    HG_TRY({
        if (id < 0 || (size_t)id >= c->gl->num_nodes()) throw std::runtime_error("no such node");
        *d_ptr = c->gl->node_value(id);
        *len = c->gl->node_out_len(id);
    })
}
int hg_gkr_prove(hg_circuit* c, size_t n_output_claims, const size_t* point_lens, const uint64_t* points_ext, const uint64_t* values_ext,
                 hg_transcript* t, int mode) {
    HG_TRY({
        typedef GlField FP;
        HG_CUDA(cudaSetDevice(c->ctx->dev.device));
        std::vector<GkrCircuitDev<FP>::InputClaim> oc(n_output_claims);
        size_t off = 0;
        for (size_t i = 0; i < n_output_claims; i++) {
            for (size_t k = 0; k < point_lens[i]; k++) oc[i].point.push_back(FP::x_from_limbs(points_ext + 2 * (off + k)));
            off += point_lens[i];
            oc[i].value = FP::x_from_limbs(values_ext + 2 * i);
        }
        c->input_claims = c->gl->prove(*t->gl, mode == HG_MODE_INTERACTIVE ? kModeInteractive : kModePrefetch, c->ctx->wire, oc);
    })
}
void hg_gkr_timing(const hg_circuit* c, double* out_us6) { for (int i = 0; i < 6; i++) out_us6[i] = c->gl->timing()[i]; }
size_t hg_gkr_num_challenges(const hg_circuit* c) { return c->gl->total_challenges(); }
size_t hg_gkr_num_inputs(const hg_circuit* c) { return c->input_claims.size(); }
size_t hg_gkr_num_input_claims(const hg_circuit* c, size_t input) { return input < c->input_claims.size() ? c->input_claims[input].size() : 0; }
size_t hg_gkr_input_claim_num_vars(const hg_circuit* c, size_t input, size_t k) { return c->input_claims.at(input).at(k).point.size(); }
int hg_gkr_input_claim(const hg_circuit* c, size_t input, size_t k, uint64_t* point_ext, uint64_t* value_ext) {
    HG_TRY({
        const auto& ic = c->input_claims.at(input).at(k);
        for (size_t q = 0; q < ic.point.size(); q++) GlField::x_to_limbs(ic.point[q], point_ext + 2 * q);
        GlField::x_to_limbs(ic.value, value_ext);
    })
}

int hg_field_selftest(hg_ctx* ctx, int op, const uint64_t* a_ext, const uint64_t* b_ext, size_t n, uint64_t* out_ext) {
    HG_TRY({
        typedef GlField FP;
        HG_CUDA(cudaSetDevice(ctx->dev.device));
        DevBuf<gl2> a, b, o;
        a.alloc(n); b.alloc(n); o.alloc(n);
        HG_CUDA(cudaMemcpy(a.p, a_ext, n * sizeof(gl2), cudaMemcpyHostToDevice));
        HG_CUDA(cudaMemcpy(b.p, b_ext, n * sizeof(gl2), cudaMemcpyHostToDevice));
        k_selftest<FP><<<(unsigned)((n + 255) / 256), 256, 0, ctx->dev.stream>>>(op, a.p, b.p, n, o.p);
        HG_LAUNCH_CHECK();
        HG_CUDA(cudaStreamSynchronize(ctx->dev.stream));
        HG_CUDA(cudaMemcpy(out_ext, o.p, n * sizeof(gl2), cudaMemcpyDeviceToHost));
    })
}

}  // extern "C"
