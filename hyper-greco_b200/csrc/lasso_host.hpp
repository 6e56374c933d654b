// Host-side mirror of the reference's lookup plugin API and Lasso preprocessing, kept so that the CUDA path drops in
// behind the same interface:
//   trait LassoSubtable / LookupType          /root/reference/lasso/src/table.rs:16-67
//   SubtableIndices                           /root/reference/lasso/src/table.rs:107-150
//   FullLimbSubtable, BoundSubtable           /root/reference/lasso/src/table/range.rs:11-175
//   RangeLookup                               /root/reference/lasso/src/table/range.rs:177-274
//   LassoPreprocessing::preprocess            /root/reference/lasso/src/lasso.rs:527-627
// Field-independent: subtable entries are produced as u64 integers (`F::from(u64)` in the reference).
#pragma once
#include <algorithm>
#include <cstdint>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace hg {

typedef std::string SubtableId;  // table.rs:13
typedef std::string LookupId;    // table.rs:14

inline unsigned ilog2u(uint64_t x) { return 63u - (unsigned)__builtin_clzll(x); }

// table.rs:107-150 (FixedBitSet over dimension indices)
class SubtableIndices {
  public:
    SubtableIndices() : bits_(0) {}
    static SubtableIndices from_index(unsigned i) { SubtableIndices s; s.bits_ = 1ULL << i; return s; }
    static SubtableIndices from_range(unsigned lo, unsigned hi) { SubtableIndices s; for (unsigned i = lo; i < hi; i++) s.bits_ |= 1ULL << i; return s; }
    static SubtableIndices from_mask(uint64_t mask) { SubtableIndices s; s.bits_ = mask; return s; }
    void union_with(const SubtableIndices& o) { bits_ |= o.bits_; }
    bool contains(unsigned i) const { return (bits_ >> i) & 1; }
    size_t len() const { return (size_t)__builtin_popcountll(bits_); }
    std::vector<unsigned> iter() const { std::vector<unsigned> v; for (unsigned i = 0; i < 64; i++) if (contains(i)) v.push_back(i); return v; }
  private:
    uint64_t bits_;
};

// table.rs:16-33
class LassoSubtable {
  public:
    virtual ~LassoSubtable() {}
    virtual SubtableId subtable_id() const = 0;
    // materialize(M): entry i as an integer to be lifted with F::from(u64)
    virtual std::vector<uint64_t> materialize(size_t M) const = 0;
};

// range.rs:11-49
class FullLimbSubtable : public LassoSubtable {
  public:
    SubtableId subtable_id() const override { return "full"; }
    std::vector<uint64_t> materialize(size_t M) const override {
        std::vector<uint64_t> t(M);
        for (size_t i = 0; i < M; i++) t[i] = i;
        return t;
    }
};

// range.rs:51-175. cutoff = 2^(ilog2(bound) % log2 M) + bound % M  (reference behaviour Q5, kept as is)
class BoundSubtable : public LassoSubtable {
  public:
    explicit BoundSubtable(uint64_t bound) : bound_(bound) {}
    SubtableId subtable_id() const override { return "bound_" + std::to_string(bound_); }
    static uint64_t cutoff(uint64_t bound, size_t M) {
        unsigned log2M = ilog2u(M);
        return (1ULL << (ilog2u(bound) % log2M)) + bound % M;
    }
    std::vector<uint64_t> materialize(size_t M) const override {
        uint64_t c = cutoff(bound_, M);
        std::vector<uint64_t> t(M, 0);
        for (size_t i = 0; i < M && i < c; i++) t[i] = i;
        return t;
    }
  private:
    uint64_t bound_;
};

// table.rs:35-67
class LookupType {
  public:
    virtual ~LookupType() {}
    virtual LookupId lookup_id() const = 0;
    virtual std::vector<std::pair<std::shared_ptr<LassoSubtable>, SubtableIndices>> subtables(size_t C, size_t M) const = 0;
    virtual std::vector<unsigned> chunk_bits(size_t M) const = 0;
    // weight of the t-th operand in combine_lookups / combine_lookup_expressions, as (base, exponent): base^t
    virtual uint64_t combine_weight_base(size_t M) const = 0;
};

// range.rs:177-274
class RangeLookup : public LookupType {
  public:
    explicit RangeLookup(uint64_t bound) : bound_(bound) {
        if (bound == 0) throw std::invalid_argument("RangeLookup: bound must be positive (ilog2 of 0 panics in the reference)");
    }
    static LookupId id_for(uint64_t bound) { return "range_" + std::to_string(bound); }
    LookupId lookup_id() const override { return id_for(bound_); }
    uint64_t bound() const { return bound_; }
    std::vector<std::pair<std::shared_ptr<LassoSubtable>, SubtableIndices>> subtables(size_t, size_t M) const override {
        auto full = std::make_shared<FullLimbSubtable>();
        auto rem = std::make_shared<BoundSubtable>(bound_);
        unsigned num_chunks = ilog2u(bound_) / ilog2u(M);
        if (bound_ % M == 0) return {{full, SubtableIndices::from_range(0, num_chunks)}};
        if (bound_ < M) return {{rem, SubtableIndices::from_index(0)}};
        return {{full, SubtableIndices::from_range(0, num_chunks)}, {rem, SubtableIndices::from_index(num_chunks)}};
    }
    std::vector<unsigned> chunk_bits(size_t M) const override {
        unsigned log2M = ilog2u(M), bound_bits = ilog2u(bound_);
        std::vector<unsigned> out(bound_bits / log2M, log2M);
        if (bound_ % M != 0) out.push_back(ilog2u(BoundSubtable::cutoff(bound_, M)));
        return out;
    }
    uint64_t combine_weight_base(size_t M) const override { return M; }  // range.rs:184-204
  private:
    uint64_t bound_;
};

// ---- plug-ins: a LassoSubtable / LookupType described by DATA, for callers that bring their own tables through the C ABI
// (hg_lasso_preprocess_lookups). The traits of table.rs:16-67 carry code (materialize, evaluate_mle, combine_lookups, ...); what the
// prove / verify path needs from them is: the subtable's M entries (materialize; its MLE is evaluated from them, the MLE being
// unique), the dimensions it serves (SubtableIndices), the bit widths of the chunks (chunk_bits) and the combination
// g(operands) = sum_t w^t operand_t (combine_lookups / combine_lookup_expressions of every lookup in the reference, range.rs:184-204).
// subtable_indices is the uniform split into log2(M)-bit chunks (range.rs:254-256).
class TableSubtable : public LassoSubtable {
  public:
    TableSubtable(SubtableId id, std::vector<uint64_t> table) : id_(std::move(id)), table_(std::move(table)) {}
    SubtableId subtable_id() const override { return id_; }
    std::vector<uint64_t> materialize(size_t M) const override {
        if (table_.size() != M) throw std::invalid_argument("TableSubtable '" + id_ + "': the table has " + std::to_string(table_.size()) + " entries, M = " + std::to_string(M));
        return table_;
    }
  private:
    SubtableId id_;
    std::vector<uint64_t> table_;
};
class TableLookup : public LookupType {
  public:
    TableLookup(LookupId id, std::vector<std::pair<std::shared_ptr<LassoSubtable>, SubtableIndices>> subtables, std::vector<unsigned> chunk_bits, uint64_t weight)
        : id_(std::move(id)), subtables_(std::move(subtables)), chunk_bits_(std::move(chunk_bits)), weight_(weight) {}
    LookupId lookup_id() const override { return id_; }
    std::vector<std::pair<std::shared_ptr<LassoSubtable>, SubtableIndices>> subtables(size_t, size_t) const override { return subtables_; }
    std::vector<unsigned> chunk_bits(size_t) const override { return chunk_bits_; }
    uint64_t combine_weight_base(size_t) const override { return weight_; }
  private:
    LookupId id_;
    std::vector<std::pair<std::shared_ptr<LassoSubtable>, SubtableIndices>> subtables_;
    std::vector<unsigned> chunk_bits_;
    uint64_t weight_;
};

// lasso.rs:513-523
struct LassoPreprocessing {
    size_t C = 4, M = 1 << 16;
    std::vector<std::shared_ptr<LookupType>> lookups;  // BTreeMap<LookupId, _> order: ascending id STRING, deduplicated
    std::map<LookupId, size_t> lookup_id_to_index;
    std::vector<std::shared_ptr<LassoSubtable>> subtables_by_idx;
    std::vector<std::vector<size_t>> subtable_to_memory_indices, lookup_to_memory_indices;
    std::vector<size_t> memory_to_subtable_index, memory_to_dimension_index;
    size_t num_memories = 0;

    // lasso.rs:527-627
    static LassoPreprocessing preprocess(const std::vector<std::shared_ptr<LookupType>>& lookups_in, size_t C, size_t M) {
        LassoPreprocessing pp;
        pp.C = C; pp.M = M;
        std::map<LookupId, std::shared_ptr<LookupType>> ordered;
        for (auto& l : lookups_in) ordered[l->lookup_id()] = l;  // later duplicates replace earlier ones, like BTreeMap::from_iter
        for (auto& kv : ordered) { pp.lookup_id_to_index[kv.first] = pp.lookups.size(); pp.lookups.push_back(kv.second); }

        std::map<SubtableId, size_t> subtable_id_to_index;
        for (auto& l : pp.lookups)
            for (auto& st : l->subtables(C, M)) {
                SubtableId id = st.first->subtable_id();
                if (!subtable_id_to_index.count(id)) { subtable_id_to_index[id] = pp.subtables_by_idx.size(); pp.subtables_by_idx.push_back(st.first); }
            }
        std::vector<SubtableIndices> subtable_indices(pp.subtables_by_idx.size());
        for (auto& l : pp.lookups)
            for (auto& st : l->subtables(C, M)) subtable_indices[subtable_id_to_index[st.first->subtable_id()]].union_with(st.second);

        for (size_t s = 0; s < subtable_indices.size(); s++) {
            std::vector<size_t> mems;
            for (unsigned d : subtable_indices[s].iter()) {
                mems.push_back(pp.num_memories++);
                pp.memory_to_subtable_index.push_back(s);
                pp.memory_to_dimension_index.push_back(d);
            }
            pp.subtable_to_memory_indices.push_back(mems);
        }
        pp.lookup_to_memory_indices.assign(pp.lookups.size(), {});
        for (size_t li = 0; li < pp.lookups.size(); li++)
            for (auto& st : pp.lookups[li]->subtables(C, M))
                for (size_t mi : pp.subtable_to_memory_indices[subtable_id_to_index[st.first->subtable_id()]])
                    if (st.second.contains((unsigned)pp.memory_to_dimension_index[mi])) pp.lookup_to_memory_indices[li].push_back(mi);
        return pp;
    }
    // lasso.rs:629-637
    std::vector<std::vector<uint64_t>> materialize_subtables() const {
        std::vector<std::vector<uint64_t>> out;
        for (auto& s : subtables_by_idx) out.push_back(s->materialize(M));
        return out;
    }
};

}  // namespace hg
