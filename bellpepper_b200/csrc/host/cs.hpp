// C++ host-side mirror of the reference's constraint-system interface, backed by the C ABI (include/bp_r1cs.h).
//
//   SynthesisError            constraint_system.rs:21-57
//   ConstraintSystem surface  constraint_system.rs:61-237   (duck-typed: gadgets are templates over CS)
//   Namespace                 constraint_system.rs:242-333  (RAII: pops on destruction, like Drop)
//   TestConstraintSystem      util_cs/test_cs.rs:19-447     (names + paths on the host, values + matrices in HBM)
//   WitnessCS                 crates/bellpepper/src/util_cs/witness_cs.rs:45-201
//
// The surface a gadget sees (same names and meaning as the Rust trait):
//   Variable alloc(name_fn, value_fn)        value_fn returns Fr or throws SynthesisError
//   Variable alloc_input(name_fn, value_fn)
//   void     enforce(name_fn, a, b, c)       a/b/c : LinearCombination -> LinearCombination
//   void     push_namespace(name_fn) / pop_namespace();  Root& get_root();  Namespace<Root> ns(name_fn)
//   static Variable one();  const Field* field()
// name_fn is only invoked by backends that keep names (exactly like the reference's `annotation` closures).
//
// Rows and witness values are batched in host staging vectors and handed to a Sink: DeviceSink forwards to
// bp_cs_alloc / bp_cs_enforce (pinned ring + async H2D inside the library); HostSink keeps them in host memory
// (CPU-only structural tests, export of a CSR sample to the CPU baseline).  Evaluation is ALWAYS the device's.
#pragma once
#include <cstdint>
#include <optional>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

#include "../../../include/bp_r1cs.h"
#include "lc.hpp"

namespace bph {

struct SynthesisError : std::runtime_error {
    enum Kind { AssignmentMissing, DivisionByZero, Unsatisfiable, Native } kind;
    SynthesisError(Kind k, const char* msg) : std::runtime_error(msg), kind(k) {}
    static SynthesisError assignment_missing() { return {AssignmentMissing, "an assignment for a variable could not be computed"}; }
    static SynthesisError division_by_zero() { return {DivisionByZero, "division by zero"}; }
    static SynthesisError unsatisfiable() { return {Unsatisfiable, "unsatisfiable constraint system"}; }
};

// ---- sinks ------------------------------------------------------------------------------------------------
struct Sink {
    virtual ~Sink() {}
    virtual void alloc(int is_aux, const uint64_t* vals, uint64_t n) = 0;
    // Packed batch: one byte per value; wide[i] = (position in the batch, 4 limbs) for the values that do not fit a byte
    // (their byte is a 0 placeholder).
    virtual void alloc_packed(int is_aux, const uint8_t* bytes, uint64_t n, const uint64_t* wide_pos, const uint64_t* wide_vals,
                              uint64_t n_wide) = 0;
    virtual void enforce(uint64_t n_rows, const uint32_t* lens, const uint32_t* cols, const uint64_t* coeffs, uint64_t nnz) = 0;
    virtual bp_cs* handle() { return nullptr; }
};

struct DeviceSink : Sink {
    bp_cs* h;
    explicit DeviceSink(bp_cs* handle_) : h(handle_) {}
    void alloc(int is_aux, const uint64_t* vals, uint64_t n) override {
        uint64_t first;
        if (bp_cs_alloc(h, is_aux, vals, n, &first) != BP_OK) throw SynthesisError(SynthesisError::Native, bp_cs_last_error(h));
    }
    void alloc_packed(int is_aux, const uint8_t* bytes, uint64_t n, const uint64_t* wide_pos, const uint64_t* wide_vals,
                      uint64_t n_wide) override {
        uint64_t first;
        if (bp_cs_alloc_u8(h, is_aux, bytes, n, &first) != BP_OK) throw SynthesisError(SynthesisError::Native, bp_cs_last_error(h));
        for (uint64_t i = 0; i < n_wide; ++i)
            if (bp_cs_set(h, is_aux, first + wide_pos[i], wide_vals + 4 * i) != BP_OK)
                throw SynthesisError(SynthesisError::Native, bp_cs_last_error(h));
    }
    void enforce(uint64_t n_rows, const uint32_t* lens, const uint32_t* cols, const uint64_t* coeffs, uint64_t) override {
        if (bp_cs_enforce(h, n_rows, lens, cols, coeffs) != BP_OK) throw SynthesisError(SynthesisError::Native, bp_cs_last_error(h));
    }
    bp_cs* handle() override { return h; }
};

struct HostSink : Sink {
    std::vector<uint64_t> inputs{1, 0, 0, 0}, aux;  // inputs start with ONE
    std::vector<uint32_t> lens, cols;
    std::vector<uint64_t> coeffs;
    void alloc(int is_aux, const uint64_t* vals, uint64_t n) override {
        auto& v = is_aux ? aux : inputs;
        v.insert(v.end(), vals, vals + 4 * n);
    }
    void alloc_packed(int is_aux, const uint8_t* bytes, uint64_t n, const uint64_t* wide_pos, const uint64_t* wide_vals,
                      uint64_t n_wide) override {
        auto& v = is_aux ? aux : inputs;
        const size_t base = v.size();
        v.resize(base + 4 * n, 0);
        for (uint64_t i = 0; i < n; ++i) v[base + 4 * i] = bytes[i];
        for (uint64_t i = 0; i < n_wide; ++i)
            for (int j = 0; j < 4; ++j) v[base + 4 * wide_pos[i] + j] = wide_vals[4 * i + j];
    }
    void enforce(uint64_t n_rows, const uint32_t* l, const uint32_t* c, const uint64_t* v, uint64_t nnz) override {
        lens.insert(lens.end(), l, l + 3 * n_rows);
        cols.insert(cols.end(), c, c + nnz);
        coeffs.insert(coeffs.end(), v, v + 4 * nnz);
    }
};

// ---- Namespace ----------------------------------------------------------------------------------------------
template <class Root> class Namespace {
  public:
    explicit Namespace(Root& r) : r_(&r) {}
    Namespace(const Namespace&) = delete;
    Namespace(Namespace&& o) noexcept : r_(o.r_) { o.r_ = nullptr; }
    // Never throws (a destructor that throws during unwinding is std::terminate): the stack cannot be empty here, since this
    // object exists because of a push; a root whose stack was tampered with is left as it is.
    ~Namespace() noexcept {
        if (!r_) return;
        try {
            r_->pop_namespace();
        } catch (...) {
        }
    }
    static Variable one() { return one_var(); }
    const Field* field() { return r_->field(); }
    template <class N, class V> Variable alloc(N&& n, V&& v) { return r_->alloc(n, v); }
    template <class N, class V> Variable alloc_input(N&& n, V&& v) { return r_->alloc_input(n, v); }
    template <class N, class A, class B, class C> void enforce(N&& n, A&& a, B&& b, C&& c) { r_->enforce(n, a, b, c); }
    Root& get_root() { return *r_; }
    template <class N> Namespace<Root> ns(N&& n) {
        r_->push_namespace(n);
        return Namespace<Root>(*r_);
    }

  private:
    Root* r_;
};

// ---- TestConstraintSystem ------------------------------------------------------------------------------------
// kNamed = true keeps the reference's path bookkeeping (named_objects, duplicate-path and '/'-in-name panics as
// exceptions, row -> path); kNamed = false skips every annotation closure for bulk circuits.
template <bool kNamed> class TestConstraintSystemT {
  public:
    using Root = TestConstraintSystemT;
    enum ObjKind { kVar, kConstraint, kNamespace };
    struct NamedObject {
        ObjKind kind;
        uint64_t id;  // Variable.tagged or row index
    };

    TestConstraintSystemT(int field, Sink* sink, size_t flush_terms = 1u << 20)
        : field_(field), sink_(sink), flush_terms_(flush_terms) {
        if (kNamed) {
            named_["ONE"] = NamedObject{kVar, one_var().tagged};
            input_names_.push_back("ONE");
        }
        count_[0] = 1;
        count_[1] = 0;
    }
    ~TestConstraintSystemT() {}

    static Variable one() { return one_var(); }
    const Field* field() const { return &field_; }
    Sink* sink() { return sink_; }

    // ---- trait ConstraintSystem (test_cs.rs:377-447) ----
    template <class N, class V> Variable alloc(N&& name, V&& value) { return alloc_impl(1, name, value); }
    template <class N, class V> Variable alloc_input(N&& name, V&& value) { return alloc_impl(0, name, value); }

    template <class N, class A, class B, class C> void enforce(N&& name, A&& a, B&& b, C&& c) {
        if (kNamed) {
            std::string path = compute_path(name());
            set_named_obj(path, NamedObject{kConstraint, n_rows_});
            row_paths_.push_back(std::move(path));
        }
        push_lc(a(LinearCombination::zero(&field_)));
        push_lc(b(LinearCombination::zero(&field_)));
        push_lc(c(LinearCombination::zero(&field_)));
        ++n_rows_;
        if (cols_.size() >= flush_terms_) flush();
    }

    template <class N> void push_namespace(N&& name) {
        if (kNamed) {
            std::string n = name();
            set_named_obj(compute_path(n), NamedObject{kNamespace, 0});
            ns_.push_back(std::move(n));
        }
    }
    void pop_namespace() {
        if (kNamed) {
            if (ns_.empty()) throw std::logic_error("pop_namespace on empty namespace stack");
            ns_.pop_back();
        }
    }
    Root& get_root() { return *this; }
    template <class N> Namespace<Root> ns(N&& name) {
        push_namespace(name);
        return Namespace<Root>(*this);
    }

    // ---- the hot path, on the device (test_cs.rs:239-264) ----
    int64_t first_unsatisfied_row() {
        flush();
        bp_cs* h = need_handle();
        int64_t row;
        if (bp_cs_first_unsatisfied(h, &row) != BP_OK) throw SynthesisError(SynthesisError::Native, bp_cs_last_error(h));
        return row;
    }
    std::optional<std::string> which_is_unsatisfied() {
        const int64_t row = first_unsatisfied_row();
        if (row < 0) return std::nullopt;
        return kNamed ? row_paths_[(size_t)row] : std::to_string(row);
    }
    bool is_satisfied() { return first_unsatisfied_row() < 0; }

    // ---- accessors (test_cs.rs:175-334) ----
    uint64_t num_constraints() const { return n_rows_; }
    uint64_t num_inputs() const { return count_[0]; }
    uint64_t num_aux() const { return count_[1]; }
    void set(const std::string& path, const Fr& to) {
        const Variable v = var_at(path);
        flush();
        bp_cs* h = need_handle();
        if (bp_cs_set(h, v.is_aux(), v.index(), to.l) != BP_OK) throw SynthesisError(SynthesisError::Native, bp_cs_last_error(h));
    }
    Fr get(const std::string& path) {
        const Variable v = var_at(path);
        flush();
        bp_cs* h = need_handle();
        Fr out;
        if (bp_cs_get(h, v.is_aux(), v.index(), out.l) != BP_OK) throw SynthesisError(SynthesisError::Native, bp_cs_last_error(h));
        return out;
    }
    const std::string& row_path(uint64_t row) const { return row_paths_.at(row); }
    const std::vector<std::string>& input_names() const { return input_names_; }
    const std::vector<std::string>& aux_names() const { return aux_names_; }

    void flush() {
        for (int k = 0; k < 2; ++k) {
            if (!pend_[k].empty()) {
                sink_->alloc_packed(k, pend_[k].data(), pend_[k].size(), wide_pos_[k].data(), wide_vals_[k].data(), wide_pos_[k].size());
                pend_[k].clear();
                wide_pos_[k].clear();
                wide_vals_[k].clear();
            }
        }
        if (!lens_.empty()) {
            sink_->enforce(lens_.size() / 3, lens_.data(), cols_.data(), coeffs_.data(), cols_.size());
            lens_.clear();
            cols_.clear();
            coeffs_.clear();
        }
    }

  private:
    template <class N, class V> Variable alloc_impl(int is_aux, N&& name, V&& value) {
        std::string path;
        if (kNamed) path = compute_path(name());
        const Fr v = value();  // may throw: nothing has been registered yet (test_cs.rs:388)
        const uint64_t idx = count_[is_aux]++;
        // packed staging: a byte when the value fits one (AllocatedBit / Boolean witnesses), else a placeholder + the limbs
        if ((v.l[0] >> 8) == 0 && (v.l[1] | v.l[2] | v.l[3]) == 0) {
            pend_[is_aux].push_back((uint8_t)v.l[0]);
        } else {
            wide_pos_[is_aux].push_back(pend_[is_aux].size());
            wide_vals_[is_aux].insert(wide_vals_[is_aux].end(), v.l, v.l + 4);
            pend_[is_aux].push_back(0);
        }
        const Variable var = is_aux ? Variable::aux((uint32_t)idx) : Variable::input((uint32_t)idx);
        if (kNamed) {
            (is_aux ? aux_names_ : input_names_).push_back(path);
            set_named_obj(path, NamedObject{kVar, var.tagged});
        }
        if (pend_[is_aux].size() >= (4u << 20)) flush();  // (rows first: flush() keeps alloc-before-enforce order)
        return var;
    }
    void push_lc(const LinearCombination& lc) {
        const uint32_t n = lc.len();
        lens_.push_back(n);
        const Term* t = lc.terms();
        for (uint32_t i = 0; i < n; ++i) {
            cols_.push_back(t[i].col);
            coeffs_.insert(coeffs_.end(), t[i].coeff.l, t[i].coeff.l + 4);
        }
    }
    std::string compute_path(const std::string& name) const {  // test_cs.rs:363-375
        if (name.find('/') != std::string::npos) throw std::logic_error("'/' is not allowed in names");
        std::string out;
        for (const auto& s : ns_) {
            out += s;
            out += '/';
        }
        return out + name;
    }
    void set_named_obj(const std::string& path, NamedObject o) {  // test_cs.rs:325-333
        if (!named_.emplace(path, o).second) throw std::logic_error("tried to create object at existing path: " + path);
    }
    Variable var_at(const std::string& path) const {
        static_assert(kNamed, "set/get by path need a named constraint system");
        auto it = named_.find(path);
        if (it == named_.end()) throw std::out_of_range("no variable exists at path: " + path);
        if (it->second.kind != kVar) throw std::out_of_range("path `" + path + "` is not a variable");
        return Variable{(uint32_t)it->second.id};
    }
    bp_cs* need_handle() {
        bp_cs* h = sink_->handle();
        if (!h) throw SynthesisError(SynthesisError::Native, "this constraint system records on the host only; evaluation needs a device handle");
        return h;
    }

    Field field_;
    Sink* sink_;
    size_t flush_terms_;
    uint64_t n_rows_ = 0;
    uint64_t count_[2];
    std::vector<uint8_t> pend_[2];
    std::vector<uint64_t> wide_pos_[2], wide_vals_[2];
    std::vector<uint32_t> lens_, cols_;
    std::vector<uint64_t> coeffs_;
    std::unordered_map<std::string, NamedObject> named_;
    std::vector<std::string> ns_, row_paths_, input_names_, aux_names_;
};

// ---- WitnessCS (crates/bellpepper/src/util_cs/witness_cs.rs:45-201) ------------------------------------------------
// The flat witness container: no names, no constraints (`enforce` is a no-op and never runs its closures,
// witness_cs.rs:125-134), values appended in allocation order.  The two assignment vectors live in HBM behind the
// handle; allocations are staged packed like in TestConstraintSystemT.
class WitnessCS {
  public:
    using Root = WitnessCS;
    // `new()` (witness_cs.rs:94-101): input_assignment = [ONE], aux_assignment = []: exactly a fresh handle
    WitnessCS(int field, bp_cs* handle) : field_(field), h_(handle) { count_[0] = 1; count_[1] = 0; }
    static Variable one() { return one_var(); }
    const Field* field() const { return &field_; }

    template <class N, class V> Variable alloc(N&&, V&& value) { return push(1, value()); }        // witness_cs.rs:103-112
    template <class N, class V> Variable alloc_input(N&&, V&& value) { return push(0, value()); }  // witness_cs.rs:114-123
    template <class N, class A, class B, class C> void enforce(N&&, A&&, B&&, C&&) {}              // witness_cs.rs:125-134
    template <class N> void push_namespace(N&&) {}
    void pop_namespace() {}
    Root& get_root() { return *this; }
    template <class N> Namespace<Root> ns(N&&) { return Namespace<Root>(*this); }

    static bool is_extensible() { return true; }  // witness_cs.rs:150-152
    void extend(WitnessCS& other) {               // witness_cs.rs:154-163: other's inputs without its ONE, then its aux
        const std::vector<Fr> in = other.inputs_slice(), ax = other.aux_slice();
        extend_inputs(in.data() + 1, in.size() - 1);
        extend_aux(ax.data(), ax.size());
    }
    static bool is_witness_generator() { return true; }  // witness_cs.rs:167-169
    void extend_inputs(const Fr* v, size_t n) {          // witness_cs.rs:171-173
        for (size_t i = 0; i < n; ++i) push(0, v[i]);
    }
    void extend_aux(const Fr* v, size_t n) {  // witness_cs.rs:175-177
        for (size_t i = 0; i < n; ++i) push(1, v[i]);
    }
    // witness_cs.rs:179-193: both vectors grow by zeros; the reference returns the two new tails (aux first) as mutable
    // slices -- here their first indices, to be filled with fill_aux / fill_inputs (one bulk upload each)
    std::pair<uint64_t, uint64_t> allocate_empty(size_t aux_n, size_t inputs_n) {
        const uint64_t a0 = count_[1], i0 = count_[0];
        for (size_t i = 0; i < aux_n; ++i) push(1, Fr::zero());
        for (size_t i = 0; i < inputs_n; ++i) push(0, Fr::zero());
        return {a0, i0};
    }
    void fill_aux(uint64_t first, const Fr* v, size_t n) { fill(1, first, v, n); }
    void fill_inputs(uint64_t first, const Fr* v, size_t n) { fill(0, first, v, n); }
    std::vector<Fr> inputs_slice() { return slice(0); }  // witness_cs.rs:195-197
    std::vector<Fr> aux_slice() { return slice(1); }     // witness_cs.rs:199-201
    uint64_t num_inputs() const { return count_[0]; }
    uint64_t num_aux() const { return count_[1]; }

    void flush() {
        for (int k = 0; k < 2; ++k) {
            if (pend_[k].empty()) continue;
            uint64_t first;
            if (bp_cs_alloc_u8(h_, k, pend_[k].data(), pend_[k].size(), &first) != BP_OK) fail();
            for (size_t i = 0; i < wide_pos_[k].size(); ++i)
                if (bp_cs_set(h_, k, first + wide_pos_[k][i], wide_vals_[k].data() + 4 * i) != BP_OK) fail();
            pend_[k].clear();
            wide_pos_[k].clear();
            wide_vals_[k].clear();
        }
    }

  private:
    Variable push(int is_aux, const Fr& v) {
        const uint64_t idx = count_[is_aux]++;
        if ((v.l[0] >> 8) == 0 && (v.l[1] | v.l[2] | v.l[3]) == 0) {
            pend_[is_aux].push_back((uint8_t)v.l[0]);
        } else {
            wide_pos_[is_aux].push_back(pend_[is_aux].size());
            wide_vals_[is_aux].insert(wide_vals_[is_aux].end(), v.l, v.l + 4);
            pend_[is_aux].push_back(0);
        }
        if (pend_[is_aux].size() >= (4u << 20)) flush();
        return is_aux ? Variable::aux((uint32_t)idx) : Variable::input((uint32_t)idx);
    }
    void fill(int is_aux, uint64_t first, const Fr* v, size_t n) {
        flush();
        static_assert(sizeof(Fr) == 32, "Fr is 4 limbs");
        if (n && bp_cs_set_range(h_, is_aux, first, n, v[0].l) != BP_OK) fail();
    }
    std::vector<Fr> slice(int is_aux) {
        flush();
        std::vector<Fr> out(count_[is_aux]);
        if (!out.empty() && bp_cs_witness(h_, is_aux, 0, out.size(), out[0].l) != BP_OK) fail();
        return out;
    }
    [[noreturn]] void fail() { throw SynthesisError(SynthesisError::Native, bp_cs_last_error(h_)); }

    Field field_;
    bp_cs* h_;
    uint64_t count_[2];
    std::vector<uint8_t> pend_[2];
    std::vector<uint64_t> wide_pos_[2], wide_vals_[2];
};

using TestConstraintSystem = TestConstraintSystemT<true>;
using BulkConstraintSystem = TestConstraintSystemT<false>;

}  // namespace bph
