// Variable / LinearCombination for the C++ host front-end.
// Mirrors /root/reference/crates/bellpepper-core/src/lc.rs: `Variable(Index)` (:8-30), the index-sorted,
// key-unique term list with "same key adds coefficients, zero coefficients are kept" (:40-129), and the
// operator algebra (:270-375).  One sorted list keyed by the tagged column (bit 31 = aux) gives the reference's
// iteration order -- inputs first, then aux, ascending index (:155-160) -- for free.
#pragma once
#include <cstdint>
#include <cstdlib>
#include <algorithm>
#include <cstring>
#include <utility>
#include <vector>

#include "fr.hpp"

namespace bph {

constexpr uint32_t kAuxTag = 0x80000000u;

struct Variable {
    uint32_t tagged;
    static Variable input(uint32_t i) { return Variable{i}; }
    static Variable aux(uint32_t i) { return Variable{i | kAuxTag}; }
    bool is_aux() const { return (tagged & kAuxTag) != 0; }
    uint32_t index() const { return tagged & ~kAuxTag; }
    bool operator==(const Variable& o) const { return tagged == o.tagged; }
};
inline Variable one_var() { return Variable::input(0); }  // constraint_system.rs:73-75

struct Term {
    uint32_t col;
    Fr coeff;
};

// Term list with inline storage for the short LCs that make up almost every gadget constraint.
class TermVec {
  public:
    static constexpr uint32_t kInline = 6;
    TermVec() : n_(0), cap_(kInline), p_(inl_) {}
    TermVec(const TermVec& o) : n_(0), cap_(kInline), p_(inl_) { assign(o); }
    TermVec(TermVec&& o) noexcept : n_(0), cap_(kInline), p_(inl_) { steal(o); }
    TermVec& operator=(const TermVec& o) {
        if (this != &o) assign(o);
        return *this;
    }
    TermVec& operator=(TermVec&& o) noexcept {
        if (this != &o) {
            release();
            steal(o);
        }
        return *this;
    }
    ~TermVec() { release(); }
    uint32_t size() const { return n_; }
    const Term* data() const { return p_; }
    Term* data() { return p_; }
    const Term& operator[](uint32_t i) const { return p_[i]; }
    Term& operator[](uint32_t i) { return p_[i]; }
    void clear() { n_ = 0; }
    void reserve(uint32_t c) {
        if (c <= cap_) return;
        uint32_t nc = cap_ * 2 > c ? cap_ * 2 : c;
        Term* np = (Term*)std::malloc(sizeof(Term) * (size_t)nc);
        if (!np) std::abort();
        std::memcpy(np, p_, sizeof(Term) * (size_t)n_);
        if (p_ != inl_) std::free(p_);
        p_ = np;
        cap_ = nc;
    }
    void insert_at(uint32_t i, const Term& t) {
        reserve(n_ + 1);
        std::memmove(p_ + i + 1, p_ + i, sizeof(Term) * (size_t)(n_ - i));
        p_[i] = t;
        ++n_;
    }
    void push_back(const Term& t) {
        reserve(n_ + 1);
        p_[n_++] = t;
    }

  private:
    void release() {
        if (p_ != inl_) std::free(p_);
        p_ = inl_;
        cap_ = kInline;
        n_ = 0;
    }
    void assign(const TermVec& o) {
        n_ = 0;
        reserve(o.n_);
        std::memcpy(p_, o.p_, sizeof(Term) * (size_t)o.n_);
        n_ = o.n_;
    }
    void steal(TermVec& o) {
        if (o.p_ == o.inl_) {
            std::memcpy(inl_, o.inl_, sizeof(Term) * (size_t)o.n_);
            p_ = inl_;
            cap_ = kInline;
        } else {
            p_ = o.p_;
            cap_ = o.cap_;
            o.p_ = o.inl_;
            o.cap_ = kInline;
        }
        n_ = o.n_;
        o.n_ = 0;
    }
    uint32_t n_, cap_;
    Term* p_;
    Term inl_[kInline];
};

class LinearCombination {
  public:
    explicit LinearCombination(const Field* f) : f_(f) {}
    static LinearCombination zero(const Field* f) { return LinearCombination(f); }
    static LinearCombination from_coeff(const Field* f, Variable v, const Fr& c) {
        LinearCombination lc(f);
        lc.add_term(v, c);
        return lc;
    }
    static LinearCombination from_variable(const Field* f, Variable v) { return from_coeff(f, v, Fr::one()); }

    const Field* field() const { return f_; }
    uint32_t len() const { return t_.size(); }
    bool is_empty() const { return t_.size() == 0; }
    const Term* terms() const { return t_.data(); }

    // lc.rs:74-113 `insert_or_update`: sorted insert; an existing key accumulates (zero results are kept).
    void add_term(Variable v, const Fr& c) {
        const uint32_t key = v.tagged, n = t_.size();
        if (n == 0 || t_[n - 1].col < key) {  // append: the common case while building gadget LCs
            t_.push_back(Term{key, c});
            return;
        }
        uint32_t lo = 0, hi = n;
        while (lo < hi) {
            const uint32_t mid = (lo + hi) / 2;
            if (t_[mid].col < key) lo = mid + 1; else hi = mid;
        }
        if (lo < n && t_[lo].col == key) t_[lo].coeff = f_->add(t_[lo].coeff, c);
        else t_.insert_at(lo, Term{key, c});
    }
    void sub_term(Variable v, const Fr& c) { add_term(v, f_->neg(c)); }
    // lc + &other / lc - &other / lc +- (k, &other)  (lc.rs:311-375).  Same result as inserting other's terms one by
    // one, computed as a single merge of the two sorted lists so that MultiEq's ~1000-term rows stay linear-time.
    // Many single-term additions at once (UInt32::addmany adds 32 bits per operand): same final list as calling
    // add_term for each in order -- field addition commutes -- but one sort + one merge instead of n sorted inserts.
    void add_terms_bulk(std::vector<Term>& ts) {
        if (ts.empty()) return;
        std::stable_sort(ts.begin(), ts.end(), [](const Term& a, const Term& b) { return a.col < b.col; });
        LinearCombination tmp(f_);
        tmp.t_.reserve((uint32_t)ts.size());
        for (const Term& t : ts) {
            const uint32_t n = tmp.t_.size();
            if (n && tmp.t_[n - 1].col == t.col) tmp.t_[n - 1].coeff = f_->add(tmp.t_[n - 1].coeff, t.coeff);
            else tmp.t_.push_back(t);
        }
        if (t_.size() == 0) t_ = std::move(tmp.t_);
        else merge(tmp, nullptr, false);
        ts.clear();
    }
    void add_lc(const LinearCombination& o) { merge(o, nullptr, false); }
    void sub_lc(const LinearCombination& o) { merge(o, nullptr, true); }
    void add_scaled(const Fr& k, const LinearCombination& o) { merge(o, &k, false); }
    void add_scaled_pow2(unsigned bits, const LinearCombination& o) { merge(o, nullptr, false, (int)bits); }
    void sub_scaled(const Fr& k, const LinearCombination& o) { merge(o, &k, true); }

  private:
    void merge(const LinearCombination& o, const Fr* k, bool negate, int pow2_bits = -1) {
        const uint32_t n = t_.size(), m = o.t_.size();
        if (m == 0) return;
        auto scaled = [&](const Fr& c0) {
            Fr c = pow2_bits >= 0 ? f_->mul_pow2(c0, (unsigned)pow2_bits) : (k ? f_->mul(c0, *k) : c0);
            return negate ? f_->neg(c) : c;
        };
        auto coeff_of = [&](uint32_t j) { return scaled(o.t_[j].coeff); };
        if (&o == this) {  // aliased: work from a copy
            LinearCombination tmp(o);
            merge(tmp, k, negate, pow2_bits);
            return;
        }
        if (m <= 2) {  // tiny: plain inserts
            for (uint32_t j = 0; j < m; ++j) add_term(Variable{o.t_[j].col}, coeff_of(j));
            return;
        }
        TermVec out;
        out.reserve(n + m);
        uint32_t i = 0, j = 0;
        while (i < n && j < m) {
            if (t_[i].col < o.t_[j].col) out.push_back(t_[i++]);
            else if (t_[i].col > o.t_[j].col) { out.push_back(Term{o.t_[j].col, coeff_of(j)}); ++j; }
            else { out.push_back(Term{t_[i].col, f_->add(t_[i].coeff, coeff_of(j))}); ++i; ++j; }
        }
        while (i < n) out.push_back(t_[i++]);
        while (j < m) { out.push_back(Term{o.t_[j].col, coeff_of(j)}); ++j; }
        t_ = std::move(out);
    }

    const Field* f_;
    TermVec t_;
};

// Operator forms, consuming the left operand like Rust's `self` (lc.rs:270-375).
using CoeffVar = std::pair<Fr, Variable>;
using CoeffLc = std::pair<Fr, const LinearCombination*>;

inline LinearCombination operator+(LinearCombination lc, Variable v) { lc.add_term(v, Fr::one()); return lc; }
inline LinearCombination operator-(LinearCombination lc, Variable v) { lc.sub_term(v, Fr::one()); return lc; }
inline LinearCombination operator+(LinearCombination lc, const CoeffVar& cv) { lc.add_term(cv.second, cv.first); return lc; }
inline LinearCombination operator-(LinearCombination lc, const CoeffVar& cv) { lc.sub_term(cv.second, cv.first); return lc; }
inline LinearCombination operator+(LinearCombination lc, const LinearCombination& o) { lc.add_lc(o); return lc; }
inline LinearCombination operator-(LinearCombination lc, const LinearCombination& o) { lc.sub_lc(o); return lc; }
inline LinearCombination operator+(LinearCombination lc, const CoeffLc& cl) { lc.add_scaled(cl.first, *cl.second); return lc; }
inline LinearCombination operator-(LinearCombination lc, const CoeffLc& cl) { lc.sub_scaled(cl.first, *cl.second); return lc; }

}  // namespace bph
