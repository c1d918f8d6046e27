// Gadget front-end: the circuit builders that produce BASELINE configs 1-3, as C++ templates over any
// constraint system with the surface described in cs.hpp.  They only call alloc / enforce / ns, exactly like the
// reference's gadgets, so "existing gadgets synthesize unchanged" is shown by construction.
//
//   AllocatedBit, Boolean   /root/reference/crates/bellpepper-core/src/gadgets/boolean.rs:10-272, 369-766
//   UInt32                  /root/reference/crates/bellpepper/src/gadgets/uint32.rs:14-406
//   MultiEq                 /root/reference/crates/bellpepper/src/gadgets/multieq.rs:6-122
//   sha256                  /root/reference/crates/bellpepper/src/gadgets/sha256.rs:16-272
//   blake2s                 /root/reference/crates/bellpepper/src/gadgets/blake2s.rs:29-406
// Pinned by the reference's structural KATs (constraint counts) and digest KATs through tests/.
#pragma once
#include <array>
#include <optional>
#include <string>
#include <vector>

#include "cs.hpp"
#include "wtape.hpp"

namespace bph {

using OptBool = int8_t;  // -1 = None
constexpr OptBool kNone = -1;
constexpr unsigned kCapacity = 254;  // ff::PrimeField::CAPACITY (all three fields)

inline Fr bit_value(OptBool v) {
    if (v < 0) throw SynthesisError::assignment_missing();
    return v ? Fr::one() : Fr::zero();
}

struct AllocatedBit {
    Variable variable;
    OptBool value;

    template <class CS> static AllocatedBit alloc(CS&& cs, OptBool value) {  // boolean.rs:68-97
        const Variable var = cs.alloc([] { return std::string("boolean"); }, [&] { return bit_value(value); });
        if (g_tape) g_tape->on_alloc(var);  // a free bit, or the next bit of the addmany sum being recorded
        cs.enforce([] { return std::string("boolean constraint"); },
                   [&](LinearCombination lc) { return std::move(lc) + one_var() - var; },
                   [&](LinearCombination lc) { return std::move(lc) + var; },
                   [&](LinearCombination lc) { return lc; });
        return AllocatedBit{var, value};
    }

    // boolean.rs:27-64: a bit that must be false whenever `must_be_false` is true:  (1 - must_be_false - a) * a = 0
    template <class CS> static AllocatedBit alloc_conditionally(CS&& cs, OptBool value, const AllocatedBit& must_be_false) {
        const Variable var = cs.alloc([] { return std::string("boolean"); }, [&] { return bit_value(value); });
        if (g_tape) g_tape->fail("alloc_conditionally is not a recorded gadget");
        cs.enforce([] { return std::string("boolean constraint"); },
                   [&](LinearCombination lc) { return std::move(lc) + one_var() - must_be_false.variable - var; },
                   [&](LinearCombination lc) { return std::move(lc) + var; },
                   [&](LinearCombination lc) { return lc; });
        return AllocatedBit{var, value};
    }

    template <class CS> static AllocatedBit xor_(CS&& cs, const AllocatedBit& a, const AllocatedBit& b) {  // boolean.rs:101-151
        OptBool rv = kNone;
        const Variable r = cs.alloc([] { return std::string("xor result"); }, [&] {
            if (a.value < 0 || b.value < 0) throw SynthesisError::assignment_missing();
            rv = a.value ^ b.value;
            return rv ? Fr::one() : Fr::zero();
        });
        if (g_tape) g_tape->on_op(kTapeXor, r, WitnessTape::operand(a.variable, false), WitnessTape::operand(b.variable, false));
        cs.enforce([] { return std::string("xor constraint"); },
                   [&](LinearCombination lc) { return std::move(lc) + a.variable + a.variable; },
                   [&](LinearCombination lc) { return std::move(lc) + b.variable; },
                   [&](LinearCombination lc) { return std::move(lc) + a.variable + b.variable - r; });
        return AllocatedBit{r, rv};
    }

    template <class CS> static AllocatedBit and_(CS&& cs, const AllocatedBit& a, const AllocatedBit& b) {  // boolean.rs:155-191
        OptBool rv = kNone;
        const Variable r = cs.alloc([] { return std::string("and result"); }, [&] {
            if (a.value < 0 || b.value < 0) throw SynthesisError::assignment_missing();
            rv = a.value & b.value;
            return rv ? Fr::one() : Fr::zero();
        });
        if (g_tape) g_tape->on_op(kTapeAnd, r, WitnessTape::operand(a.variable, false), WitnessTape::operand(b.variable, false));
        cs.enforce([] { return std::string("and constraint"); },
                   [&](LinearCombination lc) { return std::move(lc) + a.variable; },
                   [&](LinearCombination lc) { return std::move(lc) + b.variable; },
                   [&](LinearCombination lc) { return std::move(lc) + r; });
        return AllocatedBit{r, rv};
    }

    template <class CS> static AllocatedBit and_not(CS&& cs, const AllocatedBit& a, const AllocatedBit& b) {  // boolean.rs:195-231
        OptBool rv = kNone;
        const Variable r = cs.alloc([] { return std::string("and not result"); }, [&] {
            if (a.value < 0 || b.value < 0) throw SynthesisError::assignment_missing();
            rv = a.value & (b.value ^ 1);
            return rv ? Fr::one() : Fr::zero();
        });
        if (g_tape) g_tape->on_op(kTapeAndNot, r, WitnessTape::operand(a.variable, false), WitnessTape::operand(b.variable, false));
        cs.enforce([] { return std::string("and not constraint"); },
                   [&](LinearCombination lc) { return std::move(lc) + a.variable; },
                   [&](LinearCombination lc) { return std::move(lc) + one_var() - b.variable; },
                   [&](LinearCombination lc) { return std::move(lc) + r; });
        return AllocatedBit{r, rv};
    }

    template <class CS> static AllocatedBit nor(CS&& cs, const AllocatedBit& a, const AllocatedBit& b) {  // boolean.rs:235-271
        OptBool rv = kNone;
        const Variable r = cs.alloc([] { return std::string("nor result"); }, [&] {
            if (a.value < 0 || b.value < 0) throw SynthesisError::assignment_missing();
            rv = (a.value ^ 1) & (b.value ^ 1);
            return rv ? Fr::one() : Fr::zero();
        });
        if (g_tape) g_tape->on_op(kTapeNor, r, WitnessTape::operand(a.variable, false), WitnessTape::operand(b.variable, false));
        cs.enforce([] { return std::string("nor constraint"); },
                   [&](LinearCombination lc) { return std::move(lc) + one_var() - a.variable; },
                   [&](LinearCombination lc) { return std::move(lc) + one_var() - b.variable; },
                   [&](LinearCombination lc) { return std::move(lc) + r; });
        return AllocatedBit{r, rv};
    }
};

struct Boolean {  // boolean.rs:368-376
    enum Kind : uint8_t { Is, Not, Constant } kind;
    AllocatedBit bit;  // Is / Not
    bool c;            // Constant

    static Boolean constant(bool b) { return Boolean{Constant, AllocatedBit{Variable{0}, kNone}, b}; }
    static Boolean from(const AllocatedBit& b) { return Boolean{Is, b, false}; }
    bool is_constant() const { return kind == Constant; }
    OptBool get_value() const {  // boolean.rs:429-435
        if (kind == Constant) return c ? 1 : 0;
        if (bit.value < 0) return kNone;
        return kind == Is ? bit.value : (OptBool)(bit.value ^ 1);
    }
    Boolean not_() const {  // boolean.rs:463-469
        if (kind == Constant) return constant(!c);
        return Boolean{kind == Is ? Not : Is, bit, false};
    }
    // boolean.rs:437-455
    LinearCombination lc(const Field* f, Variable one, const Fr& coeff) const {
        LinearCombination z(f);
        if (kind == Constant) {
            if (c) z.add_term(one, coeff);
        } else if (kind == Is) {
            z.add_term(bit.variable, coeff);
        } else {
            z.add_term(one, coeff);
            z.sub_term(bit.variable, coeff);
        }
        return z;
    }
    // the terms of `coeff * self` appended to a scratch list (same terms as `self.lc(one, coeff)`)
    void terms_into(std::vector<Term>& out, const Field* f, Variable one, const Fr& coeff) const {
        if (kind == Constant) {
            if (c) out.push_back(Term{one.tagged, coeff});
        } else if (kind == Is) {
            out.push_back(Term{bit.variable.tagged, coeff});
        } else {
            out.push_back(Term{one.tagged, coeff});
            out.push_back(Term{bit.variable.tagged, f->neg(coeff)});
        }
    }

    uint32_t tape_operand() const {  // witness tape (wtape.hpp): constant / variable / negated variable
        if (kind == Constant) return c ? kOpConst1 : kOpConst0;
        return WitnessTape::operand(bit.variable, kind == Not);
    }

    template <class CS> static Boolean xor_(CS&& cs, const Boolean& a, const Boolean& b) {  // boolean.rs:472-491
        if (a.kind == Constant && !a.c) return b;
        if (b.kind == Constant && !b.c) return a;
        if (a.kind == Constant && a.c) return b.not_();
        if (b.kind == Constant && b.c) return a.not_();
        if (a.kind != b.kind) {
            const Boolean& is = a.kind == Is ? a : b;
            const Boolean& nt = a.kind == Is ? b : a;
            return xor_(cs, is, nt.not_()).not_();
        }
        return from(AllocatedBit::xor_(cs, a.bit, b.bit));
    }

    template <class CS> static Boolean and_(CS&& cs, const Boolean& a, const Boolean& b) {  // boolean.rs:494-516
        if ((a.kind == Constant && !a.c) || (b.kind == Constant && !b.c)) return constant(false);
        if (a.kind == Constant) return b;
        if (b.kind == Constant) return a;
        if (a.kind == Is && b.kind == Not) return from(AllocatedBit::and_not(cs, a.bit, b.bit));
        if (a.kind == Not && b.kind == Is) return from(AllocatedBit::and_not(cs, b.bit, a.bit));
        if (a.kind == Not) return from(AllocatedBit::nor(cs, a.bit, b.bit));
        return from(AllocatedBit::and_(cs, a.bit, b.bit));
    }

    // boolean.rs:519-533: a or b = not(and(not a, not b)), in the namespace "not and (not a) (not b)"
    template <class CS> static Boolean or_(CS&& cs, const Boolean& a, const Boolean& b) {
        auto ns = cs.ns([] { return std::string("not and (not a) (not b)"); });
        return and_(ns, a.not_(), b.not_()).not_();
    }

    // boolean.rs:383-427: one row `0 * 0 = ...` (two constants: nothing, or Unsatisfiable)
    template <class CS> static void enforce_equal(CS&& cs, const Boolean& a, const Boolean& b) {
        if (a.kind == Constant && b.kind == Constant) {
            if (a.c != b.c) throw SynthesisError::unsatisfiable();
            return;
        }
        const Field* f = cs.field();
        auto zero = [](LinearCombination lc) { return lc; };
        if ((a.kind == Constant && a.c) || (b.kind == Constant && b.c)) {
            const Boolean& x = (a.kind == Constant && a.c) ? b : a;
            cs.enforce([] { return std::string("enforce equal to one"); }, zero, zero,
                       [&](LinearCombination lc) { return std::move(lc) + one_var() - x.lc(f, one_var(), Fr::one()); });
            return;
        }
        if (a.kind == Constant || b.kind == Constant) {  // a false constant
            const Boolean& x = a.kind == Constant ? b : a;
            cs.enforce([] { return std::string("enforce equal to zero"); }, zero, zero,
                       [&](LinearCombination) { return x.lc(f, one_var(), Fr::one()); });
            return;
        }
        cs.enforce([] { return std::string("enforce equal"); }, zero, zero,
                   [&](LinearCombination) { return a.lc(f, one_var(), Fr::one()) - b.lc(f, one_var(), Fr::one()); });
    }

    template <class CS> static Boolean sha256_ch(CS&& cs, const Boolean& a, const Boolean& b, const Boolean& c) {  // boolean.rs:536-641
        const OptBool va = a.get_value(), vb = b.get_value(), vc = c.get_value();
        const OptBool chv = (va < 0 || vb < 0 || vc < 0) ? kNone : (OptBool)((va & vb) ^ ((va ^ 1) & vc));
        if (a.kind == Constant && b.kind == Constant && c.kind == Constant) return constant(chv != 0);
        if (a.kind == Constant && !a.c) return c;
        if (b.kind == Constant && !b.c) return and_(cs, a.not_(), c);
        if (c.kind == Constant && !c.c) return and_(cs, a, b);
        if (c.kind == Constant && c.c) return and_(cs, a, b.not_()).not_();
        if (b.kind == Constant && b.c) return and_(cs, a.not_(), c.not_()).not_();
        const Field* f = cs.field();
        const Variable ch = cs.alloc([] { return std::string("ch"); }, [&] { return bit_value(chv); });
        if (g_tape) g_tape->on_op(kTapeCh, ch, a.tape_operand(), b.tape_operand(), c.tape_operand());
        cs.enforce([] { return std::string("ch computation"); },
                   [&](LinearCombination) { return b.lc(f, one_var(), Fr::one()) - c.lc(f, one_var(), Fr::one()); },
                   [&](LinearCombination) { return a.lc(f, one_var(), Fr::one()); },
                   [&](LinearCombination lc) { return std::move(lc) + ch - c.lc(f, one_var(), Fr::one()); });
        return from(AllocatedBit{ch, chv});
    }

    template <class CS> static Boolean sha256_maj(CS&& cs, const Boolean& a, const Boolean& b, const Boolean& c) {  // boolean.rs:644-759
        const OptBool va = a.get_value(), vb = b.get_value(), vc = c.get_value();
        const OptBool mv = (va < 0 || vb < 0 || vc < 0) ? kNone : (OptBool)((va & vb) ^ (va & vc) ^ (vb & vc));
        if (a.kind == Constant && b.kind == Constant && c.kind == Constant) return constant(mv != 0);
        if (a.kind == Constant && !a.c) return and_(cs, b, c);
        if (b.kind == Constant && !b.c) return and_(cs, a, c);
        if (c.kind == Constant && !c.c) return and_(cs, a, b);
        if (c.kind == Constant && c.c) return and_(cs, a.not_(), b.not_()).not_();
        if (b.kind == Constant && b.c) return and_(cs, a.not_(), c.not_()).not_();
        if (a.kind == Constant && a.c) return and_(cs, b.not_(), c.not_()).not_();
        const Field* f = cs.field();
        const Variable maj = cs.alloc([] { return std::string("maj"); }, [&] { return bit_value(mv); });
        if (g_tape) g_tape->on_op(kTapeMaj, maj, a.tape_operand(), b.tape_operand(), c.tape_operand());
        Boolean bc = constant(false);
        {
            auto ns = cs.ns([] { return std::string("b and c"); });
            bc = and_(ns, b, c);
        }
        cs.enforce([] { return std::string("maj computation"); },
                   [&](LinearCombination) {
                       return bc.lc(f, one_var(), Fr::one()) + bc.lc(f, one_var(), Fr::one()) - b.lc(f, one_var(), Fr::one()) -
                              c.lc(f, one_var(), Fr::one());
                   },
                   [&](LinearCombination) { return a.lc(f, one_var(), Fr::one()); },
                   [&](LinearCombination) { return bc.lc(f, one_var(), Fr::one()) - maj; });
        return from(AllocatedBit{maj, mv});
    }
};

// ---- field element -> bits (boolean.rs:320-366) and AllocatedNum (gadgets/num.rs:10-460) ---------------------------------
constexpr unsigned kNumBits = 255;  // ff::PrimeField::NUM_BITS (all three fields)

// boolean.rs:320-366: NUM_BITS freshly allocated bits "bit i", little-endian, of `value` (nullptr = no assignment).
template <class CS> std::vector<AllocatedBit> field_into_allocated_bits_le(CS&& cs, const Fr* value) {
    std::vector<AllocatedBit> bits;
    bits.reserve(kNumBits);
    for (unsigned i = 0; i < kNumBits; ++i) {
        auto ns = cs.ns([&] { return "bit " + std::to_string(i); });
        bits.push_back(AllocatedBit::alloc(ns, value ? (OptBool)Field::bit(*value, i) : kNone));
    }
    return bits;
}

// boolean.rs:306-318
template <class CS> std::vector<Boolean> field_into_boolean_vec_le(CS&& cs, const Fr* value) {
    std::vector<Boolean> out;
    for (const AllocatedBit& b : field_into_allocated_bits_le(cs, value)) out.push_back(Boolean::from(b));
    return out;
}

// boolean.rs:274-304: 64 freshly allocated bits "bit i", little-endian, of `value` (nullptr = no assignment)
template <class CS> std::vector<Boolean> u64_into_boolean_vec_le(CS&& cs, const uint64_t* value) {
    std::vector<Boolean> bits;
    bits.reserve(64);
    for (unsigned i = 0; i < 64; ++i) {
        auto ns = cs.ns([&] { return "bit " + std::to_string(i); });
        bits.push_back(Boolean::from(AllocatedBit::alloc(ns, value ? (OptBool)((*value >> i) & 1) : kNone)));
    }
    return bits;
}

struct AllocatedNum {
    std::optional<Fr> value;
    Variable variable;

    template <class CS, class V> static AllocatedNum alloc(CS&& cs, V&& value_fn) {  // num.rs:27-47
        std::optional<Fr> nv;
        const Variable var = cs.alloc([] { return std::string("num"); }, [&] {
            const Fr t = value_fn();
            nv = t;
            return t;
        });
        if (g_tape) g_tape->fail("AllocatedNum is not a recorded gadget");
        return AllocatedNum{nv, var};
    }
    template <class CS, class V> static AllocatedNum alloc_input(CS&& cs, V&& value_fn) {  // num.rs:61-81
        std::optional<Fr> nv;
        const Variable var = cs.alloc_input([] { return std::string("input num"); }, [&] {
            const Fr t = value_fn();
            nv = t;
            return t;
        });
        return AllocatedNum{nv, var};
    }
    const Fr& need() const {
        if (!value) throw SynthesisError::assignment_missing();
        return *value;
    }
    template <class CS> void inputize(CS&& cs) const {  // num.rs:104-121
        const Variable input = cs.alloc_input([] { return std::string("input variable"); }, [&] { return need(); });
        cs.enforce([] { return std::string("enforce input is correct"); },
                   [&](LinearCombination lc) { return std::move(lc) + input; },
                   [&](LinearCombination lc) { return std::move(lc) + one_var(); },
                   [&](LinearCombination lc) { return std::move(lc) + variable; });
    }

    // num.rs:263-274: NUM_BITS bits and ONE 256-term row  0 * 0 = sum 2^i bit_i - self  (the repo's fattest LCs)
    template <class CS> std::vector<Boolean> to_bits_le(CS&& cs) const {
        const std::vector<AllocatedBit> bits = field_into_allocated_bits_le(cs, value ? &*value : nullptr);
        unpacking_constraint(cs, bits, /*reversed=*/false);
        std::vector<Boolean> out;
        for (auto& b : bits) out.push_back(Boolean::from(b));
        return out;
    }

    // num.rs:128-247: the same with the representation forced below the modulus.  Walks the bits of p - 1 from the top: a bit
    // in a run of ones is a plain allocated bit; at a zero bit of p - 1 the bit must be false when every bit of all the runs
    // of ones so far was set (k-ary AND chained through `last_run`).
    template <class CS> std::vector<Boolean> to_bits_le_strict(CS&& cs) const {
        const Field* f = cs.field();
        const Fr b = f->neg(Fr::one());  // p - 1
        std::vector<AllocatedBit> result;  // big-endian
        std::optional<AllocatedBit> last_run;
        std::vector<AllocatedBit> current_run;
        bool found_one = false;
        unsigned i = 0;
        for (int pos = 255; pos >= 0; --pos) {
            const bool bbit = Field::bit(b, (unsigned)pos);
            const OptBool abit = value ? (OptBool)Field::bit(*value, (unsigned)pos) : kNone;
            found_one |= bbit;
            if (!found_one) {
                if (abit > 0) throw std::logic_error("to_bits_le_strict: value has a bit above the modulus");
                continue;
            }
            if (bbit) {
                auto ns = cs.ns([&] { return "bit " + std::to_string(i); });
                const AllocatedBit a = AllocatedBit::alloc(ns, abit);
                current_run.push_back(a);
                result.push_back(a);
            } else {
                if (!current_run.empty()) {
                    if (last_run) current_run.push_back(*last_run);
                    auto ns = cs.ns([&] { return "run ending at " + std::to_string(i); });
                    AllocatedBit cur = current_run[0];  // kary_and (num.rs:133-160)
                    for (size_t k = 1; k < current_run.size(); ++k) {
                        auto n2 = ns.ns([&] { return "and " + std::to_string(k); });
                        cur = AllocatedBit::and_(n2, cur, current_run[k]);
                    }
                    last_run = cur;
                    current_run.clear();
                }
                auto ns = cs.ns([&] { return "bit " + std::to_string(i); });
                result.push_back(AllocatedBit::alloc_conditionally(ns, abit, *last_run));
            }
            ++i;
        }
        if (!current_run.empty()) throw std::logic_error("to_bits_le_strict: the modulus is odd, so p - 1 ends in a zero bit");
        unpacking_constraint(cs, result, /*reversed=*/true);
        std::vector<Boolean> out;
        for (auto it = result.rbegin(); it != result.rend(); ++it) out.push_back(Boolean::from(*it));
        return out;
    }

    template <class CS> AllocatedNum add(CS&& cs, const AllocatedNum& o) const {  // num.rs:276-306
        const Field* f = cs.field();
        std::optional<Fr> v;
        const Variable var = cs.alloc([] { return std::string("sum num"); }, [&] {
            v = f->add(need(), o.need());
            return *v;
        });
        cs.enforce([] { return std::string("addition constraint"); },
                   [&](LinearCombination lc) { return std::move(lc) + variable + o.variable; },
                   [&](LinearCombination lc) { return std::move(lc) + one_var(); },
                   [&](LinearCombination lc) { return std::move(lc) + var; });
        return AllocatedNum{v, var};
    }
    template <class CS> AllocatedNum mul(CS&& cs, const AllocatedNum& o) const {  // num.rs:308-338
        const Field* f = cs.field();
        std::optional<Fr> v;
        const Variable var = cs.alloc([] { return std::string("product num"); }, [&] {
            v = f->mul(need(), o.need());
            return *v;
        });
        cs.enforce([] { return std::string("multiplication constraint"); },
                   [&](LinearCombination lc) { return std::move(lc) + variable; },
                   [&](LinearCombination lc) { return std::move(lc) + o.variable; },
                   [&](LinearCombination lc) { return std::move(lc) + var; });
        return AllocatedNum{v, var};
    }
    template <class CS> AllocatedNum square(CS&& cs) const {  // num.rs:340-370
        const Field* f = cs.field();
        std::optional<Fr> v;
        const Variable var = cs.alloc([] { return std::string("squared num"); }, [&] {
            v = f->square(need());
            return *v;
        });
        cs.enforce([] { return std::string("squaring constraint"); },
                   [&](LinearCombination lc) { return std::move(lc) + variable; },
                   [&](LinearCombination lc) { return std::move(lc) + variable; },
                   [&](LinearCombination lc) { return std::move(lc) + var; });
        return AllocatedNum{v, var};
    }
    template <class CS> void assert_nonzero(CS&& cs) const {  // num.rs:372-401
        const Field* f = cs.field();
        const Variable inv = cs.alloc([] { return std::string("ephemeral inverse"); }, [&] {
            const Fr& t = need();
            if (t.is_zero()) throw SynthesisError::division_by_zero();
            return f->invert(t);
        });
        cs.enforce([] { return std::string("nonzero assertion constraint"); },
                   [&](LinearCombination lc) { return std::move(lc) + variable; },
                   [&](LinearCombination lc) { return std::move(lc) + inv; },
                   [&](LinearCombination lc) { return std::move(lc) + one_var(); });
    }
    // num.rs:403-455: (b, a) when the condition holds, else (a, b)
    template <class CS> static std::pair<AllocatedNum, AllocatedNum> conditionally_reverse(CS&& cs, const AllocatedNum& a, const AllocatedNum& b,
                                                                                         const Boolean& condition) {
        const Field* f = cs.field();
        auto pick = [&](bool swap_first) {
            const OptBool cv = condition.get_value();
            if (cv < 0) throw SynthesisError::assignment_missing();
            return (cv != 0) == swap_first ? b.need() : a.need();
        };
        AllocatedNum c{std::nullopt, Variable{0}}, d{std::nullopt, Variable{0}};
        {
            auto ns = cs.ns([] { return std::string("conditional reversal result 1"); });
            c = alloc(ns, [&] { return pick(true); });
        }
        cs.enforce([] { return std::string("first conditional reversal"); },
                   [&](LinearCombination lc) { return std::move(lc) + a.variable - b.variable; },
                   [&](LinearCombination) { return condition.lc(f, one_var(), Fr::one()); },
                   [&](LinearCombination lc) { return std::move(lc) + a.variable - c.variable; });
        {
            auto ns = cs.ns([] { return std::string("conditional reversal result 2"); });
            d = alloc(ns, [&] { return pick(false); });
        }
        cs.enforce([] { return std::string("second conditional reversal"); },
                   [&](LinearCombination lc) { return std::move(lc) + b.variable - a.variable; },
                   [&](LinearCombination) { return condition.lc(f, one_var(), Fr::one()); },
                   [&](LinearCombination lc) { return std::move(lc) + b.variable - d.variable; });
        return {c, d};
    }

  private:
    // 0 * 0 = sum_i 2^i bit_i - self   (num.rs:236-247, 263-274); `bits` little-endian, or big-endian when reversed
    template <class CS> void unpacking_constraint(CS&& cs, const std::vector<AllocatedBit>& bits, bool reversed) const {
        const Field* f = cs.field();
        cs.enforce([] { return std::string("unpacking constraint"); },
                   [&](LinearCombination lc) { return lc; },
                   [&](LinearCombination lc) { return lc; },
                   [&](LinearCombination) {
                       LinearCombination lc(f);
                       Fr coeff = Fr::one();
                       for (size_t k = 0; k < bits.size(); ++k) {
                           const AllocatedBit& bt = reversed ? bits[bits.size() - 1 - k] : bits[k];
                           lc.add_term(bt.variable, coeff);
                           coeff = f->dbl(coeff);
                       }
                       return std::move(lc) - variable;
                   });
    }
};

// ---- MultiEq (multieq.rs:6-122) ------------------------------------------------------------------------------
template <class CS> class MultiEq {
  public:
    using Root = MultiEq;
    explicit MultiEq(CS& cs) : cs_(cs), lhs_(cs.field()), rhs_(cs.field()) {}
    ~MultiEq() noexcept(false) {  // Drop (multieq.rs:61-67)
        if (bits_used_ > 0) accumulate();
    }
    static Variable one() { return one_var(); }
    const Field* field() { return cs_.field(); }
    template <class N, class V> Variable alloc(N&& n, V&& v) { return cs_.alloc(n, v); }
    template <class N, class V> Variable alloc_input(N&& n, V&& v) { return cs_.alloc_input(n, v); }
    template <class N, class A, class B, class C> void enforce(N&& n, A&& a, B&& b, C&& c) { cs_.enforce(n, a, b, c); }
    template <class N> void push_namespace(N&& n) { cs_.get_root().push_namespace(n); }
    void pop_namespace() { cs_.get_root().pop_namespace(); }
    Root& get_root() { return *this; }
    template <class N> Namespace<Root> ns(N&& n) {
        push_namespace(n);
        return Namespace<Root>(*this);
    }

    void enforce_equal(unsigned num_bits, const LinearCombination& lhs, const LinearCombination& rhs) {  // multieq.rs:41-58
        if (kCapacity <= bits_used_ + num_bits) accumulate();
        if (!(kCapacity > bits_used_ + num_bits)) throw std::logic_error("MultiEq: equality wider than the field capacity");
        // coeff = 2^bits_used (multieq.rs:54); (coeff, &lc) scales every coefficient (lc.rs:339-356)
        lhs_.add_scaled_pow2(bits_used_, lhs);
        rhs_.add_scaled_pow2(bits_used_, rhs);
        bits_used_ += num_bits;
    }

  private:
    void accumulate() {  // multieq.rs:25-39
        const unsigned ops = ops_;
        cs_.enforce([&] { return "multieq " + std::to_string(ops); },
                    [&](LinearCombination) { return std::move(lhs_); },
                    [&](LinearCombination lc) { return std::move(lc) + one_var(); },
                    [&](LinearCombination) { return std::move(rhs_); });
        lhs_ = LinearCombination(cs_.field());
        rhs_ = LinearCombination(cs_.field());
        bits_used_ = 0;
        ops_ += 1;
    }
    CS& cs_;
    unsigned ops_ = 0, bits_used_ = 0;
    LinearCombination lhs_, rhs_;
};

// ---- UInt32 (uint32.rs) -----------------------------------------------------------------------------------------
struct UInt32 {
    std::array<Boolean, 32> bits;  // least significant first
    std::optional<uint32_t> value;

    static UInt32 constant(uint32_t v) {
        UInt32 u{{}, v};
        for (int i = 0; i < 32; ++i) u.bits[i] = Boolean::constant((v >> i) & 1);
        return u;
    }
    template <class CS> static UInt32 alloc(CS&& cs, std::optional<uint32_t> value) {  // uint32.rs:43-72
        UInt32 u{{}, value};
        for (int i = 0; i < 32; ++i) {
            auto ns = cs.ns([&] { return "allocated bit " + std::to_string(i); });
            u.bits[i] = Boolean::from(AllocatedBit::alloc(ns, value ? (OptBool)((*value >> i) & 1) : kNone));
        }
        return u;
    }
    static std::optional<uint32_t> value_of(const std::array<Boolean, 32>& le) {
        uint32_t v = 0;
        for (int i = 0; i < 32; ++i) {
            const OptBool b = le[i].get_value();
            if (b < 0) return std::nullopt;
            v |= (uint32_t)b << i;
        }
        return v;
    }
    static UInt32 from_bits_be(const Boolean* be) {  // uint32.rs:80-109
        UInt32 u{{}, std::nullopt};
        for (int i = 0; i < 32; ++i) u.bits[i] = be[31 - i];
        u.value = value_of(u.bits);
        return u;
    }
    static UInt32 from_bits(const Boolean* le) {  // uint32.rs:118-163
        UInt32 u{{}, std::nullopt};
        for (int i = 0; i < 32; ++i) u.bits[i] = le[i];
        u.value = value_of(u.bits);
        return u;
    }
    void into_bits_be(std::vector<Boolean>& out) const {
        for (int i = 31; i >= 0; --i) out.push_back(bits[i]);
    }
    void into_bits(std::vector<Boolean>& out) const {
        for (int i = 0; i < 32; ++i) out.push_back(bits[i]);
    }
    UInt32 rotr(unsigned by) const {  // uint32.rs:165-181
        by %= 32;
        UInt32 u{{}, std::nullopt};
        for (unsigned i = 0; i < 32; ++i) u.bits[i] = bits[(i + by) % 32];
        if (value) u.value = by ? ((*value >> by) | (*value << (32 - by))) : *value;
        return u;
    }
    UInt32 shr(unsigned by) const {  // uint32.rs:183-201
        by %= 32;
        UInt32 u{{}, std::nullopt};
        for (unsigned i = 0; i < 32; ++i) u.bits[i] = (i + by < 32) ? bits[i + by] : Boolean::constant(false);
        if (value) u.value = *value >> by;
        return u;
    }
    template <class CS> UInt32 xor_(CS&& cs, const UInt32& o) const {  // uint32.rs:281-303
        UInt32 u{{}, std::nullopt};
        if (value && o.value) u.value = *value ^ *o.value;
        for (int i = 0; i < 32; ++i) {
            auto ns = cs.ns([&] { return "xor of bit " + std::to_string(i); });
            u.bits[i] = Boolean::xor_(ns, bits[i], o.bits[i]);
        }
        return u;
    }
    template <class CS> static UInt32 sha256_maj(CS&& cs, const UInt32& a, const UInt32& b, const UInt32& c) {  // uint32.rs:238-257
        UInt32 u{{}, std::nullopt};
        if (a.value && b.value && c.value) u.value = (*a.value & *b.value) ^ (*a.value & *c.value) ^ (*b.value & *c.value);
        for (int i = 0; i < 32; ++i) {
            auto ns = cs.ns([&] { return "maj " + std::to_string(i); });
            u.bits[i] = Boolean::sha256_maj(ns, a.bits[i], b.bits[i], c.bits[i]);
        }
        return u;
    }
    template <class CS> static UInt32 sha256_ch(CS&& cs, const UInt32& a, const UInt32& b, const UInt32& c) {  // uint32.rs:259-278
        UInt32 u{{}, std::nullopt};
        if (a.value && b.value && c.value) u.value = (*a.value & *b.value) ^ ((~*a.value) & *c.value);
        for (int i = 0; i < 32; ++i) {
            auto ns = cs.ns([&] { return "ch " + std::to_string(i); });
            u.bits[i] = Boolean::sha256_ch(ns, a.bits[i], b.bits[i], c.bits[i]);
        }
        return u;
    }

    // uint32.rs:306-406.  M is a constraint system whose Root is a MultiEq.
    template <class M> static UInt32 addmany(M&& cs, const UInt32* operands, size_t n_ops) {
        if (n_ops < 2 || n_ops > 10) throw std::logic_error("addmany: 2..10 operands");
        const Field* f = cs.field();
        uint64_t max_value = (uint64_t)n_ops * 0xffffffffull;
        std::optional<uint64_t> result_value = 0;
        LinearCombination lc(f);
        std::vector<Term> scratch;
        scratch.reserve(n_ops * 64);
        bool all_constants = true;
        for (size_t k = 0; k < n_ops; ++k) {
            const UInt32& op = operands[k];
            if (!op.value) result_value = std::nullopt;
            else if (result_value) *result_value += *op.value;
            Fr coeff = Fr::one();
            for (int i = 0; i < 32; ++i) {
                op.bits[i].terms_into(scratch, f, one_var(), coeff);
                all_constants &= op.bits[i].is_constant();
                coeff = f->dbl(coeff);
            }
        }
        std::optional<uint32_t> modular_value;
        if (result_value) modular_value = (uint32_t)*result_value;
        if (all_constants && modular_value) return constant(*modular_value);
        lc.add_terms_bulk(scratch);  // == the reference's per-bit `lc = lc + &bit.lc(one, coeff)` (uint32.rs:349-355)
        UInt32 out{{}, modular_value};
        LinearCombination result_lc(f);
        Fr coeff = Fr::one();
        unsigned i = 0;
        if (g_tape) {  // the result bits allocated below are the bits of this integer sum
            g_tape->begin_sum();
            for (size_t k = 0; k < n_ops; ++k)
                for (unsigned bi = 0; bi < 32; ++bi) g_tape->sum_operand(operands[k].bits[bi].tape_operand(), bi);
        }
        while (max_value != 0) {
            AllocatedBit b{Variable{0}, kNone};
            {
                auto ns = cs.ns([&] { return "result bit " + std::to_string(i); });
                b = AllocatedBit::alloc(ns, result_value ? (OptBool)((*result_value >> i) & 1) : kNone);
            }
            result_lc.add_term(b.variable, coeff);
            if (i < 32) out.bits[i] = Boolean::from(b);
            max_value >>= 1;
            ++i;
            coeff = f->dbl(coeff);
        }
        if (g_tape) g_tape->end_sum();
        cs.get_root().enforce_equal(i, lc, result_lc);
        return out;
    }
    template <class M> static UInt32 addmany(M&& cs, const std::vector<UInt32>& ops) { return addmany(cs, ops.data(), ops.size()); }
};

// ---- sha256 (sha256.rs) -------------------------------------------------------------------------------------------
namespace sha256_detail {
static const uint32_t K[64] = {
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01, 0x243185be,
    0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa,
    0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967, 0x27b70a85,
    0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85, 0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3,
    0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f,
    0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};
static const uint32_t IV[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};

struct Maybe {  // sha256.rs:128-148
    bool deferred;
    UInt32 concrete;
    std::vector<UInt32> ops;
    template <class M> UInt32 compute(M&& cs, const UInt32* others, size_t n_others) const {
        if (!deferred) return concrete;
        std::vector<UInt32> v = ops;
        v.insert(v.end(), others, others + n_others);
        return UInt32::addmany(cs, v);
    }
};
}  // namespace sha256_detail

inline std::vector<UInt32> sha256_iv() {
    std::vector<UInt32> v;
    for (int i = 0; i < 8; ++i) v.push_back(UInt32::constant(sha256_detail::IV[i]));
    return v;
}

// sha256.rs:83-272
template <class CS> std::vector<UInt32> sha256_compression_function(CS&& cs_in, const Boolean* input /*512*/, const std::vector<UInt32>& cur) {
    using namespace sha256_detail;
    using CSv = std::remove_reference_t<CS>;
    std::vector<UInt32> w;
    w.reserve(64);
    for (int i = 0; i < 16; ++i) w.push_back(UInt32::from_bits_be(input + 32 * i));
    std::vector<UInt32> out;
    {
        MultiEq<CSv> cs(cs_in);
        for (int i = 16; i < 64; ++i) {
            auto ns = cs.ns([&] { return "w extension " + std::to_string(i); });
            UInt32 s0 = w[i - 15].rotr(7);
            {
                auto n2 = ns.ns([] { return std::string("first xor for s0"); });
                s0 = s0.xor_(n2, w[i - 15].rotr(18));
            }
            {
                auto n2 = ns.ns([] { return std::string("second xor for s0"); });
                s0 = s0.xor_(n2, w[i - 15].shr(3));
            }
            UInt32 s1 = w[i - 2].rotr(17);
            {
                auto n2 = ns.ns([] { return std::string("first xor for s1"); });
                s1 = s1.xor_(n2, w[i - 2].rotr(19));
            }
            {
                auto n2 = ns.ns([] { return std::string("second xor for s1"); });
                s1 = s1.xor_(n2, w[i - 2].shr(10));
            }
            auto n2 = ns.ns([] { return std::string("computation of w[i]"); });
            const UInt32 ops[4] = {w[i - 16], s0, w[i - 7], s1};
            w.push_back(UInt32::addmany(n2, ops, 4));
        }
        Maybe a{false, cur[0], {}}, e{false, cur[4], {}};
        UInt32 b = cur[1], c = cur[2], d = cur[3], f = cur[5], g = cur[6], h = cur[7];
        for (int i = 0; i < 64; ++i) {
            auto ns = cs.ns([&] { return "compression round " + std::to_string(i); });
            UInt32 new_e = [&] {
                auto n2 = ns.ns([] { return std::string("deferred e computation"); });
                return e.compute(n2, nullptr, 0);
            }();
            UInt32 s1 = new_e.rotr(6);
            {
                auto n2 = ns.ns([] { return std::string("first xor for s1"); });
                s1 = s1.xor_(n2, new_e.rotr(11));
            }
            {
                auto n2 = ns.ns([] { return std::string("second xor for s1"); });
                s1 = s1.xor_(n2, new_e.rotr(25));
            }
            UInt32 ch = [&] {
                auto n2 = ns.ns([] { return std::string("ch"); });
                return UInt32::sha256_ch(n2, new_e, f, g);
            }();
            std::vector<UInt32> temp1 = {h, s1, ch, UInt32::constant(K[i]), w[i]};
            UInt32 new_a = [&] {
                auto n2 = ns.ns([] { return std::string("deferred a computation"); });
                return a.compute(n2, nullptr, 0);
            }();
            UInt32 s0 = new_a.rotr(2);
            {
                auto n2 = ns.ns([] { return std::string("first xor for s0"); });
                s0 = s0.xor_(n2, new_a.rotr(13));
            }
            {
                auto n2 = ns.ns([] { return std::string("second xor for s0"); });
                s0 = s0.xor_(n2, new_a.rotr(22));
            }
            UInt32 maj = [&] {
                auto n2 = ns.ns([] { return std::string("maj"); });
                return UInt32::sha256_maj(n2, new_a, b, c);
            }();
            h = g;
            g = f;
            f = new_e;
            e = Maybe{true, UInt32{}, temp1};
            e.ops.push_back(d);
            d = c;
            c = b;
            b = new_a;
            a = Maybe{true, UInt32{}, temp1};
            a.ops.push_back(s0);
            a.ops.push_back(maj);
        }
        auto add2 = [&](const char* name, const UInt32& x, const UInt32& y) {
            auto ns = cs.ns([&] { return std::string(name); });
            const UInt32 ops[2] = {x, y};
            return UInt32::addmany(ns, ops, 2);
        };
        UInt32 h0 = [&] {
            auto ns = cs.ns([] { return std::string("deferred h0 computation"); });
            return a.compute(ns, &cur[0], 1);
        }();
        UInt32 h1 = add2("new h1", cur[1], b);
        UInt32 h2 = add2("new h2", cur[2], c);
        UInt32 h3 = add2("new h3", cur[3], d);
        UInt32 h4 = [&] {
            auto ns = cs.ns([] { return std::string("deferred h4 computation"); });
            return e.compute(ns, &cur[4], 1);
        }();
        UInt32 h5 = add2("new h5", cur[5], f);
        UInt32 h6 = add2("new h6", cur[6], g);
        UInt32 h7 = add2("new h7", cur[7], h);
        out = {h0, h1, h2, h3, h4, h5, h6, h7};
    }  // MultiEq drops here: flushes the last packed equality
    return out;
}

// sha256.rs:50-77
template <class CS> std::vector<Boolean> sha256(CS&& cs, const std::vector<Boolean>& input) {
    if (input.size() % 8 != 0) throw std::logic_error("sha256: input must be whole bytes");
    std::vector<Boolean> padded = input;
    const uint64_t plen = padded.size();
    padded.push_back(Boolean::constant(true));
    while ((padded.size() + 64) % 512 != 0) padded.push_back(Boolean::constant(false));
    for (int i = 63; i >= 0; --i) padded.push_back(Boolean::constant((plen >> i) & 1));
    std::vector<UInt32> cur = sha256_iv();
    for (size_t blk = 0; blk < padded.size() / 512; ++blk) {
        auto ns = cs.ns([&] { return "block " + std::to_string(blk); });
        cur = sha256_compression_function(ns, padded.data() + 512 * blk, cur);
    }
    std::vector<Boolean> out;
    for (auto& wd : cur) wd.into_bits_be(out);
    return out;
}

// ---- blake2s (crates/bellpepper/src/gadgets/blake2s.rs) -----------------------------------------------------------------
namespace blake2s_detail {
static const unsigned R1 = 16, R2 = 12, R3 = 8, R4 = 7;  // blake2s.rs:29-32
static const uint8_t SIGMA[10][16] = {  // blake2s.rs:50-61
    {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3},
    {11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4}, {7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8},
    {9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13}, {2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9},
    {12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11}, {13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10},
    {6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5}, {10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0}};
static const uint8_t G_ARGS[8][4] = {{0, 4, 8, 12}, {1, 5, 9, 13}, {2, 6, 10, 14}, {3, 7, 11, 15},
                                     {0, 5, 10, 15}, {1, 6, 11, 12}, {2, 7, 8, 13}, {3, 4, 9, 14}};  // blake2s.rs:226-305
static const uint32_t IV[8] = {0x6A09E667, 0xBB67AE85, 0x3C6EF372, 0xA54FF53A, 0x510E527F, 0x9B05688C, 0x1F83D9AB, 0x5BE0CD19};

// blake2s.rs:86-121 (the root of `cs` is a MultiEq)
template <class M> void mixing_g(M&& cs, std::vector<UInt32>& v, unsigned a, unsigned b, unsigned c, unsigned d, const UInt32& x, const UInt32& y) {
    auto step = [&](int i) { return cs.ns([i] { return "mixing step " + std::to_string(i); }); };
    { auto ns = step(1); const UInt32 ops[3] = {v[a], v[b], x}; v[a] = UInt32::addmany(ns, ops, 3); }
    { auto ns = step(2); v[d] = v[d].xor_(ns, v[a]).rotr(R1); }
    { auto ns = step(3); const UInt32 ops[2] = {v[c], v[d]}; v[c] = UInt32::addmany(ns, ops, 2); }
    { auto ns = step(4); v[b] = v[b].xor_(ns, v[c]).rotr(R2); }
    { auto ns = step(5); const UInt32 ops[3] = {v[a], v[b], y}; v[a] = UInt32::addmany(ns, ops, 3); }
    { auto ns = step(6); v[d] = v[d].xor_(ns, v[a]).rotr(R3); }
    { auto ns = step(7); const UInt32 ops[2] = {v[c], v[d]}; v[c] = UInt32::addmany(ns, ops, 2); }
    { auto ns = step(8); v[b] = v[b].xor_(ns, v[c]).rotr(R4); }
}
}  // namespace blake2s_detail

// blake2s.rs:171-315: h is updated in place
template <class CS> void blake2s_compression(CS&& cs_in, std::vector<UInt32>& h, const std::vector<UInt32>& m, uint64_t t, bool f) {
    using namespace blake2s_detail;
    using CSv = std::remove_reference_t<CS>;
    std::vector<UInt32> v(h);
    for (int i = 0; i < 8; ++i) v.push_back(UInt32::constant(IV[i]));
    { auto ns = cs_in.ns([] { return std::string("first xor"); }); v[12] = v[12].xor_(ns, UInt32::constant((uint32_t)t)); }
    { auto ns = cs_in.ns([] { return std::string("second xor"); }); v[13] = v[13].xor_(ns, UInt32::constant((uint32_t)(t >> 32))); }
    if (f) { auto ns = cs_in.ns([] { return std::string("third xor"); }); v[14] = v[14].xor_(ns, UInt32::constant(0xffffffffu)); }
    {
        MultiEq<CSv> cs(cs_in);
        for (int i = 0; i < 10; ++i) {
            auto rns = cs.ns([i] { return "round " + std::to_string(i); });
            const uint8_t* s = SIGMA[i % 10];
            for (int j = 0; j < 8; ++j) {
                auto ns = rns.ns([j] { return "mixing invocation " + std::to_string(j + 1); });
                mixing_g(ns, v, G_ARGS[j][0], G_ARGS[j][1], G_ARGS[j][2], G_ARGS[j][3], m[s[2 * j]], m[s[2 * j + 1]]);
            }
        }
    }  // MultiEq drops here
    for (int i = 0; i < 8; ++i) {
        auto ns = cs_in.ns([i] { return "h[" + std::to_string(i) + "] ^ v[" + std::to_string(i) + "] ^ v[" + std::to_string(i) + " + 8]"; });
        { auto n2 = ns.ns([] { return std::string("first xor"); }); h[i] = h[i].xor_(n2, v[i]); }
        { auto n2 = ns.ns([] { return std::string("second xor"); }); h[i] = h[i].xor_(n2, v[i + 8]); }
    }
}

// blake2s.rs:344-406: input bits little-endian per byte; output through into_bits (little-endian)
template <class CS> std::vector<Boolean> blake2s(CS&& cs, const std::vector<Boolean>& input, const uint8_t personalization[8]) {
    using namespace blake2s_detail;
    if (input.size() % 8 != 0) throw std::logic_error("blake2s: input must be whole bytes");
    auto le32 = [](const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); };
    std::vector<UInt32> h;
    for (int i = 0; i < 8; ++i) h.push_back(UInt32::constant(IV[i]));
    h[0] = UInt32::constant(IV[0] ^ 0x01010000u ^ 32u);
    h[6] = UInt32::constant(IV[6] ^ le32(personalization));
    h[7] = UInt32::constant(IV[7] ^ le32(personalization + 4));
    std::vector<std::vector<UInt32>> blocks;
    for (size_t b0 = 0; b0 < input.size(); b0 += 512) {
        std::vector<UInt32> words;
        const size_t b1 = std::min(input.size(), b0 + 512);
        for (size_t w0 = b0; w0 < b1; w0 += 32) {
            Boolean tmp[32];
            for (size_t i = 0; i < 32; ++i) tmp[i] = (w0 + i < b1) ? input[w0 + i] : Boolean::constant(false);
            words.push_back(UInt32::from_bits(tmp));
        }
        while (words.size() < 16) words.push_back(UInt32::constant(0));
        blocks.push_back(std::move(words));
    }
    if (blocks.empty()) blocks.push_back(std::vector<UInt32>(16, UInt32::constant(0)));
    // witness tape (wtape.hpp): a compression is a unit -- its 512-bit message window, and which variables carry the chaining
    // value h (word j, bit i = state bit 32 j + i; a negated Boolean carries the inverted bit)
    auto note_unit = [&](size_t blk) {
        if (!g_tape) return;
        g_tape->begin_unit(512 * (uint64_t)blk);
        for (uint32_t j = 0; j < 8; ++j)
            for (uint32_t i = 0; i < 32; ++i) {
                const Boolean& b = h[j].bits[i];
                if (!b.is_constant()) g_tape->units.back().state[b.bit.variable.index()] = (32 * j + i) | (b.kind == Boolean::Not ? 0x80000000u : 0u);
            }
    };
    for (size_t i = 0; i + 1 < blocks.size(); ++i) {
        note_unit(i);
        auto ns = cs.ns([i] { return "block " + std::to_string(i); });
        blake2s_compression(ns, h, blocks[i], (uint64_t)(i + 1) * 64, false);
    }
    {
        note_unit(blocks.size() - 1);
        auto ns = cs.ns([] { return std::string("final block"); });
        blake2s_compression(ns, h, blocks.back(), input.size() / 8, true);
    }
    std::vector<Boolean> out;
    for (auto& wd : h) wd.into_bits(out);
    return out;
}

}  // namespace bph
