// Witness tape: while a bit-logic gadget circuit (boolean / uint32 / sha256) is synthesized, record HOW every aux variable's
// value follows from earlier ones, so that the next witness of the same circuit can be generated on the device instead of by
// re-running the host closures (the producer it replaces: crates/bellpepper/src/gadgets/sha256.rs:83-272 through
// boolean.rs:68-272, 536-759 and uint32.rs:306-406; the consumer: SizedWitness::generate_witness_into, witness_cs.rs:7-41).
//
// One entry per aux variable, in allocation order:
//   FREE                       the value comes from outside (a message bit)
//   XOR / AND / AND_NOT / NOR  AllocatedBit::{xor, and, and_not, nor}           (boolean.rs:101-271)
//   CH / MAJ                   the bit Boolean::sha256_ch / sha256_maj allocates (boolean.rs:536-759)
//   SUMBIT(s, j)               bit j of the integer sum s of UInt32::addmany's operands (uint32.rs:306-406)
// Operands are Booleans: a constant, a variable, or a negated variable.
//
// build_program() turns the tape of a chained, block-structured circuit into the flat program bp_cs_set_witness_program takes
// (include/bp_r1cs.h): per unit (compression block) a tape with operands made RELATIVE (own variable / message bit of the
// unit / chaining-state bit), entries sorted into dependency levels, identical tapes shared between units.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "lc.hpp"

namespace bph {

enum : uint32_t { kTapeFree = 0, kTapeXor = 1, kTapeAnd = 2, kTapeAndNot = 3, kTapeNor = 4, kTapeCh = 5, kTapeMaj = 6, kTapeSumBit = 7 };
// recorded operand: low 2 bits = kind, the rest = aux index
enum : uint32_t { kOpConst0 = 0, kOpConst1 = 1, kOpVar = 2, kOpNotVar = 3 };
// program operand (device): bits 31..29 = kind, bits 28..0 = payload; sum operands carry their bit position in bits 28..24
enum : uint32_t { kPConst0 = 0, kPConst1 = 1, kPOwn = 2, kPNotOwn = 3, kPMsg = 4, kPNotMsg = 5, kPState = 6, kPNotState = 7 };
constexpr uint32_t kWprogMagic = 0x50575042u;  // "BPWP"
constexpr uint32_t kWprogHeaderWords = 16;

struct TapeEntry {
    uint32_t op, a, b, c;
};
struct TapeSum {
    uint32_t first_op, n_ops;
    uint64_t constant;
};

struct WitnessTape {
    bool ok = true;
    std::string why;
    uint64_t aux_base = 0;  // aux index of entries[0]
    std::vector<TapeEntry> entries;
    std::vector<TapeSum> sums;
    std::vector<uint32_t> sum_ops;     // recorded operand
    std::vector<uint8_t> sum_shift;    // its bit position (weight 2^shift)
    int open_sum = -1;
    uint32_t next_bit = 0;
    // unit (block) structure, noted by the circuit driver
    struct Unit {
        uint64_t first_entry, msg_bit_base;
        std::map<uint32_t, uint32_t> state;  // aux index -> state bit | (inverted << 31): the variable's value is that chaining-state bit
    };
    std::vector<Unit> units;
    uint64_t n_msg_bits = 0;

    void fail(const char* w) {
        if (ok) why = w;
        ok = false;
    }
    static uint32_t operand(const Variable& v, bool negated) {
        return (negated ? kOpNotVar : kOpVar) | ((uint32_t)v.index() << 2);
    }
    void check_next(const Variable& r) {
        if (!r.is_aux() || r.index() != aux_base + entries.size()) fail("an aux variable was allocated outside the recorded gadgets");
    }
    void on_alloc(const Variable& r) {  // AllocatedBit::alloc
        check_next(r);
        if (open_sum >= 0) entries.push_back(TapeEntry{kTapeSumBit, (uint32_t)open_sum, next_bit++, 0});
        else entries.push_back(TapeEntry{kTapeFree, 0, 0, 0});
    }
    void on_op(uint32_t op, const Variable& r, uint32_t a, uint32_t b, uint32_t c = 0) {
        check_next(r);
        if (open_sum >= 0) fail("a logic gate was allocated inside an addmany");
        entries.push_back(TapeEntry{op, a, b, c});
    }
    void begin_sum() {
        if (open_sum >= 0) fail("nested addmany");
        sums.push_back(TapeSum{(uint32_t)sum_ops.size(), 0, 0});
        open_sum = (int)sums.size() - 1;
        next_bit = 0;
    }
    void sum_operand(uint32_t recorded, unsigned shift) {
        TapeSum& s = sums.back();
        const uint32_t kind = recorded & 3u;
        if (kind == kOpConst0) return;
        if (kind == kOpConst1) {
            s.constant += 1ull << shift;
            return;
        }
        sum_ops.push_back(recorded);
        sum_shift.push_back((uint8_t)shift);
        s.n_ops++;
    }
    void end_sum() { open_sum = -1; }
    void begin_unit(uint64_t msg_bit_base) { units.push_back(Unit{entries.size(), msg_bit_base, {}}); }

    // ---- flat program ------------------------------------------------------------------------------------------------
    // entries [0, n_msg_bits) must be FREE (the message bits, in message order); every later entry belongs to a unit.
    std::vector<uint32_t> build_program(uint64_t n_aux_total, bool msb_first) {
        std::vector<uint32_t> blob;
        if (!ok) return blob;
        if (open_sum >= 0 || units.empty() || entries.size() + aux_base != n_aux_total || units[0].first_entry != n_msg_bits) {
            fail("tape does not cover the witness");
            return blob;
        }
        for (uint64_t i = 0; i < n_msg_bits; ++i)
            if (entries[i].op != kTapeFree) {
                fail("message bits are not the first allocations");
                return blob;
            }
        struct Tape {
            uint32_t n_vars;
            std::vector<uint32_t> levels, ents, sums, sumops;  // flat records
        };
        std::vector<Tape> tapes;
        std::map<std::vector<uint32_t>, uint32_t> seen;  // serialised tape -> id
        std::vector<uint32_t> unit_tape(units.size());
        for (size_t u = 0; u < units.size() && ok; ++u) {
            const uint64_t e0 = units[u].first_entry, e1 = u + 1 < units.size() ? units[u + 1].first_entry : entries.size();
            const uint64_t v0 = aux_base + e0;  // aux index of the unit's first variable
            const uint32_t n = (uint32_t)(e1 - e0);
            if (n >= (1u << 24)) { fail("unit too large"); break; }
            std::vector<uint32_t> level(n, 0);
            auto rel = [&](uint32_t recorded, uint32_t* lvl) -> uint32_t {  // recorded operand -> program operand (no shift bits)
                const uint32_t kind = recorded & 3u, idx = recorded >> 2;
                if (kind == kOpConst0) return kPConst0 << 29;
                if (kind == kOpConst1) return kPConst1 << 29;
                const bool neg = kind == kOpNotVar;
                if (idx >= v0 && idx < v0 + n) {
                    *lvl = std::max(*lvl, level[idx - v0]);
                    return ((neg ? kPNotOwn : kPOwn) << 29) | (uint32_t)(idx - v0);
                }
                if (idx >= aux_base && idx < aux_base + n_msg_bits) {
                    const uint64_t bit = idx - aux_base;
                    if (bit < units[u].msg_bit_base || bit - units[u].msg_bit_base >= (1u << 24)) { fail("message bit of another unit"); return 0; }
                    return ((neg ? kPNotMsg : kPMsg) << 29) | (uint32_t)(bit - units[u].msg_bit_base);
                }
                auto it = units[u].state.find(idx);
                if (it == units[u].state.end()) { fail("operand outside unit, message and chaining state"); return 0; }
                const bool inv = ((it->second >> 31) != 0) != neg;
                return ((inv ? kPNotState : kPState) << 29) | (it->second & 0x7fffffffu);
            };
            // levels: a value is available one level after everything it reads; SUMBITs share the level of their sum
            struct E { uint32_t lvl, res_op, a, b, c; };
            std::vector<E> es;
            es.reserve(n);
            struct S { uint32_t lvl, first, nops; uint64_t constant; };
            std::vector<S> ss;
            std::vector<uint32_t> sumops;
            std::map<uint32_t, uint32_t> sum_local;  // global sum id -> id within the unit
            for (uint32_t i = 0; i < n && ok; ++i) {
                const TapeEntry& t = entries[e0 + i];
                uint32_t lvl = 0, a = 0, b = 0, c = 0;
                if (t.op == kTapeFree) { fail("a free variable inside a unit"); break; }
                if (t.op == kTapeSumBit) {
                    auto it = sum_local.find(t.a);
                    if (it == sum_local.end()) {
                        const TapeSum& gs = sums[t.a];
                        S s{0, (uint32_t)sumops.size(), gs.n_ops, gs.constant};
                        for (uint32_t k = 0; k < gs.n_ops; ++k) {
                            uint32_t ol = 0;
                            const uint32_t po = rel(sum_ops[gs.first_op + k], &ol);
                            if ((po & 0x00ffffffu) != (po & 0x1fffffffu)) { fail("sum operand index too large"); break; }
                            s.lvl = std::max(s.lvl, ol);
                            sumops.push_back((po & 0xe0ffffffu) | ((uint32_t)sum_shift[gs.first_op + k] << 24));
                        }
                        s.lvl += 1;
                        it = sum_local.emplace(t.a, (uint32_t)ss.size()).first;
                        ss.push_back(s);
                    }
                    lvl = ss[it->second].lvl;
                    a = it->second;
                    b = t.b;
                    if (b >= 64) { fail("sum bit out of range"); break; }
                } else {
                    a = rel(t.a, &lvl);
                    b = rel(t.b, &lvl);
                    if (t.op == kTapeCh || t.op == kTapeMaj) c = rel(t.c, &lvl);
                    lvl += 1;
                }
                level[i] = lvl;
                es.push_back(E{lvl, i | (t.op << 28), a, b, c});
            }
            if (!ok) break;
            uint32_t n_levels = 0;
            for (auto& e : es) n_levels = std::max(n_levels, e.lvl + 1);
            for (auto& s : ss) n_levels = std::max(n_levels, s.lvl + 1);
            // stable counting sort by level
            std::vector<uint32_t> ecount(n_levels + 1, 0), scount(n_levels + 1, 0);
            for (auto& e : es) ecount[e.lvl + 1]++;
            for (auto& s : ss) scount[s.lvl + 1]++;
            for (uint32_t l = 0; l < n_levels; ++l) { ecount[l + 1] += ecount[l]; scount[l + 1] += scount[l]; }
            Tape tp;
            tp.n_vars = n;
            tp.ents.resize(4 * es.size());
            tp.sums.resize(4 * ss.size());
            std::vector<uint32_t> epos(ecount.begin(), ecount.end() - 1), spos(scount.begin(), scount.end() - 1), sum_new(ss.size());
            for (size_t k = 0; k < ss.size(); ++k) sum_new[k] = spos[ss[k].lvl]++;
            for (size_t k = 0; k < ss.size(); ++k) {
                uint32_t* r = &tp.sums[4 * sum_new[k]];
                r[0] = ss[k].first; r[1] = ss[k].nops; r[2] = (uint32_t)ss[k].constant; r[3] = (uint32_t)(ss[k].constant >> 32);
            }
            for (auto& e : es) {
                uint32_t* r = &tp.ents[4 * epos[e.lvl]++];
                r[0] = e.res_op;
                r[1] = ((e.res_op >> 28) == kTapeSumBit) ? sum_new[e.a] : e.a;
                r[2] = e.b;
                r[3] = e.c;
            }
            tp.levels.resize(4 * n_levels);
            for (uint32_t l = 0; l < n_levels; ++l) {
                tp.levels[4 * l] = ecount[l]; tp.levels[4 * l + 1] = ecount[l + 1];
                tp.levels[4 * l + 2] = scount[l]; tp.levels[4 * l + 3] = scount[l + 1];
            }
            tp.sumops = std::move(sumops);
            // identical tapes are shared
            std::vector<uint32_t> key;
            key.push_back(n);
            for (auto* v : {&tp.levels, &tp.ents, &tp.sums, &tp.sumops}) {
                key.push_back((uint32_t)v->size());
                key.insert(key.end(), v->begin(), v->end());
            }
            auto it = seen.find(key);
            if (it == seen.end()) {
                it = seen.emplace(std::move(key), (uint32_t)tapes.size()).first;
                tapes.push_back(std::move(tp));
            }
            unit_tape[u] = it->second;
        }
        if (!ok) return blob;
        // ---- serialise ----
        const uint32_t n_units = (uint32_t)units.size(), n_tapes = (uint32_t)tapes.size();
        uint32_t max_vars = 0, max_sums = 0;
        for (auto& t : tapes) { max_vars = std::max(max_vars, t.n_vars); max_sums = std::max<uint32_t>(max_sums, (uint32_t)t.sums.size() / 4); }
        blob.assign(kWprogHeaderWords, 0);
        const uint32_t unit_off = kWprogHeaderWords, tape_off = unit_off + 4 * n_units;
        blob.resize(tape_off + 8 * n_tapes, 0);
        for (uint32_t u = 0; u < n_units; ++u) {
            uint32_t* r = &blob[unit_off + 4 * u];
            r[0] = unit_tape[u];
            r[1] = (uint32_t)(aux_base + units[u].first_entry);
            r[2] = (uint32_t)units[u].msg_bit_base;
            r[3] = u;  // chaining-state index
        }
        for (uint32_t t = 0; t < n_tapes; ++t) {
            uint32_t rec[8] = {tapes[t].n_vars, 0, (uint32_t)tapes[t].levels.size() / 4, 0, (uint32_t)tapes[t].ents.size() / 4, 0,
                               (uint32_t)tapes[t].sums.size() / 4, 0};
            auto align4 = [&] { while (blob.size() % 4) blob.push_back(0); };  // sections start on 16-byte boundaries
            align4();
            rec[1] = (uint32_t)blob.size();
            blob.insert(blob.end(), tapes[t].levels.begin(), tapes[t].levels.end());
            rec[3] = (uint32_t)blob.size();
            blob.insert(blob.end(), tapes[t].ents.begin(), tapes[t].ents.end());
            rec[5] = (uint32_t)blob.size();
            blob.insert(blob.end(), tapes[t].sums.begin(), tapes[t].sums.end());
            rec[7] = (uint32_t)blob.size();
            blob.insert(blob.end(), tapes[t].sumops.begin(), tapes[t].sumops.end());
            std::memcpy(&blob[tape_off + 8 * t], rec, sizeof rec);
        }
        blob[0] = kWprogMagic;
        blob[1] = 1;
        blob[2] = n_units;
        blob[3] = n_tapes;
        blob[4] = (uint32_t)aux_base;
        blob[5] = (uint32_t)n_msg_bits;
        blob[6] = msb_first ? 1u : 0u;
        blob[7] = (uint32_t)n_aux_total;
        blob[8] = max_vars;
        blob[9] = max_sums;
        blob[10] = unit_off;
        blob[11] = tape_off;
        blob[12] = (uint32_t)blob.size();
        return blob;
    }
};

inline thread_local WitnessTape* g_tape = nullptr;  // non-null while a synthesis is being recorded

}  // namespace bph
