// Example: the reference's `test_full_block`-style use (crates/bellpepper/src/gadgets/sha256.rs:310-336, 365-417) through the
// C++ host mirror: synthesize sha256 of a message with the gadget, check satisfaction on the B200, flip a witness bit, ask which
// constraint fails.
//
//   g++ -std=c++17 -O2 -I. examples/sha256_check.cpp -Lbellpepper_b200 -lbp_r1cs -Wl,-rpath,$PWD/bellpepper_b200 -o /tmp/sha256_check
//   /tmp/sha256_check 200
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "bellpepper_b200/csrc/host/gadgets.hpp"

using namespace bph;

int main(int argc, char** argv) {
    const size_t n_bytes = argc > 1 ? (size_t)atoi(argv[1]) : 64;
    bp_cs* h = nullptr;
    if (bp_cs_new(/*field: Pallas Fr*/ 1, /*device*/ 0, 0, 0, 0, &h) != BP_OK) {
        std::fprintf(stderr, "no CUDA device (there is no CPU path)\n");
        return 2;
    }
    int rc = 1;
    {
        DeviceSink sink(h);
        TestConstraintSystem cs(1, &sink);
        std::vector<Boolean> bits;
        for (size_t i = 0; i < n_bytes; ++i)
            for (int j = 7; j >= 0; --j) {
                auto ns = cs.ns([&] { return "input bit " + std::to_string(i) + " " + std::to_string(j); });
                bits.push_back(Boolean::from(AllocatedBit::alloc(ns, (OptBool)(((i * 131 + 7) >> j) & 1))));
            }
        std::vector<Boolean> digest = sha256(cs, bits);
        std::printf("%zu bytes -> %llu constraints, %llu aux variables, %zu output bits\n", n_bytes,
                    (unsigned long long)cs.num_constraints(), (unsigned long long)cs.num_aux(), digest.size());
        const bool ok = cs.is_satisfied();
        std::printf("is_satisfied: %s\n", ok ? "true" : "false");
        const std::string victim = "block 0/w extension 16/computation of w[i]/result bit 0/boolean";
        const Fr old = cs.get(victim);
        cs.set(victim, old.is_zero() ? Fr::one() : Fr::zero());
        auto bad = cs.which_is_unsatisfied();
        std::printf("after flipping \"%s\": which_is_unsatisfied = %s\n", victim.c_str(), bad ? bad->c_str() : "(none)");
        cs.set(victim, old);
        rc = (ok && bad && cs.is_satisfied()) ? 0 : 1;
    }
    bp_cs_free(h);
    return rc;
}
