// Host build of bellpepper_b200/csrc/field.cuh (plain-C twins of the PTX carry chains) for CPU unit tests.
#include "../../bellpepper_b200/csrc/field.cuh"
#include <cstring>

template <int F> static void run(int op, const uint32_t* a, const uint32_t* b, uint32_t* out) {
    switch (op) {
        case 0: bp::mul_wide(out, a, b); break;                       // out[16]
        case 1: bp::redc_acc<F>(out, a); break;                       // a[17] -> out[8] in [0,2p)
        case 2: bp::mont_mul<F>(out, a, b); break;                    // out[8] canonical
        case 3: { uint32_t acc[17]; std::memcpy(acc, out, 68); bp::mac_wide(acc, a, b); std::memcpy(out, acc, 68); } break;
        case 4: { std::memcpy(out, a, 32); bp::reduce_once<F>(out); } break;
        case 5: out[0] = bp::is_zero_mod_p<F>(a); break;
        case 6: out[0] = bp::is_canonical<F>(a); break;
        case 7: bp::redc8<F>(out, a); break;                          // a[8] -> out[8]
    }
}
extern "C" int field_host_op(int field, int op, const uint32_t* a, const uint32_t* b, uint32_t* out) {
    switch (field) {
        case 0: run<0>(op, a, b, out); return 0;
        case 1: run<1>(op, a, b, out); return 0;
        case 2: run<2>(op, a, b, out); return 0;
    }
    return -1;
}

// ---- plain-arithmetic primitives ----
template <int F> static void run2(int op, const uint32_t* a, const uint32_t* b, uint32_t* out) {
    switch (op) {
        case 10: bp::neg_mod<F>(out, a); break;                                  // a[8] -> out[8]
        case 11: bp::acc_add8<9>(out, a); break;                                 // out[9] += a[8]
        case 12: bp::acc_add8<17>(out, a); break;                                // out[17] += a[8]
        case 13: bp::acc_mad_small<9>(out, a, b[0]); break;                      // out[9] += b0 * a[8]
        case 14: bp::acc_mad_small<17>(out, a, b[0]); break;
        case 15: bp::reduce_8p<F>(out, a); break;                                // a[9] -> out[8]
        case 16: { uint32_t s; out[0] = bp::classify_coeff<F>(a, &s); out[1] = s; } break;
    }
}
extern "C" int field_host_op2(int field, int op, const uint32_t* a, const uint32_t* b, uint32_t* out) {
    switch (field) {
        case 0: run2<0>(op, a, b, out); return 0;
        case 1: run2<1>(op, a, b, out); return 0;
        case 2: run2<2>(op, a, b, out); return 0;
    }
    return -1;
}
