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
