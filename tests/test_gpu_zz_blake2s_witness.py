"""GPU (-m gpu): SURVEY 8 f-3 for BASELINE configs[2] -- the blake2s witness generated on the device from the message bytes by
the program the front-end recorded at synthesis (one unit per compression, message bits least significant first, chaining
values by plain BLAKE2s on the host) equals, bit for bit, what the gadgets' host closures produce, for the recorded message
and for another one, and the circuit holds with it.  Same kernels as the sha256 chain (wprog_expand_msg / wprog_run); the
program itself is checked without a GPU by tests/test_witness_program_cpu.py."""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("fid,n_bytes", [(2, 200), (0, 64 * 3), (1, 64)])
def test_blake2s_witness_generated_on_device_equals_the_front_end(fid, n_bytes):
    from bellpepper_b200 import ffi, fixtures

    L = ffi.load()
    msg = fixtures.xorshift_bytes(n_bytes)
    with fixtures.Tcs(fid, device=0, named=False) as t:
        t.record_witness_program()
        t.blake2s(msg)
        prog = t.witness_program()
        h = ffi.vp(t.handle)
        n_aux = t.num_aux()
        want = np.zeros((n_aux, 4), np.uint64)
        assert L.bp_cs_witness(h, 1, 0, n_aux, want.ctypes.data) == 0
        assert t.first_unsatisfied_row() == -1
        assert L.bp_cs_set_witness_program(h, prog.ctypes.data, prog.size) == 0, L.bp_cs_last_error(h)
        for m in (msg, bytes((b * 5 + 3) & 0xFF for b in msg)):
            junk = np.ones(n_aux, np.uint8)  # wreck the witness first: the generator must rewrite every aux value
            assert L.bp_cs_set_range_u8(h, 1, 0, n_aux, junk.ctypes.data) == 0
            assert t.first_unsatisfied_row() >= 0
            st = fixtures.blake2s_chain_states(m)
            assert L.bp_cs_generate_witness_async(h, m, len(m), st.ctypes.data, st.size) == 0, L.bp_cs_last_error(h)
            assert t.first_unsatisfied_row() == -1
            got = np.zeros((n_aux, 4), np.uint64)
            assert L.bp_cs_witness(h, 1, 0, n_aux, got.ctypes.data) == 0
            if m is msg:
                assert (got == want).all()
            else:
                with fixtures.Tcs(fid, device=-1, named=False) as rec:
                    rec.blake2s(m)
                    assert (got == rec.host_csr()[4]).all()
        if st.shape[0] > 1:  # a wrong chaining value is caught by the circuit itself
            bad = fixtures.blake2s_chain_states(msg).copy()
            bad[1][3] ^= 4
            assert L.bp_cs_generate_witness_async(h, msg, len(msg), bad.ctypes.data, bad.size) == 0
            assert t.first_unsatisfied_row() >= 0
