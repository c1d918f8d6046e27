"""CPU checks of the witness program: the tape the front-end records while it synthesizes a sha256 chain,
replayed by a plain-Python twin of the device kernel (kernels.cuh: wprog_run), must reproduce the front-end's own witness bit
for bit -- for the recorded message and for ANY other message of the same length."""
import hashlib

import numpy as np
import pytest

from bellpepper_b200 import fixtures


def replay(prog: np.ndarray, msg: bytes, states: np.ndarray) -> np.ndarray:
    """Python twin of wprog_expand_msg + wprog_run: returns the aux witness (0/1 per variable)."""
    w = [int(x) for x in prog]
    assert w[0] == 0x50575042 and w[1] == 1
    n_units, n_tapes, msg_base, n_msg, msb, n_aux = w[2], w[3], w[4], w[5], w[6], w[7]
    unit_off, tape_off = w[10], w[11]
    assert 8 * len(msg) == n_msg and states.shape == (n_units, 8)
    out = np.zeros(n_aux, np.uint8)

    def msg_bit(g):
        b = msg[g >> 3]
        return (b >> (7 - (g & 7))) & 1 if msb else (b >> (g & 7)) & 1

    for i in range(n_msg):
        out[msg_base + i] = msg_bit(i)
    for u in range(n_units):
        tape, aux0, msg0, st = w[unit_off + 4 * u: unit_off + 4 * u + 4]
        n_vars, lev, n_lev, ent, n_ent, sm, n_sum, sop = w[tape_off + 8 * tape: tape_off + 8 * tape + 8]
        vals = [0] * n_vars
        sumv = [0] * max(n_sum, 1)

        def value(op, mask):
            kind, p = op >> 29, op & mask
            if kind < 2:
                v = kind
            elif kind < 4:
                v = vals[p]
            elif kind < 6:
                v = msg_bit(msg0 + p)
            else:
                v = (int(states[st][p >> 5]) >> (p & 31)) & 1
            return v ^ 1 if kind >= 2 and (kind & 1) else v

        for l in range(n_lev):
            e0, e1, s0, s1 = w[lev + 4 * l: lev + 4 * l + 4]
            for s in range(s0, s1):
                first, n_ops, lo, hi = w[sm + 4 * s: sm + 4 * s + 4]
                acc = lo | (hi << 32)
                for k in range(n_ops):
                    op = w[sop + first + k]
                    acc += value(op, 0x00FFFFFF) << ((op >> 24) & 31)
                sumv[s] = acc
            new = []
            for e in range(e0, e1):
                r0, a, b, c = w[ent + 4 * e: ent + 4 * e + 4]
                op, res = r0 >> 28, r0 & 0x0FFFFFFF
                if op == 7:
                    v = (sumv[a] >> b) & 1
                else:
                    x, y = value(a, 0x1FFFFFFF), value(b, 0x1FFFFFFF)
                    if op == 1:
                        v = x ^ y
                    elif op == 2:
                        v = x & y
                    elif op == 3:
                        v = x & (y ^ 1)
                    elif op == 4:
                        v = (x ^ 1) & (y ^ 1)
                    else:
                        z = value(c, 0x1FFFFFFF)
                        v = (x & y) ^ ((x ^ 1) & z) if op == 5 else (x & y) ^ (x & z) ^ (y & z)
                new.append((res, v))
            for res, v in new:  # a level only reads lower levels
                vals[res] = v
        out[aux0: aux0 + n_vars] = vals
    return out


_K256 = [
    0x428A2F98, 0x71374491, 0xB5C0FBCF, 0xE9B5DBA5, 0x3956C25B, 0x59F111F1, 0x923F82A4, 0xAB1C5ED5, 0xD807AA98, 0x12835B01, 0x243185BE,
    0x550C7DC3, 0x72BE5D74, 0x80DEB1FE, 0x9BDC06A7, 0xC19BF174, 0xE49B69C1, 0xEFBE4786, 0x0FC19DC6, 0x240CA1CC, 0x2DE92C6F, 0x4A7484AA,
    0x5CB0A9DC, 0x76F988DA, 0x983E5152, 0xA831C66D, 0xB00327C8, 0xBF597FC7, 0xC6E00BF3, 0xD5A79147, 0x06CA6351, 0x14292967, 0x27B70A85,
    0x2E1B2138, 0x4D2C6DFC, 0x53380D13, 0x650A7354, 0x766A0ABB, 0x81C2C92E, 0x92722C85, 0xA2BFE8A1, 0xA81A664B, 0xC24B8B70, 0xC76C51A3,
    0xD192E819, 0xD6990624, 0xF40E3585, 0x106AA070, 0x19A4C116, 0x1E376C08, 0x2748774C, 0x34B0BCB5, 0x391C0CB3, 0x4ED8AA4A, 0x5B9CCA4F,
    0x682E6FF3, 0x748F82EE, 0x78A5636F, 0x84C87814, 0x8CC70208, 0x90BEFFFA, 0xA4506CEB, 0xBEF9A3F7, 0xC67178F2]
_IV256 = [0x6A09E667, 0xBB67AE85, 0x3C6EF372, 0xA54FF53A, 0x510E527F, 0x9B05688C, 0x1F83D9AB, 0x5BE0CD19]


def _compress256(state, block):
    """FIPS 180-4 section 6.2.2 in plain Python (hashlib exposes no intermediate states)."""
    M = 0xFFFFFFFF
    rotr = lambda x, r: ((x >> r) | (x << (32 - r))) & M  # noqa: E731
    w = [int.from_bytes(block[4 * i:4 * i + 4], "big") for i in range(16)]
    for i in range(16, 64):
        s0 = rotr(w[i - 15], 7) ^ rotr(w[i - 15], 18) ^ (w[i - 15] >> 3)
        s1 = rotr(w[i - 2], 17) ^ rotr(w[i - 2], 19) ^ (w[i - 2] >> 10)
        w.append((w[i - 16] + s0 + w[i - 7] + s1) & M)
    a, b, c, d, e, f, g, h = state
    for i in range(64):
        t1 = (h + (rotr(e, 6) ^ rotr(e, 11) ^ rotr(e, 25)) + ((e & f) ^ (~e & M & g)) + _K256[i] + w[i]) & M
        t2 = ((rotr(a, 2) ^ rotr(a, 13) ^ rotr(a, 22)) + ((a & b) ^ (a & c) ^ (b & c))) & M
        a, b, c, d, e, f, g, h = (t1 + t2) & M, a, b, c, (d + t1) & M, e, f, g
    return [(x + y) & M for x, y in zip(state, (a, b, c, d, e, f, g, h))]


def _chain_states_py(msg):
    padded = msg + b"\x80" + b"\x00" * ((55 - len(msg)) % 64) + (8 * len(msg)).to_bytes(8, "big")
    assert len(padded) % 64 == 0
    states, st = [], list(_IV256)
    for b in range(len(padded) // 64):
        states.append(st)
        st = _compress256(st, padded[64 * b:64 * b + 64])
    return states, st


def check_chain_states():
    import hashlib

    for n in (0, 1, 55, 56, 63, 64, 65, 119, 120, 128, 1000, 64 * 40 - 9):
        msg = fixtures.xorshift_bytes(n)
        want, final = _chain_states_py(msg)
        assert b"".join(x.to_bytes(4, "big") for x in final) == hashlib.sha256(msg).digest()  # the restatement itself
        st = fixtures.sha256_chain_states(msg)
        assert st.shape == ((n + 9 + 63) // 64, 8)
        assert st.tolist() == want, n


def test_chain_states_match_sha256():
    """bp_sha256_chain_states (the hash state before every compression block: what bp_cs_generate_witness_async takes) against
    plain-Python SHA-256, itself checked against hashlib -- on whichever kernel this CPU selects ..."""
    check_chain_states()


def test_chain_states_match_sha256_portable_kernel(monkeypatch):
    """... and on the portable one (BP_NO_SHANI is read at every call)."""
    monkeypatch.setenv("BP_NO_SHANI", "1")
    check_chain_states()


def test_recorded_program_reproduces_the_witness():
    fid, blocks = 1, 3
    msg = fixtures.chain_message(blocks)
    with fixtures.Tcs(fid, device=-1, named=False) as t:
        t.record_witness_program()
        digest, _ = t.sha256(msg)
        prog = t.witness_program()
        _, _, _, _, aux = t.host_csr()
    assert digest == hashlib.sha256(msg).digest()
    assert prog[2] == blocks and prog[7] == aux.shape[0] and prog[5] == 8 * len(msg)
    assert prog[3] <= 3  # tapes are shared between identical blocks (first / middle / last)
    want = aux[:, 0].astype(np.uint8)
    assert not aux[:, 1:].any() and want.max() <= 1
    got = replay(prog, msg, fixtures.sha256_chain_states(msg))
    assert (got == want).all()
    # ANOTHER message of the same length through the same program == a fresh synthesis of that message
    msg2 = bytes((b * 7 + 13) & 0xFF for b in msg)
    with fixtures.Tcs(fid, device=-1, named=False) as t:
        t.sha256(msg2)
        aux2 = t.host_csr()[4]
    got2 = replay(prog, msg2, fixtures.sha256_chain_states(msg2))
    assert (got2 == aux2[:, 0].astype(np.uint8)).all()


def test_generic_blocks_share_one_tape():
    fid, blocks = 1, 6
    msg = fixtures.chain_message(blocks)
    with fixtures.Tcs(fid, device=-1, named=False) as t:
        t.record_witness_program()
        t.sha256(msg)
        prog = t.witness_program()
    n_units, n_tapes, unit_off = int(prog[2]), int(prog[3]), int(prog[10])
    tapes = [int(prog[unit_off + 4 * u]) for u in range(n_units)]
    assert n_units == blocks and n_tapes == 3
    assert tapes[0] != tapes[1] and len(set(tapes[1:-1])) == 1 and tapes[-1] != tapes[1]


# ---- blake2s (configs[2]): one unit per compression, message bits least significant first ------------------------------------
def test_blake2s_chain_states_match_hashlib():
    """The state after the last compression is the digest; intermediate states by re-hashing the prefix is not possible with
    hashlib (the last block is flagged), so: the final value from the last state + the documented compression via the
    gadget itself in test_blake2s_program_reproduces_the_witness, and here the shape, the parameter block and determinism."""
    for n in (0, 1, 63, 64, 65, 128, 1000):
        msg = fixtures.xorshift_bytes(n)
        st = fixtures.blake2s_chain_states(msg)
        assert st.shape == (max(1, (n + 63) // 64), 8)
        iv = [0x6A09E667, 0xBB67AE85, 0x3C6EF372, 0xA54FF53A, 0x510E527F, 0x9B05688C, 0x1F83D9AB, 0x5BE0CD19]
        want0 = list(iv)
        want0[0] ^= 0x01010020
        want0[6] ^= int.from_bytes(b"1234", "little")
        want0[7] ^= int.from_bytes(b"5678", "little")
        assert list(st[0]) == want0
    # a message of whole blocks hashed without the final flag chains exactly like the unflagged prefix of a longer one
    a = fixtures.blake2s_chain_states(fixtures.xorshift_bytes(64 * 3 + 5))
    b = fixtures.blake2s_chain_states(fixtures.xorshift_bytes(64 * 5))
    assert (a[:4] == b[:4]).all()  # states before blocks 0..3 depend on the first three whole blocks only


@pytest.mark.parametrize("fid,n_bytes", [(2, 64), (2, 200), (0, 130), (1, 64 * 4)])
def test_blake2s_program_reproduces_the_witness(fid, n_bytes):
    msg = fixtures.xorshift_bytes(n_bytes)
    with fixtures.Tcs(fid, device=-1, named=False) as t:
        t.record_witness_program()
        digest = t.blake2s(msg)
        prog = t.witness_program()
        lens, cols, coeffs, inputs, aux = t.host_csr()
    assert digest == hashlib.blake2s(msg, digest_size=32, person=b"12345678").digest()
    n_blocks = max(1, (n_bytes + 63) // 64)
    assert int(prog[2]) == n_blocks and int(prog[6]) == 0 and int(prog[5]) == 8 * n_bytes and int(prog[7]) == aux.shape[0]
    want = aux[:, 0].astype(np.uint8)
    assert not aux[:, 1:].any() and want.max() <= 1
    got = replay(prog, msg, fixtures.blake2s_chain_states(msg))
    assert (got == want).all()
    # recording does not change the circuit
    with fixtures.Tcs(fid, device=-1, named=False) as t:
        t.blake2s(msg)
        plain = t.host_csr()
    for x, y in zip(plain, (lens, cols, coeffs, inputs, aux)):
        assert x.shape == y.shape and (x == y).all()
    # ANOTHER message of the same length through the same program == a fresh synthesis of that message
    msg2 = bytes((b * 11 + 5) & 0xFF for b in msg)
    with fixtures.Tcs(fid, device=-1, named=False) as t:
        t.blake2s(msg2)
        aux2 = t.host_csr()[4]
    got2 = replay(prog, msg2, fixtures.blake2s_chain_states(msg2))
    assert (got2 == aux2[:, 0].astype(np.uint8)).all()
    # a wrong chaining value gives a different witness (the states are really read)
    if n_blocks > 1:
        st = fixtures.blake2s_chain_states(msg).copy()
        st[1][2] ^= 1 << 7
        assert (replay(prog, msg, st) != want).any()
