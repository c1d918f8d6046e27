"""CPU checks of the witness program: the tape the front-end records while it synthesizes a sha256 chain,
replayed by a plain-Python twin of the device kernel (kernels.cuh: wprog_run), must reproduce the front-end's own witness bit
for bit -- for the recorded message and for ANY other message of the same length."""
import hashlib

import numpy as np

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


def test_chain_states_match_hashlib():
    for n in (0, 1, 55, 56, 64, 119, 120, 1000):
        msg = fixtures.xorshift_bytes(n)
        st = fixtures.sha256_chain_states(msg)
        assert st.shape[0] == (n + 9 + 63) // 64
        assert list(st[0]) == [0x6A09E667, 0xBB67AE85, 0x3C6EF372, 0xA54FF53A, 0x510E527F, 0x9B05688C, 0x1F83D9AB, 0x5BE0CD19]
    # the state after the last block is the digest: run one more byte-block's worth by hashing a longer padded message by hand
    msg = fixtures.xorshift_bytes(200)
    st = fixtures.sha256_chain_states(msg)
    # hashlib cannot expose intermediate states; check the first block boundary through a 64-byte prefix trick instead:
    # sha256(prefix64 || rest) state after block 0 == state after compressing prefix64 from the IV, which the gadget circuit
    # also computes -- covered by test_recorded_program_reproduces_the_witness below (witness of block 1 depends on it)
    assert st.shape == (4, 8)


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
