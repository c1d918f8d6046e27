"""TEST INFRASTRUCTURE (oracle) -- the counter-based synthetic R1CS recipe in pure Python.

This is the specification the C (oracle/bp_oracle.c) and CUDA (bellpepper_b200/csrc/synth.cuh)
generators are checked against.  See DESIGN.md "Synthetic instances" for the recipe and why the
columns are stratified (strictly ascending and unique by construction, which is the reference's
LinearCombination invariant, /root/reference/crates/bellpepper-core/src/lc.rs:74-113).
"""

from __future__ import annotations

from .fields import FIELDS

M64 = (1 << 64) - 1
SEED = 0x5962BE3D763D318D  # the reference's test-RNG seed bytes (crates/bellpepper/src/gadgets/sha256.rs:312-315)
N_INPUTS = 16  # ONE + 15 public inputs


def mix(x: int) -> int:
    x = (x + 0x9E3779B97F4A7C15) & M64
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & M64
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & M64
    return x ^ (x >> 31)


def gkey(seed: int, stream: int, i: int, j: int) -> int:
    return mix((mix((mix(seed ^ ((stream << 56) & M64)) + i) & M64) + j) & M64)


def sample(p: int, h: int) -> int:
    limbs = [0, 0, 0, 0]
    for a in range(64):
        limbs = [mix((h + 4 * a + k + 1) & M64) for k in range(4)]
        limbs[3] &= (1 << 63) - 1
        v = limbs[0] | (limbs[1] << 64) | (limbs[2] << 128) | (limbs[3] << 192)
        if v < p:
            return v
    limbs[3] >>= 2
    return limbs[0] | (limbs[1] << 64) | (limbs[2] << 128) | (limbs[3] << 192)


def lc_len(seed: int, t: int, row: int, lc: int) -> int:
    return 1 + gkey(seed, 1, 3 * row + lc, 0) % (2 * t - 1)


def lc_terms(field_id: int, seed: int, t: int, n_vars: int, n_inputs: int, row: int, lc: int):
    """[(tagged col, canonical coeff)] for one linear combination."""
    p = FIELDS[field_id].p
    lcid = 3 * row + lc
    ln = lc_len(seed, t, row, lc)
    out = []
    for k in range(ln):
        lo, hi = (k * n_vars) // ln, ((k + 1) * n_vars) // ln
        col = lo + gkey(seed, 2, lcid, k) % (hi - lo)
        tagged = col if col < n_inputs else ((col - n_inputs) | (1 << 31))
        out.append((tagged, sample(p, gkey(seed, 3, lcid, k))))
    return out


def witness(field_id: int, seed: int, i: int) -> int:
    return 1 if i == 0 else sample(FIELDS[field_id].p, gkey(seed, 4, i, 0))
