"""TEST INFRASTRUCTURE (oracle) -- not part of the product path.

Field definitions for the three 255-bit prime fields the hot path runs over.

The reference is generic over `ff::PrimeField` (`ff = "0.13.0"`, /root/reference/Cargo.toml:12)
and its only concrete field is `blstrs::Scalar` (`blstrs = "0.7.0"`, Cargo.toml:10; dev-dependency,
crates/bellpepper-core/Cargo.toml:26).  Neither crate is vendored under /root/reference, so the
arithmetic is restated here from the published field definitions:

* BLS12-381 Fr  (blstrs::Scalar)                         -- PINNED by the reference's own tests
* Pallas Fr = pasta Fq (pasta_curves 0.5 `Fq`)           -- parity UNPINNED (no reference test, F5)
* Vesta  Fr = pasta Fp (pasta_curves 0.5 `Fp`)           -- parity UNPINNED

`PrimeField::to_repr()` is the canonical little-endian 32-byte encoding (the reference reverses it to
get big-endian at crates/bellpepper-core/src/util_cs/test_cs.rs:108-111); elements cross every
boundary in this repo as 4 x u64 little-endian limbs of the canonical residue.
"""

from __future__ import annotations

from dataclasses import dataclass

FIELD_BLS12_381_FR = 0
FIELD_PALLAS_FR = 1  # pasta Fq
FIELD_VESTA_FR = 2  # pasta Fp


@dataclass(frozen=True)
class Field:
    fid: int
    name: str
    p: int
    num_bits: int = 255  # ff::PrimeField::NUM_BITS
    capacity: int = 254  # ff::PrimeField::CAPACITY

    @property
    def limbs64(self):
        return [(self.p >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)]

    @property
    def inv64(self) -> int:
        """-p^-1 mod 2^64 (Montgomery constant for 64-bit CIOS)."""
        return (-pow(self.p, -1, 1 << 64)) % (1 << 64)

    @property
    def inv32(self) -> int:
        """-p^-1 mod 2^32 (Montgomery constant for the 32-bit device path)."""
        return (-pow(self.p, -1, 1 << 32)) % (1 << 32)

    def pow2(self, k: int) -> int:
        return pow(2, k, self.p)


FIELDS = {
    FIELD_BLS12_381_FR: Field(
        FIELD_BLS12_381_FR,
        "bls12_381_fr",
        0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001,
    ),
    FIELD_PALLAS_FR: Field(
        FIELD_PALLAS_FR,
        "pallas_fr",
        0x40000000000000000000000000000000224698FC0994A8DD8C46EB2100000001,
    ),
    FIELD_VESTA_FR: Field(
        FIELD_VESTA_FR,
        "vesta_fr",
        0x40000000000000000000000000000000224698FC094CF91B992D30ED00000001,
    ),
}

BY_NAME = {f.name: f for f in FIELDS.values()}


def to_limbs(x: int):
    """Canonical residue -> 4 little-endian u64 limbs."""
    return [(x >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)]


def from_limbs(l) -> int:
    return int(l[0]) | (int(l[1]) << 64) | (int(l[2]) << 128) | (int(l[3]) << 192)
