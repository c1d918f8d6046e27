"""Extract the known-answer constants of the hot path's producers from the reference's OWN tests (run in the build
container, where /root/reference is mounted; the GPU box only sees the committed reference_kats.json).

    python tests/golden/extract_reference_kats.py  > tests/golden/reference_kats.json

Every entry carries the reference file and line it was read from."""
import json
import os
import re
import sys

REF = os.environ.get("BP_REFERENCE", "/root/reference")


def grab(path, pattern, group=1, all_=False, conv=str):
    full = os.path.join(REF, path)
    out = []
    with open(full) as fh:
        for ln, line in enumerate(fh, 1):
            m = re.search(pattern, line)
            if m:
                out.append({"value": conv(m.group(group)), "at": f"{path}:{ln}"})
    if not out:
        sys.exit(f"pattern {pattern!r} not found in {path}")
    return out if all_ else out[0]


def main():
    sha = "crates/bellpepper/src/gadgets/sha256.rs"
    b2 = "crates/bellpepper/src/gadgets/blake2s.rs"
    kats = {
        "source": "argumentcomputer/bellpepper @ 7275595cc7121b0adc8ce5e88ce1281aa6aff601 (tests of the reference itself)",
        "sha256": {
            "num_constraints_asserts": grab(sha, r"assert_eq!\(cs\.num_constraints\(\)(?: - 512)?, (\d+)\)", all_=True, conv=int),
            "blank_block_digest": grab(sha, r'hex!\("([0-9a-f]{64})"\)'),
            "rng_seed_bytes": grab(sha, r"0x59, 0x62, 0xbe, (0x[0-9a-f]{2}), 0x76", conv=lambda s: int(s, 16)),
        },
        "blake2s": {
            "num_constraints_asserts": grab(b2, r"assert_eq!\(cs\.num_constraints\(\), (\d+)\)", all_=True, conv=int),
            "digests": grab(b2, r'hex!\("([0-9a-f]{64})"\)', all_=True),
            "rng_seed_bytes": grab(b2, r"0x59, 0x62, 0xbe, (0x[0-9a-f]{2}), 0x76", conv=lambda s: int(s, 16)),
            "personalization": grab(b2, r'b"(12345678)"'),
        },
        "test_cs": {
            "which_is_unsatisfied": grab("crates/bellpepper-core/src/util_cs/test_cs.rs", r'cs\.which_is_unsatisfied\(\) == Some\("(\w+)"\)'),
        },
        "multieq_capacity_source": grab("crates/bellpepper/src/gadgets/multieq.rs", r"(Scalar::CAPACITY)"),
    }
    json.dump(kats, sys.stdout, indent=1, sort_keys=True)
    print()


if __name__ == "__main__":
    main()
