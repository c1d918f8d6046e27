"""CPU-side checks of the boundary: the library loads without a GPU and exports every declared symbol."""
import ctypes
import os
import re

from bellpepper_b200 import ffi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "bp_r1cs.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bp_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    syms = declared_symbols()
    assert len(syms) >= 20
    assert set(syms) == set(ffi.SIGNATURES), set(syms) ^ set(ffi.SIGNATURES)


def test_library_loads_and_exports_every_symbol():
    L = ffi.load()
    for s in declared_symbols():
        assert hasattr(L, s), s
    assert L.bp_abi_version() == 1


def test_no_cpu_fallback():
    """Without a CUDA device bp_cs_new must fail with BP_E_CUDA (there is no CPU path)."""
    import torch

    if torch.cuda.is_available():
        return
    L = ffi.load()
    h = ffi.vp()
    assert L.bp_cs_new(0, 0, 0, 0, 0, ctypes.byref(h)) == ffi.BP_E_CUDA
    assert not h.value
    assert L.bp_cs_new(7, 0, 0, 0, 0, ctypes.byref(h)) == ffi.BP_E_ARG


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "bellpepper_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert not re.search(r"#\s*include[^\n]*oracle|libbp_oracle", src), f
                # the only library the product may bind at run time is NCCL (multi-GPU groups)
                for m in re.finditer(r"dlopen\s*\(", src):
                    assert "libnccl" in src[max(0, m.start() - 400):m.start()], (f, "dlopen of something that is not NCCL")


def test_split_rows_by_nnz_is_host_only_and_balances_terms():
    """bp_split_rows_by_nnz (shards balanced by terms, not rows) needs no device."""
    import ctypes

    import numpy as np

    from bellpepper_b200 import ffi

    L = ffi.load()
    rng = np.random.default_rng(1)
    lens = rng.integers(0, 6, size=3 * 5000).astype(np.uint32)
    lens[3 * 100] = 900  # a MultiEq-like fat row
    lens[3 * 4000 + 2] = 700
    n = lens.size // 3
    w = lens.reshape(-1, 3).sum(1).astype(np.int64) + 1
    for world in (1, 2, 4, 8):
        b = np.zeros(world + 1, np.uint64)
        assert L.bp_split_rows_by_nnz(lens.ctypes.data, n, world, b.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64))) == 0
        bi = b.astype(np.int64)
        assert bi[0] == 0 and bi[-1] == n and (np.diff(bi) >= 0).all()
        per = [int(w[bi[r]:bi[r + 1]].sum()) for r in range(world)]
        assert max(per) - min(per) <= 2 * 901, per  # within a fat row of each other
    assert L.bp_split_rows_by_nnz(None, 0, 3, b.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64))) == 0
    assert L.bp_split_rows_by_nnz(lens.ctypes.data, n, 0, b.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64))) == -5


def test_fixture_header_binding_and_both_libraries_agree():
    """include/bp_fixtures.h (the C++ front-end's C entry points) <-> bellpepper_b200/fixtures.py <-> what libbp_r1cs.so and the
    host-only libbp_frontend.so export."""
    from bellpepper_b200 import fixtures

    text = open(os.path.join(ROOT, "include", "bp_fixtures.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    syms = sorted(set(re.findall(r"\b(bp_[a-z0-9_]+)\s*\(", text)))
    assert len(syms) >= 25
    assert set(syms) == set(fixtures._SIGS), set(syms) ^ set(fixtures._SIGS)
    for lib in (fixtures._lib(), fixtures._host_lib()):
        for s in syms:
            assert hasattr(lib, s), s


def test_rust_binding_declares_the_header():
    """rust/bellpepper-b200/src/ffi.rs (uncompiled here: no Rust toolchain) names exactly the header's entry points, each with
    the header's number of arguments."""
    hdr = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "bp_r1cs.h")).read(), flags=re.S)
    rs = open(os.path.join(ROOT, "rust", "bellpepper-b200", "src", "ffi.rs")).read()

    def arity(args):
        args = args.strip()
        return 0 if args in ("", "void") else args.count(",") + 1

    c = {m.group(1): arity(m.group(2)) for m in re.finditer(r"\b(bp_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", hdr)}
    r = {m.group(1): arity(m.group(2)) for m in re.finditer(r"pub fn (bp_[a-z0-9_]+)\s*\(([^)]*)\)", rs)}
    assert set(c) == set(r), set(c) ^ set(r)
    assert c == r, {k: (c[k], r[k]) for k in c if c[k] != r[k]}


def _c_to_rust(arg):
    """One C parameter (or return type) of include/bp_r1cs.h -> the Rust FFI type that binds it."""
    arg = re.sub(r"\[[^\]]*\]", "*", arg.strip())  # arrays decay to pointers
    toks = re.findall(r"[A-Za-z_][A-Za-z0-9_]*|\*", arg)
    types = {"const", "void", "int", "char", "uint8_t", "uint32_t", "uint64_t", "int64_t", "bp_cs", "bp_group"}
    if toks and toks[-1] not in types and toks[-1] != "*":
        toks = toks[:-1]  # the parameter's name
    elif len(toks) >= 2 and toks[-1] == "*" and toks[-2] not in types:
        toks = toks[:-2] + ["*"]  # `uint8_t id[N]`: the name sits before the decayed `*`
    base = [t for t in toks if t not in ("const", "*")]
    assert len(base) == 1, arg
    r = {"void": "c_void", "int": "c_int", "char": "c_char", "uint8_t": "u8", "uint32_t": "u32", "uint64_t": "u64", "int64_t": "i64",
         "bp_cs": "bp_cs", "bp_group": "bp_group"}[base[0]]
    stars, const = toks.count("*"), "const" in toks
    if stars == 0:
        return r
    inner = f"*{'const' if const else 'mut'} {r}"
    return inner if stars == 1 else f"*mut {inner}"


def test_rust_binding_types_match_the_header():
    """Every parameter and return type of rust/bellpepper-b200/src/ffi.rs is the Rust spelling of the header's C type
    (pointer depth, constness, integer width): what a compiler would check, checked here because the image has no rustc."""
    hdr = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "bp_r1cs.h")).read(), flags=re.S)
    rs = open(os.path.join(ROOT, "rust", "bellpepper-b200", "src", "ffi.rs")).read()
    c = {m.group(2): (m.group(1).strip(), [a for a in m.group(3).split(",") if a.strip() not in ("", "void")])
         for m in re.finditer(r"([A-Za-z_][A-Za-z0-9_ \*]*?)\b(bp_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", hdr)}
    r = {m.group(1): ([a.split(":", 1)[1].strip() for a in m.group(2).split(",") if ":" in a], (m.group(3) or "").strip())
         for m in re.finditer(r"pub fn (bp_[a-z0-9_]+)\s*\(([^)]*)\)\s*(?:->\s*([^;]+))?;", rs)}
    assert len(c) >= 50 and set(c) == set(r)
    for name, (ret, args) in c.items():
        assert [_c_to_rust(a) for a in args] == r[name][0], name
        assert ("" if ret == "void" else _c_to_rust(ret + " x" if "*" not in ret else ret)) == r[name][1], name
    # the checker itself notices a wrong width / constness / depth
    assert _c_to_rust("const uint64_t* vals_le") == "*const u64" and _c_to_rust("bp_cs** out") == "*mut *mut bp_cs"
    assert _c_to_rust("uint8_t id[BP_GROUP_ID_BYTES]") == "*mut u8" and _c_to_rust("const uint8_t id[BP_GROUP_ID_BYTES]") == "*const u8"
    assert _c_to_rust("int64_t* row") != _c_to_rust("uint64_t* row") and _c_to_rust("int is_aux") == "c_int"


def test_ctypes_binding_arities_and_scalar_types_match_the_header():
    """bellpepper_b200/ffi.py: every entry has the header's number of parameters; non-pointer parameters have the header's width
    and signedness (a pointer is a c_void_p / POINTER / c_char_p there)."""
    hdr = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "bp_r1cs.h")).read(), flags=re.S)
    scalars = {"c_int": ctypes.c_int, "u32": ctypes.c_uint32, "u64": ctypes.c_uint64, "i64": ctypes.c_int64, "u8": ctypes.c_uint8}
    for m in re.finditer(r"([A-Za-z_][A-Za-z0-9_ \*]*?)\b(bp_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", hdr):
        ret, name, args = m.group(1).strip(), m.group(2), [a for a in m.group(3).split(",") if a.strip() not in ("", "void")]
        res, argtypes = ffi.SIGNATURES[name]
        assert len(argtypes) == len(args), name
        for a, t in zip(args, argtypes):
            r = _c_to_rust(a)
            if r.startswith("*"):
                assert t in (ctypes.c_void_p, ctypes.c_char_p) or hasattr(t, "_type_") and not isinstance(t._type_, str), (name, a, t)
            else:
                assert t is scalars[r], (name, a, t)
        if ret == "void":
            assert res is None, name
        elif "*" in ret:
            assert res in (ctypes.c_char_p, ctypes.c_void_p), name
        else:
            assert res is scalars[_c_to_rust(ret + " x")], name


def test_fixture_binding_arities_and_scalar_types_match_the_header():
    from bellpepper_b200 import fixtures

    hdr = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "bp_fixtures.h")).read(), flags=re.S)
    hdr = hdr.replace("bp_tcs", "bp_cs")  # (an opaque handle either way: reuse the type table)
    scalars = {"c_int": ctypes.c_int, "u32": ctypes.c_uint32, "u64": ctypes.c_uint64, "i64": ctypes.c_int64, "u8": ctypes.c_uint8}
    n = 0
    for m in re.finditer(r"([A-Za-z_][A-Za-z0-9_ \*]*?)\b(bp_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", hdr):
        ret, name, args = m.group(1).strip(), m.group(2).replace("bp_cs_", "bp_tcs_", 1), [a for a in m.group(3).split(",") if a.strip() not in ("", "void")]
        if name not in fixtures._SIGS:
            name = m.group(2)  # bp_sha256_chain_states, bp_blake2s_chain_states, bp_wcs_selftest keep their names
        res, argtypes = fixtures._SIGS[name]
        assert len(argtypes) == len(args), name
        for a, t in zip(args, argtypes):
            r = _c_to_rust(a)
            if r.startswith("*"):
                assert t in (ctypes.c_void_p, ctypes.c_char_p) or hasattr(t, "_type_") and not isinstance(t._type_, str), (name, a, t)
            else:
                assert t is scalars[r], (name, a, t)
        if ret != "void" and "*" not in ret:
            assert res is scalars[_c_to_rust(ret + " x")], name
        n += 1
    assert n == len(fixtures._SIGS)
