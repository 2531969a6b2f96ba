"""CPU checks of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/spear_b200.h declares, host-only entry points work, and compute entry points fail loudly
(no CPU fallback) when there is no CUDA device."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "spear_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(spear_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from fhe_spear_b200 import _native
    names = _declared()
    assert len(names) >= 55
    for n in names:
        assert hasattr(_native.lib, n), f"libspear_b200.so does not export {n}"
    assert sorted(_native.EXPORTED) == names, "ctypes signature table and header disagree"


def test_ctypes_signatures_match_the_header_prototypes():
    """Every prototype of include/spear_b200.h has as many parameters as its ctypes signature, and integer / pointer /
    double parameters sit in the same positions (a drifted binding would pass garbage across the C ABI)."""
    from fhe_spear_b200 import _native
    src = open(os.path.join(ROOT, "include", "spear_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    protos = dict(re.findall(r"\b(spear_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", src))
    assert set(protos) == set(_native.EXPORTED)

    def kind_of_c(param):
        param = param.strip()
        if "*" in param or "[" in param:
            return "ptr"
        if re.match(r"(const\s+)?double\b", param):
            return "f64"
        return "int"

    def kind_of_ctypes(t):
        if t in (C.c_double, C.c_float):
            return "f64"
        if t in (C.c_int, C.c_uint32, C.c_uint64, C.c_size_t, C.c_int64):
            return "int"
        return "ptr"
    for name, params in protos.items():
        plist = [] if params.strip() in ("", "void") else [p for p in params.split(",")]
        argtypes = getattr(_native.lib, name).argtypes or []
        assert len(plist) == len(argtypes), f"{name}: header has {len(plist)} parameters, ctypes {len(argtypes)}"
        for i, (cp, ct) in enumerate(zip(plist, argtypes)):
            assert kind_of_c(cp) == kind_of_ctypes(ct), f"{name}: parameter {i} ({cp.strip()}) bound as {ct}"


def test_host_only_entry_points():
    from fhe_spear_b200 import pyPhantom as ph
    from oracle.oracle import Oracle
    bits = [59] * 27
    mods = [int(m) for m in ph.create_coeff_modulus(32768, bits)]
    assert mods == [int(x) for x in Oracle.create_coeff_modulus(32768, bits)]
    assert all(m % 65536 == 1 and m.bit_length() == 59 for m in mods) and len(set(mods)) == 27
    # generator 5 / conjugation 2N-1, as pinned by reference scripts/bootstrap_generation.py:18-26
    N = 32768
    assert ph.get_elts_from_steps([1, 2, 46], N) == [pow(5, s, 2 * N) for s in (1, 2, 46)]
    assert ph.get_elt_from_step(0, N) == 2 * N - 1
    assert ph.get_elt_from_step(-1, N) == pow(5, N // 2 - 1, 2 * N)
    assert ph.scheme_type.ckks and ph.ckks == ph.scheme_type.ckks


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from fhe_spear_b200 import pyPhantom as ph
    parms = ph.params(ph.scheme_type.ckks)
    parms.set_poly_modulus_degree(1024)
    parms.set_coeff_modulus(ph.create_coeff_modulus(1024, [59, 59, 59]))
    parms.set_special_modulus_size(1)
    with pytest.raises(RuntimeError, match="CUDA"):
        ph.context(parms)


def test_product_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "fhe_spear_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                # comments may cite the oracle; code may not import, include, link or dlopen it
                assert not re.search(r"^\s*(import|from)\s+oracle\b", text, flags=re.M), f
                assert "libspear_oracle" not in text and not re.search(r"#include[^\n]*oracle", text), f
