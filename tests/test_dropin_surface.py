"""The drop-in boundary, statically (no GPU): imported the way the reference imports it (`sys.path` pointing at
fhe_spear_b200/, `import pyPhantom`), the module offers every `ph.*` / `phantom.*` name the reference's scripts use and
every method of SURVEY.md section 8(b)'s table.  The scan of the reference's sources runs in the build container only
(/root/reference does not exist on the GPU box); the method table is checked everywhere."""
import glob
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"

METHODS = {   # class -> methods the reference calls on its instances
    "params": ["set_poly_modulus_degree", "set_coeff_modulus", "set_special_modulus_size", "set_galois_elts"],
    "secret_key": ["gen_publickey", "gen_relinkey", "create_galois_keys", "encrypt_symmetric", "decrypt"],
    "public_key": ["encrypt_asymmetric"],
    "ckks_encoder": ["slot_count", "encode_double_vector", "encode_complex_vector", "decode_double_vector",
                     "decode_complex_vector", "encode_double_vector_batch", "encode_complex_vector_batch"],
    "ciphertext": ["set_scale", "chain_index", "scale", "coeff_modulus_size"],
    "plaintext": ["chain_index", "scale", "coeff_modulus_size"],
    "ckks_bootstrapper": ["get_galois_elements", "get_bootstrap_depth", "setup", "keygen", "bootstrap"],
}
FUNCTIONS = ["create_coeff_modulus", "get_elts_from_steps", "get_elt_from_step", "add", "add_plain", "sub", "sub_plain",
             "negate", "add_many", "multiply", "multiply_and_relin", "multiply_plain", "relinearize", "rescale_to_next",
             "mod_switch_to_next", "mod_switch_to", "apply_galois", "rotate", "hoisting", "bsgs_multiply_accumulate",
             "bsgs_from_cpu", "bsgs_complete_from_cpu", "offload_plaintexts", "upload_plaintexts"]


def _top_level_import(code):
    """run `code` in a fresh interpreter whose sys.path points at fhe_spear_b200/ like the reference's PHANTOM_PATH"""
    prog = f"import sys; sys.path.insert(0, {os.path.join(ROOT, 'fhe_spear_b200')!r}); import pyPhantom as ph\n" + code
    return subprocess.run([sys.executable, "-c", prog], capture_output=True, text=True, cwd="/tmp", timeout=300)


def test_top_level_import_offers_the_table_of_section_8b():
    code = f"""
methods = {METHODS!r}
functions = {FUNCTIONS!r}
missing = [f for f in functions if not callable(getattr(ph, f, None))]
missing += [c + '.' + m for c, ms in methods.items() for m in ms if not callable(getattr(getattr(ph, c, None), m, None))]
assert ph.__name__ == 'pyPhantom' and ph.scheme_type.ckks is not None
print('MISSING', missing)
"""
    p = _top_level_import(code)
    assert p.returncode == 0, p.stderr[-2000:]
    assert "MISSING []" in p.stdout, p.stdout


@pytest.mark.skipif(not os.path.isdir(REF), reason="the reference tree exists in the build container only")
def test_every_module_level_name_the_reference_uses_exists():
    names = set()
    for f in glob.glob(os.path.join(REF, "**", "*.py"), recursive=True):
        src = open(f, errors="ignore").read()
        if "pyPhantom" in src:
            names |= set(re.findall(r"\b(?:ph|phantom)\.([A-Za-z_][A-Za-z_0-9]*)", src))
    assert len(names) >= 20
    p = _top_level_import(f"print('MISSING', [n for n in {sorted(names)!r} if not hasattr(ph, n)])")
    assert p.returncode == 0, p.stderr[-2000:]
    assert "MISSING []" in p.stdout, p.stdout
