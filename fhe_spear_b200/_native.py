"""ctypes binding of libspear_b200.so (include/spear_b200.h).

The library is the product: there is no Python or CPU fallback.  If it is missing or
cannot be loaded, importing this module raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libspear_b200.so")

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        "(or `make -C fhe_spear_b200/csrc`). The CUDA library is mandatory; there is no CPU fallback.")

lib = C.CDLL(LIB_PATH)

vp = C.c_void_p
vpp = C.POINTER(C.c_void_p)
u64p = C.POINTER(C.c_uint64)
u32p = C.POINTER(C.c_uint32)
f64p = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int)

_SIG = {
    "spear_last_error": (C.c_char_p, []),
    "spear_version": (C.c_char_p, []),
    "spear_launch_count": (C.c_uint64, []),
    "spear_create_coeff_modulus": (C.c_int, [C.c_uint64, ip, C.c_int, u64p]),
    "spear_get_elt_from_step": (C.c_uint64, [C.c_int, C.c_uint64]),
    "spear_context_create": (C.c_int, [C.c_uint64, u64p, C.c_int, C.c_int, C.c_int, vpp]),
    "spear_context_destroy": (None, [vp]),
    "spear_context_sync": (C.c_int, [vp]),
    "spear_context_stream": (vp, [vp]),
    "spear_timer_start": (C.c_int, [vp]),
    "spear_timer_stop": (C.c_int, [vp, C.POINTER(C.c_float)]),
    "spear_profile_enable": (C.c_int, [vp, C.c_int]),
    "spear_profile_read": (C.c_int, [vp, f64p, u64p, C.c_int]),
    "spear_pinned_alloc": (C.c_int, [C.c_size_t, vpp]),
    "spear_pinned_free": (None, [vp]),
    "spear_mem_info": (C.c_int, [vp, u64p, u64p]),
    "spear_mem_reserve": (C.c_int, [vp, C.c_uint64]),
    "spear_secret_key_create": (C.c_int, [vp, C.c_char_p, vpp]),
    "spear_secret_key_destroy": (None, [vp]),
    "spear_gen_public_key": (C.c_int, [vp, vp, vpp]),
    "spear_public_key_destroy": (None, [vp]),
    "spear_gen_relin_key": (C.c_int, [vp, vp, vpp]),
    "spear_kswitch_key_destroy": (None, [vp]),
    "spear_gen_galois_keys": (C.c_int, [vp, vp, u32p, C.c_int, vpp]),
    "spear_galois_keys_add": (C.c_int, [vp, vp, vp, u32p, C.c_int]),
    "spear_galois_keys_has": (C.c_int, [vp, C.c_uint32]),
    "spear_galois_keys_destroy": (None, [vp]),
    "spear_obj_destroy": (None, [vp]),
    "spear_obj_info": (C.c_int, [vp, ip, ip, ip, ip, f64p, ip]),
    "spear_obj_set_scale": (C.c_int, [vp, C.c_double]),
    "spear_obj_export": (C.c_int, [vp, vp, C.c_size_t]),
    "spear_obj_import": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, vpp]),
    "spear_secret_key_export": (C.c_int, [vp, vp, C.c_size_t]),
    "spear_kswitch_key_export": (C.c_int, [vp, vp, C.c_size_t]),
    "spear_galois_key_export": (C.c_int, [vp, C.c_uint32, vp, C.c_size_t]),
    "spear_public_key_export": (C.c_int, [vp, vp, C.c_size_t]),
    "spear_encode": (C.c_int, [vp, f64p, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, vpp]),
    "spear_decode": (C.c_int, [vp, vp, f64p]),
    "spear_encrypt_symmetric": (C.c_int, [vp, vp, vp, C.c_uint64, vpp]),
    "spear_encrypt_asymmetric": (C.c_int, [vp, vp, vp, C.c_uint64, vpp]),
    "spear_encrypt_vector": (C.c_int, [vp, vp, f64p, C.c_int, C.c_int, C.c_double, C.c_uint64, vpp]),
    "spear_decrypt_decode": (C.c_int, [vp, vp, vp, f64p, C.c_int]),
    "spear_decrypt": (C.c_int, [vp, vp, vp, vpp]),
    "spear_negate": (C.c_int, [vp, vp, vpp]),
    "spear_add": (C.c_int, [vp, vp, vp, vpp]),
    "spear_sub": (C.c_int, [vp, vp, vp, vpp]),
    "spear_add_plain": (C.c_int, [vp, vp, vp, vpp]),
    "spear_sub_plain": (C.c_int, [vp, vp, vp, vpp]),
    "spear_multiply": (C.c_int, [vp, vp, vp, vpp]),
    "spear_multiply_plain": (C.c_int, [vp, vp, vp, vpp]),
    "spear_relinearize": (C.c_int, [vp, vp, vp, vpp]),
    "spear_rescale_to_next": (C.c_int, [vp, vp, vpp]),
    "spear_mod_raise": (C.c_int, [vp, vp, C.c_int, vpp]),
    "spear_mod_switch_to_next": (C.c_int, [vp, vp, vpp]),
    "spear_apply_galois": (C.c_int, [vp, vp, C.c_uint32, vp, vpp]),
    "spear_hoisted_rotations": (C.c_int, [vp, vp, u32p, C.c_int, vp, vpp]),
    "spear_bsgs_multiply_accumulate": (C.c_int, [vp, vpp, C.c_int, vpp, C.c_int, C.c_int, C.c_int, C.c_int, vp, vpp]),
    "spear_bsgs_from_host": (C.c_int, [vp, vpp, C.c_int, vp, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int, vp, vpp]),
    "spear_objs_export": (C.c_int, [vp, vpp, C.c_int, vp, C.c_size_t]),
    "spear_diagset_encode": (C.c_int, [vp, f64p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, vpp]),
    "spear_diagset_encode_matrix": (C.c_int, [vp, f64p, f64p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double,
                                               C.c_int, C.c_int, vpp]),
    "spear_diagset_encode_matrix_view": (C.c_int, [vp, f64p, f64p, C.c_int, C.c_int, C.c_int, C.c_size_t, C.c_int, C.c_int,
                                                    C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, vpp]),
    "spear_diagset_encode_shard": (C.c_int, [vp, f64p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double,
                                            C.c_int, C.c_int, vpp]),
    "spear_diagset_destroy": (None, [vp]),
    "spear_diagset_info": (C.c_int, [vp, ip, ip, ip, ip, ip, f64p, u64p]),
    "spear_diagset_export": (C.c_int, [vp, vp, C.c_size_t]),
    "spear_bsgs_hoisted": (C.c_int, [vp, vp, vp, vp, vpp]),
    "spear_bsgs_hoisted_batch": (C.c_int, [vp, vpp, vpp, C.c_int, vp, vpp]),
    "spear_bsgs_hoisted_shared": (C.c_int, [vp, vp, vpp, C.c_int, vp, vpp]),
    "spear_bsgs_hoisted_batch_host": (C.c_int, [vp, vpp, C.c_int, C.c_double, vpp, C.c_int, vp, vpp, f64p]),
    "spear_bsgs_hoisted_partial": (C.c_int, [vp, vp, vp, vp, vpp]),
    "spear_bsgs_hoisted_partial_batch": (C.c_int, [vp, vpp, vpp, C.c_int, vp, vpp]),
    "spear_bsgs_finish": (C.c_int, [vp, vp, vpp]),
    "spear_obj_reduce": (C.c_int, [vp, vp]),
    "spear_obj_device_ptr": (vp, [vp]),
    "spear_peer_window_create": (C.c_int, [vp, C.c_int, C.c_int, C.c_uint64, C.c_int, C.c_char_p, vpp]),
    "spear_peer_window_connect": (C.c_int, [vp, vp, C.c_char_p]),
    "spear_peer_allreduce": (C.c_int, [vp, vp, C.c_int, vp]),
    "spear_peer_window_status": (C.c_int, [vp]),
    "spear_peer_window_destroy": (None, [vp]),
    "spear_peer_selftest": (C.c_int, [vp, vpp, C.c_int]),
    "spear_diagset_slice_rows": (C.c_int, [vp, vp, C.c_int, C.c_int, vpp]),
    "spear_diagset_slice_share": (C.c_int, [vp, vp, C.c_int, C.c_int, vpp]),
    "spear_split_share": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, ip, ip, ip, ip]),
    "spear_bsgs_split": (C.c_int, [vp, vp, vp, vp, vp, C.c_int, vpp]),
    "spear_bsgs_split_batch": (C.c_int, [vp, vpp, vpp, C.c_int, vp, vp, C.c_int, vpp]),
    "spear_bsgs_split_shared": (C.c_int, [vp, vp, vpp, C.c_int, vp, vp, C.c_int, vpp]),
    "spear_bsgs_split_selftest": (C.c_int, [vp, vp, vpp, C.c_int, vp, vpp]),
    "spear_ntt_host": (C.c_int, [vp, vp, C.c_int, ip, C.c_int, C.c_int]),
}

for _name, (_res, _args) in _SIG.items():
    _f = getattr(lib, _name)   # AttributeError here = library does not export a declared symbol
    _f.restype = _res
    _f.argtypes = _args

EXPORTED = tuple(_SIG)


def check(rc):
    """Turn a C status into the Python exception the reference scripts expect (RuntimeError)."""
    if rc:
        raise RuntimeError(lib.spear_last_error().decode("utf-8", "replace"))


def launch_count():
    return int(lib.spear_launch_count())
