"""ctypes binding of libb200groth16.so - the same symbols the Go shim binds with cgo
(davinci-node_b200/go/prover_b200.go).  Every call raises B200Error on a non-zero status; nothing
here computes on the CPU."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B200_LIB_PATH") or os.path.join(_HERE, "libb200groth16.so")


class B200Error(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise B200Error(
            "libb200groth16.so is missing (%s): build it with `python davinci-node_b200/build.py` - "
            "this backend has no CPU fallback" % LIB_PATH)
    return C.CDLL(LIB_PATH)


lib = _load()

_vp, _u64, _u32, _i = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int



class Slice(C.Structure):
    _fields_ = [("ptr", _vp), ("len", _u64)]


class PkDesc(C.Structure):
    _fields_ = [
        ("curve", _i), ("domain_size", _u64), ("generator", _vp), ("coset_gen", _vp),
        ("g1_alpha", _vp), ("g1_beta", _vp), ("g1_delta", _vp),
        ("g1_A", Slice), ("g1_B", Slice), ("g1_Z", Slice), ("g1_K", Slice),
        ("g2_beta", _vp), ("g2_delta", _vp), ("g2_B", Slice),
        ("infinity_a", Slice), ("infinity_b", Slice),
        ("nb_wires", _u64), ("nb_public", _u64), ("krs_skip", Slice),
        ("nb_commitments", _u32), ("commit_basis", C.POINTER(Slice)), ("commit_basis_exp_sigma", C.POINTER(Slice)),
        ("z_offset", _u64),
    ]


class ProveIn(C.Structure):
    _fields_ = [
        ("wires", Slice), ("a", Slice), ("b", Slice), ("c", Slice), ("r", _vp), ("s", _vp),
        ("nb_commitments", _u32), ("priv_committed", C.POINTER(Slice)), ("fold_challenge", _vp), ("abc_form", _u32),
    ]


class ProofOut(C.Structure):
    _fields_ = [("ar", _vp), ("bs", _vp), ("krs", _vp), ("pok", _vp)]


_PROTOS = {
    "b200_domain_create": (_i, [_i, _u64, _vp, _vp, C.POINTER(_u64)]),
    "b200_domain_release": (_i, [_u64]),
    "b200_ntt_dev": (_i, [_u64, _vp, _i, _i, _i, _vp]),
    "b200_compute_h_dev": (_i, [_u64, _vp, _vp, _vp, _vp]),
    "b200_pk_register": (_i, [C.POINTER(PkDesc), C.POINTER(_u64)]),
    "b200_pk_release": (_i, [_u64]),
    "b200_set_pk_table_budget": (_i, [_u64]),
    "b200_pk_info": (_i, [_u64, C.POINTER(_u64)]),
    "b200_commit": (_i, [_u64, _u32, Slice, _vp, _i]),
    "b200_prove": (_i, [_u64, C.POINTER(ProveIn), C.POINTER(ProofOut), _i]),
    "b200_prove_dev": (_i, [_u64, C.POINTER(ProveIn), C.POINTER(ProofOut), _i]),
    "b200_compressed_bytes": (_u64, [_i, _i]),
    "b200_points_decompress_dev": (_i, [_i, _i, _vp, _u64, _vp, _vp, _vp]),
    "b200_points_compress_dev": (_i, [_i, _i, _vp, _u64, _vp, _vp]),
    "b200_gt_bytes": (_u64, [_i]),
    "b200_pairing_check": (_i, [_i, _vp, _vp, _u32, C.POINTER(_i), _vp, _i]),
    "b200_pairing_check_batch": (_i, [_i, _vp, _vp, _u32, _u32, _vp, _i]),
    "b200_fixed_base_dev": (_i, [_i, _i, _vp, _vp, _u64, _vp, _vp]),
    "b200_bases_create_dev": (_i, [_i, _i, _vp, _u64, _i, C.POINTER(_u64), _vp]),
    "b200_bases_release": (_i, [_u64]),
    "b200_msm_bases_dev": (_i, [_u64, _vp, _u64, _vp, _vp, _vp]),
    "b200_sum_partials_dev": (_i, [_i, _i, _vp, _u32, _vp, _vp]),
    "b200_launch_count": (_u64, []),
    "b200_profile_enable": (_i, [_i]),
    "b200_profile_collect": (_i, [C.POINTER(C.c_double), C.POINTER(_u64)]),
    "b200_profile_timeline": (_i, [C.POINTER(C.c_double), _u64, C.POINTER(_u64)]),
    "b200_prove_partial_dev": (_i, [_u64, C.POINTER(ProveIn), _vp, _i]),
    "b200_pk_coset_evals_dev": (_i, [_u64, _vp, _i, _vp]),
    "b200_assemble_dev": (_i, [_i, _vp, _u32, _vp, _vp, _i, _vp, _vp]),
    "b200_kzg_srs_register": (_i, [_vp, _u32, C.POINTER(_u64)]),
    "b200_kzg_srs_release": (_i, [_u64]),
    "b200_blob_commit": (_i, [_u64, _vp, _vp, _i]),
    "b200_blob_proof": (_i, [_u64, _vp, _vp, _vp, _vp, _i]),
    "b200_kzg_srs_add_monomial": (_i, [_u64, _vp, _u32]),
    "b200_blob_cell_proofs": (_i, [_u64, _vp, _vp, _i]),
    "b200_init": (_i, [_u32]),
    "b200_host_register": (_i, [_vp, _u64]),
    "b200_host_unregister": (_i, [_vp]),
    "b200_device_count": (_i, []),
    "b200_last_error": (C.c_char_p, []),
    "b200_version": (C.c_char_p, []),
    "b200_fr_bytes": (_u64, [_i]),
    "b200_fp_bytes": (_u64, [_i]),
    "b200_affine_bytes": (_u64, [_i, _i]),
    "b200_xyzz_bytes": (_u64, [_i, _i]),
    "b200_msm": (_i, [_i, _i, _vp, _vp, _u64, _vp, _i]),
    "b200_msm_dev": (_i, [_i, _i, _vp, _vp, _u64, _vp, _i, _vp]),
    "b200_to_affine_dev": (_i, [_i, _i, _vp, _vp, _u32, _vp]),
    "b200_msm_plan": (_i, [_i, _u64, _i, C.POINTER(_u32)]),
    "b200_dbg_field_op_dev": (_i, [_i, _i, _i, _vp, _vp, _vp, _u64, _vp]),
    "b200_dbg_ec_op_dev": (_i, [_i, _i, _i, _vp, _vp, _vp, _u64, _vp]),
    "b200_calib_mul_dev": (_i, [_i, _i, _vp, _u64, _i, _vp]),
}

EXPORTS = sorted(_PROTOS)

for _name, (_res, _args) in _PROTOS.items():
    _fn = getattr(lib, _name)
    _fn.restype = _res
    _fn.argtypes = _args


def check(status):
    if status != 0:
        raise B200Error(lib.b200_last_error().decode("utf-8", "replace"))


_inited = False


def init(device_mask=0):
    global _inited
    check(lib.b200_init(device_mask))
    _inited = True


def init_once():
    if not _inited:
        init(int(os.environ.get("B200_DEVICE_MASK", "0"), 0))


def pk_info(handle):
    out = (_u64 * 6)()
    check(lib.b200_pk_info(handle, out))
    return dict(table_stride=out[0], table_bytes=out[1], slots=out[2], wire_window=out[3], z_window=out[4], gpus=out[5])


def msm_plan(curve, n, window_bits=0):
    out = (_u32 * 5)()
    check(lib.b200_msm_plan(curve, n, window_bits, out))
    return dict(c=out[0], nwin=out[1], nb=out[2], task=out[3], group=out[4])
