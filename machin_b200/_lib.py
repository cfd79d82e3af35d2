"""ctypes binding of libmachin_b200.so (the C ABI declared in include/machin_b200.h).

The library is the product: there is no Python/NumPy/torch fallback for any operation it
exports.  A missing library raises ImportError at import; a missing CUDA device raises
RuntimeError the first time a handle is created.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmachin_b200.so")

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        "(or `make -C machin_b200/csrc`).  machin_b200 has no CPU fallback."
    )

lib = C.CDLL(LIB_PATH)

MB_OK = 0
MB_ERR_CUDA = -1
MB_ERR_INDEX_OOB = -2
MB_ERR_DIM_MISMATCH = -3
MB_ERR_SIZE_MISMATCH = -4
MB_ERR_INVALID = -5
MB_ERR_NOMEM = -6
MB_ERR_NO_DEVICE = -7
MB_F32, MB_F64 = 0, 1
MB_MAX_FIELDS = 16
MB_MAX_SOFT_TENSORS = 24          # include/machin_b200.h
MB_MAX_PEERS = 16

i64, i32, f64, f32, vp = C.c_int64, C.c_int, C.c_double, C.c_float, C.c_void_p


class MbField(C.Structure):
    _fields_ = [("row_offset", i64), ("nbytes", i64), ("d_dense", vp), ("dense_pitch", i64)]


class MbPerParams(C.Structure):
    _fields_ = [("beta", f64), ("all_weight_sum", f64), ("n_stored", i64), ("step", i64), ("n_draw", i64)]


PER_PARAMS_BYTES = C.sizeof(MbPerParams)      # 40: the device-side block every graph-mode draw reads


# name -> (restype, argtypes); every symbol declared in include/machin_b200.h
SIGNATURES = {
    "mb_abi_version": (i32, []),
    "mb_status_string": (C.c_char_p, [i32]),
    "mb_last_cuda_error": (C.c_char_p, []),
    "mb_device_count": (i32, []),
    "mb_launch_count": (C.c_longlong, []),
    "mb_tree_create": (i32, [i64, i32, C.POINTER(vp)]),
    "mb_tree_destroy": (i32, [vp]),
    "mb_tree_geometry": (i32, [vp, C.POINTER(i64), C.POINTER(i32), C.POINTER(i64), C.POINTER(i64)]),
    "mb_tree_weights": (vp, [vp]),
    "mb_tree_stats": (i32, [vp, C.POINTER(f64), C.POINTER(f64), C.POINTER(i32), vp]),
    "mb_tree_root_ptr": (vp, [vp]),
    "mb_tree_max_leaf_ptr": (vp, [vp]),
    "mb_tree_state_blob": (vp, [vp, C.POINTER(i64)]),
    "mb_tree_check": (i32, [vp, vp]),
    "mb_tree_clear": (i32, [vp, vp]),
    "mb_tree_build": (i32, [vp, vp, i64, vp]),
    "mb_tree_find": (i32, [vp, vp, i64, vp, vp]),
    "mb_tree_leaf_weight": (i32, [vp, vp, i64, vp, vp]),
    "mb_tree_update_batch": (i32, [vp, vp, vp, i32, vp, i64, vp]),
    "mb_tree_update_priority": (i32, [vp, vp, vp, i32, vp, i64, f64, f64, vp]),
    "mb_tree_update_leaf": (i32, [vp, f64, i64, vp]),
    "mb_per_sample": (i32, [vp, vp, i64, i64, f64, f64, vp, vp, vp, vp]),
    "mb_per_sample_multi": (i32, [vp, vp, i64, i32, vp, vp, vp, vp]),
    "mb_tree_build_host": (i32, [vp, vp, i64]),
    "mb_tree_find_host": (i32, [vp, vp, i64, vp]),
    "mb_tree_leaf_weight_host": (i32, [vp, vp, i64, vp]),
    "mb_tree_update_batch_host": (i32, [vp, vp, vp, i32, i64]),
    "mb_tree_read_host": (i32, [vp, i64, i64, vp]),
    "mb_per_sample_host": (i32, [vp, vp, i64, i64, f64, f64, vp, vp]),
    "mb_rows_gather": (i32, [vp, i64, i64, vp, i64, C.POINTER(MbField), i32, vp]),
    "mb_rows_scatter": (i32, [vp, i64, i64, i64, i64, C.POINTER(MbField), i32, vp]),
    "mb_rows_gather_packed": (i32, [vp, i64, i64, vp, i64, vp, vp]),
    "mb_rows_pack_meta": (i32, [vp, i64, i64, i64, vp, vp, vp, vp, vp, vp]),
    "mb_rows_unpack_meta": (i32, [vp, i64, i64, i64, vp, vp, vp, vp, vp]),
    "mb_wire_compact": (i32, [vp, i32, i64, i64, i64, vp, vp, i64, vp, vp, vp, vp, vp]),
    "mb_wire_scatter_update": (i32, [vp, i32, i64, vp, vp, vp, vp, i64, vp]),
    "mb_wire_apportion": (i32, [vp, i32, i64, i32, i64, vp, vp, f64, vp, vp, vp, vp]),
    "mb_wire_update_mask": (i32, [vp, i32, i64, vp, i32, vp, i64, vp, vp, vp, vp]),
    "mb_wire_push_rows": (i32, [vp, i64, i64, vp, vp, vp, i32, i64, i32, i64, C.POINTER(vp), C.POINTER(vp), vp, vp, vp]),
    "mb_wire_push_update": (i32, [i32, i64, i32, vp, vp, vp, vp, i64, C.POINTER(vp), vp]),
    "mb_peer_push": (i32, [vp, i64, i32, i32, i64, C.POINTER(vp), C.POINTER(vp), vp, vp, vp]),
    "mb_peer_sum": (i32, [vp, i32, i64, i64, f32, vp, C.POINTER(vp), i32, vp, vp]),
    "mb_wire_compact_wait": (i32, [vp, i32, i64, i64, i64, vp, vp, i64, vp, vp, vp, vp, C.POINTER(vp), i32, vp, vp]),
    "mb_wire_compact_fields": (i32, [vp, i32, i64, i64, i64, vp, C.POINTER(MbField), i32, i64, vp, vp, vp, vp, C.POINTER(vp), i32, vp, vp]),
    "mb_peer_sums_apportion": (i32, [C.POINTER(vp), C.POINTER(vp), vp, i32, i32, vp, i64, i64, i64, vp, vp, f64, vp, vp, vp, vp]),
    "mb_peer_barrier": (i32, [C.POINTER(vp), i32, i32, vp, i64, vp]),
    "mb_td_dqn": (i32, [vp, vp, vp, i64, i32, vp, vp, vp, vp, f32, i32, vp, vp, vp, vp, vp, vp]),
    "mb_td_target": (i32, [vp, vp, vp, i64, f32, vp, vp, vp, vp]),
    "mb_td_rainbow": (i32, [vp, vp, vp, i64, i32, i32, vp, vp, vp, vp, vp, vp, f64, f64, f64, vp, vp, vp, vp, vp, vp]),
    "mb_soft_update": (i32, [vp, vp, i64, f64, vp]),
    "mb_soft_update_multi": (i32, [C.POINTER(vp), C.POINTER(vp), C.POINTER(i64), i32, f64, vp]),
    "mb_mlp_linear_forward": (i32, [vp, vp, vp, vp, i32, i32, i32, i32, vp, i64, vp, i32, vp]),
    "mb_mlp_linear_backward_input": (i32, [vp, vp, vp, vp, i32, i32, i32, vp, i64, vp, i32, vp]),
    "mb_mlp_linear_backward_weight": (i32, [vp, vp, vp, vp, i32, i32, i32, vp, i64, vp, i32, vp]),
    "mb_dqn_mlp_program_create": (i32, [i32, i32, C.POINTER(i32), i32, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), vp, vp,
                                        vp, vp, vp, vp, f32, i32, vp, vp, vp, vp, C.POINTER(vp)]),
    "mb_dqn_mlp_rows_create": (i32, [i32, i32, C.POINTER(i32), i32, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), vp, vp,
                                     vp, vp, vp, vp, f32, i32, vp, vp, vp, vp, C.POINTER(vp)]),
    "mb_mlp_program_run": (i32, [vp, vp]),
    "mb_mlp_program_run_part": (i32, [vp, i32, vp]),
    "mb_mlp_program_destroy": (i32, [vp]),
    "mb_adam_step_multi": (i32, [C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(i64), i32, vp,
                                 f64, f64, f64, f64, f64, vp]),
    "mb_adam_soft_step_multi": (i32, [C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(vp),
                                      C.POINTER(i64), i32, vp, vp, f64, f64, f64, f64, f64, f64, vp]),
    "mb_vtrace": (i32, [vp, i64, i64, vp, vp, vp, vp, vp, vp, f32, f32, f32, vp, vp, vp, vp, vp]),
    "mb_vtrace_len": (i32, [vp, i64, i64, i64, vp, vp, vp, vp, vp, vp, f32, f32, f32, vp, vp, vp, vp, vp]),
    "mb_gae": (i32, [vp, i64, i64, vp, vp, vp, vp, f64, f64, vp, vp, vp]),
    "mb_gae_len": (i32, [vp, i64, i64, i64, vp, vp, vp, vp, f64, f64, vp, vp, vp]),
    "mb_vtrace_host": (i32, [i32, vp, i64, i64, vp, vp, vp, vp, vp, vp, f32, f32, f32, vp, vp, vp, vp]),
    "mb_gae_host": (i32, [i32, vp, i64, i64, vp, vp, vp, vp, f64, f64, vp, vp]),
}

for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(lib, _name)  # AttributeError here == the library does not match the header
    _fn.restype = _res
    _fn.argtypes = _args

if lib.mb_abi_version() != 1:
    raise ImportError("libmachin_b200.so ABI version mismatch")

_MESSAGES = {
    MB_ERR_INDEX_OOB: (ValueError, "Index has elements out of boundary!"),
    MB_ERR_DIM_MISMATCH: (ValueError, "Dimension of weights and indexes doesn't match!"),
    MB_ERR_SIZE_MISMATCH: (ValueError, "Weights size must match tree size!"),
}


def check(status: int):
    """Map an mb_status to the exact exception the reference raises."""
    if status == MB_OK:
        return
    if status in _MESSAGES:
        exc, msg = _MESSAGES[status]
        raise exc(msg)
    text = lib.mb_status_string(status).decode()
    if status == MB_ERR_CUDA:
        text += ": " + lib.mb_last_cuda_error().decode()
    raise RuntimeError(f"libmachin_b200: {text}")


def require_device(index: int = 0):
    n = lib.mb_device_count()
    if n <= index:
        raise RuntimeError(
            f"machin_b200 needs CUDA device {index} but {n} device(s) are visible; "
            "there is no CPU fallback for this path."
        )


def stream_ptr(stream=None):
    """cudaStream_t of a torch stream (default: torch's current stream) as a void*."""
    import torch

    s = stream if stream is not None else torch.cuda.current_stream()
    return vp(s.cuda_stream)


def ptr(tensor):
    """Device/host address of a torch tensor or NumPy array (None -> NULL)."""
    if tensor is None:
        return vp(0)
    if hasattr(tensor, "data_ptr"):
        return vp(tensor.data_ptr())
    return vp(tensor.ctypes.data)
