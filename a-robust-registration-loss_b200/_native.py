"""ctypes binding of librrl_b200.so (include/rrl_b200.h).  There is NO fallback: if the CUDA library is
missing or fails to load, importing any op raises -- the product path never routes through oracle/ or
through eager PyTorch."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# RRL_LIB_PATH: an A/B build of the same library (build.py, RRL_VARIANT) for the measurement tools
LIB_PATH = os.environ.get("RRL_LIB_PATH") or os.path.join(_HERE, "librrl_b200.so")

HIT_CAP = 5
NSTAT = 8
STATUS_EMPTY, STATUS_NAN, STATUS_NAN_RISK, STATUS_COMM = 1, 2, 4, 8
REUSE_ORDER = 1                  # RRL_REUSE_ORDER flag of rrl_loss_forward_ex / rrl_shard_stage1_ex
REUSE_TARGET = 2                 # RRL_REUSE_TARGET: cloud 2 unchanged since the previous forward in the workspace

_lib = None


class NativeError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NativeError(
            "librrl_b200.so is not built (%s). Run `python a-robust-registration-loss_b200/build.py` "
            "(or __graft_entry__.build()); there is no CPU fallback." % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, ci, cl, cz, cull = C.c_void_p, C.c_int, C.c_longlong, C.c_size_t, C.c_ulonglong
    sig = {
        "rrl_version": (ci, []),
        "rrl_error_string": (C.c_char_p, [ci]),
        "rrl_launch_count": (cl, []),
        "rrl_workspace_bytes": (cz, [ci, ci, ci, ci]),
        "rrl_loss_forward": (ci, [vp, vp, vp, ci, ci, ci, ci, ci, ci, ci, ci, vp, cz, vp, vp, vp, vp, vp]),
        "rrl_loss_forward_ex": (ci, [vp, vp, vp, ci, ci, ci, ci, ci, ci, ci, ci, vp, cz, vp, vp, vp, vp, ci, vp]),
        "rrl_shard_stage1_ex": (ci, [vp, vp, vp, ci, ci, ci, ci, ci, ci, ci, vp, cz, ci, vp]),
        "rrl_loss_backward": (ci, [vp, cz, vp, ci, ci, ci, ci, vp, vp, vp]),
        "rrl_loss_export_hits": (ci, [vp, cz, ci, ci, ci, ci, ci, vp, vp, vp]),
        "rrl_loss_backward_twist": (ci, [vp, cz, vp, ci, ci, ci, ci, vp, vp, vp, vp, vp]),
        "rrl_se3_chain": (ci, [vp, vp, ci, vp, vp]),
        "rrl_shard_stage1": (ci, [vp, vp, vp, ci, ci, ci, ci, ci, ci, ci, vp, cz, vp]),
        "rrl_shard_counts": (ci, [vp, cz, ci, ci, ci, vp, vp]),
        "rrl_shard_pack_entries": (ci, [vp, cz, ci, ci, ci, vp, cl, vp]),
        "rrl_select_lower_median": (ci, [vp, cl, vp, vp]),
        "rrl_shard_select_hist": (ci, [vp, cz, ci, ci, ci, ci, vp, vp, vp]),
        "rrl_shard_select_pick": (ci, [ci, vp, vp, vp, vp, vp]),
        "rrl_shard_stage2": (ci, [vp, cz, ci, ci, ci, vp, vp, vp, vp]),
        "rrl_shard_stage3": (ci, [vp, cz, ci, ci, ci, vp, vp, vp, vp]),
        "rrl_comm_create": (ci, [ci, ci, cz, C.POINTER(vp)]),
        "rrl_comm_destroy": (None, [vp]),
        "rrl_comm_slot_bytes": (cz, [vp]),
        "rrl_comm_local_base": (vp, [vp]),
        "rrl_comm_ipc_handle": (ci, [vp, vp]),
        "rrl_comm_connect_ipc": (ci, [vp, vp]),
        "rrl_comm_connect_ptrs": (ci, [vp, C.POINTER(vp)]),
        "rrl_comm_error": (ci, [vp]),
        "rrl_comm_allreduce_f64": (ci, [vp, vp, ci, vp]),
        "rrl_shard_tail": (ci, [vp, cz, ci, ci, ci, vp, vp, vp, vp, vp, vp]),
        "rrl_se3_exp": (ci, [vp, ci, vp, vp, vp]),
        "rrl_se3_exp4": (ci, [vp, ci, vp, vp]),
        "rrl_se3_expmap_backward": (ci, [vp, vp, ci, vp, vp]),
        "rrl_se3_apply": (ci, [vp, vp, ci, ci, vp, vp]),
        "rrl_se3_apply_backward": (ci, [vp, vp, vp, ci, ci, vp, vp, vp]),
        "rrl_rigid_apply": (ci, [vp, vp, vp, ci, ci, vp, vp]),
        "rrl_rigid_apply_backward": (ci, [vp, vp, vp, ci, ci, vp, vp, vp, vp, vp]),
        "rrl_sampler_workspace_bytes": (cz, [ci, ci, ci]),
        "rrl_sample_lines": (ci, [vp, vp, vp, vp, ci, ci, ci, ci, ci, cull, cull, vp, vp, vp, vp, cz, vp]),
        "rrl_sampler_num_chunks": (ci, [ci, ci]),
        "rrl_sample_lines_shard_flags": (ci, [vp, vp, vp, vp, ci, ci, ci, ci, ci, cull, cull, vp, ci, ci, vp, vp, cz, vp]),
        "rrl_sample_lines_shard_scatter": (ci, [vp, vp, ci, ci, ci, cull, cull, vp, ci, ci, vp, vp, vp, vp, vp, cz, vp]),
        "rrl_chamfer": (ci, [vp, vp, ci, ci, ci, vp, vp, vp]),
        "rrl_chamfer_forward": (ci, [vp, vp, ci, ci, ci, vp, vp, vp, vp]),
        "rrl_chamfer_backward": (ci, [vp, vp, vp, vp, ci, ci, ci, vp, vp, vp]),
        "rrl_fps_workspace_bytes": (cz, [ci]),
        "rrl_fps": (ci, [vp, ci, ci, ci, ci, vp, vp, cz, vp]),
        "rrl_knn": (ci, [vp, ci, ci, vp, ci, ci, vp, vp]),
        "rrl_host_create": (ci, [ci, ci, ci, ci, ci, C.POINTER(vp)]),
        "rrl_host_destroy": (None, [vp]),
        "rrl_host_pinned_tri1": (vp, [vp]),
        "rrl_host_pinned_tri2": (vp, [vp]),
        "rrl_host_pinned_lines": (vp, [vp]),
        "rrl_host_subbatches": (ci, [vp]),
        "rrl_host_loss_fwd_bwd": (ci, [vp, vp, vp, vp, ci, ci, ci, ci, vp, vp, vp]),
        "rrl_host_slots": (ci, [vp]),
        "rrl_host_submit": (ci, [vp, vp, vp, vp, ci, ci, ci, ci, ci, C.POINTER(ci)]),
        "rrl_host_wait": (ci, [vp, ci, vp, vp, vp]),
        "rrl_measure_fp32_peak": (ci, [ci, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
        "rrl_measure_dense": (ci, [vp, vp, vp, ci, ci, ci, ci, vp, cz, ci, C.POINTER(C.c_float), C.POINTER(C.c_float), vp]),
        "rrl_measure_stages": (ci, [vp, vp, vp, ci, ci, ci, ci, vp, cz, ci, C.POINTER(C.c_float), vp]),
        "rrl_debug_set_dense_variant": (ci, [ci]),
        "rrl_debug_set_param": (ci, [ci, ci]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)          # AttributeError here == header and library out of sync: fail loudly
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


EXPORTED = ["rrl_version", "rrl_error_string", "rrl_launch_count", "rrl_workspace_bytes", "rrl_loss_forward",
            "rrl_loss_forward_ex", "rrl_shard_stage1_ex",
            "rrl_loss_backward", "rrl_loss_backward_twist", "rrl_se3_chain", "rrl_loss_export_hits", "rrl_shard_stage1", "rrl_shard_counts",
            "rrl_shard_pack_entries", "rrl_select_lower_median", "rrl_shard_select_hist", "rrl_shard_select_pick",
            "rrl_shard_stage2", "rrl_shard_stage3", "rrl_shard_tail",
            "rrl_comm_create", "rrl_comm_destroy", "rrl_comm_slot_bytes", "rrl_comm_local_base", "rrl_comm_ipc_handle",
            "rrl_comm_connect_ipc", "rrl_comm_connect_ptrs", "rrl_comm_error", "rrl_comm_allreduce_f64",
            "rrl_se3_exp", "rrl_se3_exp4", "rrl_se3_expmap_backward", "rrl_se3_apply", "rrl_se3_apply_backward", "rrl_rigid_apply", "rrl_rigid_apply_backward",
            "rrl_sampler_workspace_bytes", "rrl_sample_lines", "rrl_sampler_num_chunks", "rrl_sample_lines_shard_flags", "rrl_sample_lines_shard_scatter", "rrl_chamfer", "rrl_chamfer_forward", "rrl_chamfer_backward", "rrl_fps_workspace_bytes", "rrl_fps", "rrl_knn", "rrl_host_create", "rrl_host_destroy",
            "rrl_host_pinned_tri1", "rrl_host_pinned_tri2", "rrl_host_pinned_lines", "rrl_host_subbatches", "rrl_host_loss_fwd_bwd",
            "rrl_host_slots", "rrl_host_submit", "rrl_host_wait",
            "rrl_measure_fp32_peak", "rrl_measure_dense", "rrl_measure_stages", "rrl_debug_set_dense_variant",
            "rrl_debug_set_param"]


def check(rc, what=""):
    if rc != 0:
        raise NativeError("%s failed: %s (code %d)" % (what or "rrl call", lib().rrl_error_string(rc).decode(), rc))


def launch_count() -> int:
    return int(lib().rrl_launch_count())
