"""ctypes binding of libmaus_b200.so (include/maus_b200.h).  No CPU fallback: a missing library is an ImportError-
class failure (``MausError``), a missing GPU makes ``maus_create`` fail."""
import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_NAME = "libmaus_b200.so"

# status words / enums of include/maus_b200.h
ST_OK, ST_ZERO_PIVOT, ST_NONFINITE, ST_GMRES_NOCONV, ST_V_COLLAPSED, ST_MIX_COLLAPSED, ST_SKIPPED = range(7)
EIGENVALUE, SOLVE_LINEAR_SYSTEM, SVD = 1, 2, 3
METHOD_LU, METHOD_GMRES = 0, 1
SLOT_CURRENT, SLOT_CTOR = 0, 1

EXPORTS = [
    "maus_create", "maus_destroy", "maus_last_error", "maus_set_workspace_limit", "maus_info", "maus_alloc_pinned",
    "maus_free_pinned", "maus_set_dense", "maus_add_dense_form", "maus_set_csc", "maus_set_rhs", "maus_upload_vectors",
    "maus_download_vectors", "maus_download_vector_range", "maus_rq", "maus_solve_shifted", "maus_solve_with_R", "maus_mix_residual",
    "maus_residual", "maus_step", "maus_launch_count", "maus_profile_reset", "maus_profile_read", "maus_profile_read_kind", "maus_stream", "maus_debug_zgemm", "maus_svd_set_matrix", "maus_svd_step", "maus_svd_residual",
    "maus_gram", "maus_project", "maus_heev", "maus_diag_dense", "maus_cond2_estimate", "maus_nccl_unique_id", "maus_dist_init", "maus_dist_info", "maus_gather", "maus_set_csr_rowblock",
    "maus_rs_set_rhs", "maus_rs_matvec", "maus_rs_gmres", "maus_rs_step",
]


class MausError(RuntimeError):
    pass


def library_path():
    return os.path.join(HERE, LIB_NAME)


def build_library(verbose=False):
    """Compile libmaus_b200.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    r = subprocess.run(["make", "-C", HERE, "-j8"], capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout[-4000:])
        print(r.stderr[-4000:])
    if r.returncode != 0:
        raise MausError("building libmaus_b200.so failed")
    return library_path()


_lib = None


def load_library():
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.isfile(path):
        raise MausError(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                        f"(there is no CPU fallback)")
    lib = C.CDLL(path)
    vp, i32, i64, dp = C.c_void_p, C.c_int, C.c_int64, C.POINTER(C.c_double)
    u8p, u64p, i32p, i64p = C.POINTER(C.c_uint8), C.POINTER(C.c_uint64), C.POINTER(C.c_int32), C.POINTER(C.c_int64)
    sig = {
        "maus_create": (i32, [C.POINTER(vp), i32]),
        "maus_destroy": (i32, [vp]),
        "maus_last_error": (C.c_char_p, [vp]),
        "maus_set_workspace_limit": (i32, [vp, i64]),
        "maus_info": (i32, [vp, i32p, i32p, i64p]),
        "maus_alloc_pinned": (vp, [i64]),
        "maus_free_pinned": (None, [vp]),
        "maus_set_dense": (i32, [vp, i32, i64, dp]),
        "maus_add_dense_form": (i32, [vp, i64, dp]),
        "maus_set_csc": (i32, [vp, i32, i64, i64, i64p, i64p, dp]),
        "maus_set_rhs": (i32, [vp, dp]),
        "maus_upload_vectors": (i32, [vp, i64, dp]),
        "maus_download_vectors": (i32, [vp, i64, dp]),
        "maus_download_vector_range": (i32, [vp, i64, i64, dp]),
        "maus_rq": (i32, [vp, i64, dp, dp, dp]),
        "maus_solve_shifted": (i32, [vp, i64, dp, dp, u64p, i32, u8p, dp, i32, dp, i32p, i32p]),
        "maus_solve_with_R": (i32, [vp, dp, dp, dp, dp, dp, i32p]),
        "maus_mix_residual": (i32, [vp, i64, i32, dp, dp, u8p, i32, dp, dp, dp, i32p]),
        "maus_residual": (i32, [vp, i64, i32, dp, dp, i32, dp]),
        "maus_step": (i32, [vp, i64, i32, i32, dp, dp, dp, u64p, u8p, i32, dp, dp, dp, i32p, i32p]),
        "maus_launch_count": (i64, [vp]),
        "maus_profile_reset": (i32, [vp, i32]),
        "maus_profile_read": (i32, [vp, dp, i64p, dp, dp, i64p, dp]),
        "maus_profile_read_kind": (i32, [vp, i32, dp, i64p, dp]),
        "maus_stream": (vp, [vp]),
        "maus_debug_zgemm": (i32, [vp, i32, i32, i32, i32, dp, dp, dp, i32, i32, i32]),
        "maus_svd_set_matrix": (i32, [vp, i64, i64, dp]),
        "maus_svd_step": (i32, [vp, i64, dp, dp, dp, dp, i32p]),
        "maus_svd_residual": (i32, [vp, i64, dp, dp, dp, dp]),
        "maus_gram": (i32, [vp, i64, i64, dp, dp]),
        "maus_project": (i32, [vp, i64, i64, dp, i64, dp, dp]),
        "maus_heev": (i32, [vp, i64, dp, i32, dp, dp, i32p, dp]),
        "maus_diag_dense": (i32, [vp, C.c_double, C.c_double, i64p, i32p, i32p]),
        "maus_cond2_estimate": (i32, [vp, i32, i32, dp, dp, dp, i32p]),
        "maus_nccl_unique_id": (i32, [C.c_char_p, C.c_char_p]),
        "maus_dist_init": (i32, [vp, C.c_char_p, i32, i32, C.c_char_p]),
        "maus_set_csr_rowblock": (i32, [vp, i64, i64, i64, i64p, i64p, dp]),
        "maus_dist_info": (i32, [vp, i32p, i32p, i32p]),
        "maus_gather": (i32, [vp, dp, i64, dp]),
        "maus_rs_set_rhs": (i32, [vp, dp]),
        "maus_rs_step": (i32, [vp, i64, i32, i32, dp, dp, dp, u8p, dp, dp, dp, dp, i32p, i32p]),
        "maus_rs_matvec": (i32, [vp, i64, dp, dp]),
        "maus_rs_gmres": (i32, [vp, i64, dp, dp, u8p, dp, dp, i32p, i32p]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)      # AttributeError if the .so does not export it
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
