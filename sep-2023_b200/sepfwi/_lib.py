"""ctypes binding of libsepfwi.so (include/sepfwi.h).

The library is the product; there is no Python or CPU fallback.  Importing this
module never touches the GPU, but every compute entry raises RuntimeError when the
shared library is missing or no CUDA device is present.
"""
import ctypes as C
import os
import subprocess

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))     # .../sep-2023_b200
LIB_PATH = os.environ.get("SEPFWI_LIB") or os.path.join(_PKG, "libsepfwi.so")   # SEPFWI_LIB: alternate build (tuning experiments)

FIBER_EXX, FIBER_EZZ = 0, 1
FLAVOUR_CPML, FLAVOUR_SPONGE = 0, 1
MEM_HOST, MEM_DEVICE = 0, 1
T_PR, T_VX, T_VZ, T_ETT, T_EXX, T_EZZ, T_EXZ = range(7)


class Params(C.Structure):
    _fields_ = [("nz", C.c_int), ("nx", C.c_int), ("nPml", C.c_int), ("nPad", C.c_int), ("nSteps", C.c_int),
                ("dz", C.c_float), ("dx", C.c_float), ("dt", C.c_float), ("f0", C.c_float),
                ("fiber", C.c_int), ("flavour", C.c_int), ("max_batch", C.c_int), ("max_nrec", C.c_int),
                ("with_adjoint", C.c_int), ("kernels", C.c_int), ("ref_race_compat", C.c_int),
                ("reserved", C.c_int * 5)]


class Shot(C.Structure):
    _fields_ = [("zs", C.c_int), ("xs", C.c_int), ("nrec", C.c_int),
                ("zrec", C.c_void_p), ("xrec", C.c_void_p), ("stf", C.c_void_p),
                ("src_rxz", C.c_float), ("obs_ett", C.c_void_p), ("out", C.c_void_p * 7),
                ("gstf", C.c_void_p), ("weights", C.c_void_p),
                ("win_start", C.c_void_p), ("win_end", C.c_void_p), ("trace_weights", C.c_void_p),
                ("src_weight", C.c_float), ("src_updated", C.c_void_p)]


class DataOptions(C.Structure):
    """sepfwi_data_options (include/sepfwi.h): the data-side switches of para_file.json."""
    _fields_ = [("if_win", C.c_int), ("win_ratio", C.c_float), ("if_filter", C.c_int), ("filter", C.c_float * 4),
                ("if_cross_misfit", C.c_int), ("if_src_update", C.c_int), ("reserved", C.c_int * 6)]


# every symbol include/sepfwi.h declares (tests check the export list against the header)
SYMBOLS = ["sepfwi_last_error", "sepfwi_version", "sepfwi_create", "sepfwi_destroy", "sepfwi_set_model",
           "sepfwi_courant", "sepfwi_forward", "sepfwi_gradient", "sepfwi_cufd", "sepfwi_cufd_clear_cache",
           "sepfwi_ring_len", "sepfwi_ring_save", "sepfwi_ring_restore", "sepfwi_get_cpml",
           "sepfwi_launch_count", "sepfwi_last_timing", "sepfwi_set_profile", "sepfwi_get_profile",
           "sepfwi_kernel_name", "sepfwi_resident_launches", "sepfwi_forward_snapshots", "sepfwi_plan_resident", "sepfwi_plan_stream", "sepfwi_plan_backward",
           "sepfwi_last_misfit", "sepfwi_bytes_per_slot", "sepfwi_set_data_options", "sepfwi_condition"]
NKERNEL = 14

_lib = None


def build(verbose=False):
    """Compile libsepfwi.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    cmd = ["make", "-C", _PKG, LIB_PATH]
    if not verbose:
        cmd.insert(1, "-s")
    subprocess.check_call(cmd)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("libsepfwi.so is not built (%s); run `python -c 'import __graft_entry__ as g; g.build()'`. "
                               "There is no CPU fallback." % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        L.sepfwi_last_error.restype = C.c_char_p
        L.sepfwi_launch_count.restype = C.c_longlong
        L.sepfwi_create.argtypes = [C.POINTER(Params), C.c_int, C.POINTER(C.c_void_p)]
        L.sepfwi_destroy.argtypes = [C.c_void_p]
        L.sepfwi_set_model.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.sepfwi_courant.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
        L.sepfwi_forward.argtypes = [C.c_void_p, C.c_int, C.POINTER(Shot), C.c_int, C.c_void_p]
        L.sepfwi_forward_snapshots.argtypes = [C.c_void_p, C.POINTER(Shot), C.c_int, C.c_void_p, C.c_int, C.c_void_p]
        L.sepfwi_gradient.argtypes = [C.c_void_p, C.c_int, C.POINTER(Shot), C.c_int, C.POINTER(C.c_float),
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.sepfwi_cufd.argtypes = [C.c_void_p] * 9 + [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_char_p]
        L.sepfwi_ring_len.argtypes = [C.POINTER(Params)]
        L.sepfwi_ring_save.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.sepfwi_ring_restore.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.sepfwi_get_cpml.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.sepfwi_launch_count.argtypes = [C.c_void_p]
        L.sepfwi_plan_stream.argtypes = [C.POINTER(Params), C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_int)]
        L.sepfwi_plan_backward.argtypes = [C.POINTER(Params), C.c_int, C.c_int, C.POINTER(C.c_int)]
        L.sepfwi_plan_resident.argtypes = [C.POINTER(Params), C.c_int, C.c_int, C.c_size_t, C.POINTER(C.c_int)]
        L.sepfwi_resident_launches.argtypes = [C.c_void_p]
        L.sepfwi_resident_launches.restype = C.c_longlong
        L.sepfwi_last_timing.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float)]
        L.sepfwi_set_data_options.argtypes = [C.c_void_p, C.POINTER(DataOptions)]
        L.sepfwi_condition.argtypes = [C.c_void_p, C.POINTER(Shot), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_double),
                                       C.c_int, C.c_void_p]
        L.sepfwi_last_misfit.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
        L.sepfwi_bytes_per_slot.argtypes = [C.POINTER(Params)]
        L.sepfwi_bytes_per_slot.restype = C.c_longlong
        L.sepfwi_set_profile.argtypes = [C.c_void_p, C.c_int]
        L.sepfwi_get_profile.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_longlong)]
        L.sepfwi_kernel_name.argtypes = [C.c_int]
        L.sepfwi_kernel_name.restype = C.c_char_p
        _lib = L
    return _lib


class SepfwiError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libsepfwi error %d: %s" % (code, msg))
        self.code = code


def check(rc):
    if rc != 0:
        raise SepfwiError(rc, lib().sepfwi_last_error().decode("utf-8", "replace"))
