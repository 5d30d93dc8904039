"""ctypes binding of libradnet_b200.so (C ABI declared in include/radnet_b200.h).

There is no CPU fallback: if the shared library is missing the import of any
compute entry point fails loudly, and every compute call without a CUDA device
returns RADNET_E_CUDA, which is raised as `RadnetError`.
"""
import ctypes
import os

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RADNET_B200_LIB", os.path.join(_PKG_DIR, "_C", "libradnet_b200.so"))

c_int = ctypes.c_int
c_size_t = ctypes.c_size_t
c_void_p = ctypes.c_void_p
c_double = ctypes.c_double
c_float = ctypes.c_float
c_longlong = ctypes.c_longlong

ABI_VERSION = 2      # RADNET_ABI_VERSION of include/radnet_b200.h

# name -> (restype, argtypes); mirrors include/radnet_b200.h one to one
SIGNATURES = {
    "radnet_det_record_bytes": (c_size_t, [c_int]),
    "radnet_version": (c_int, []),
    "radnet_last_error_string": (ctypes.c_char_p, []),
    "radnet_error_name": (ctypes.c_char_p, [c_int]),
    "radnet_device_info": (c_int, [c_void_p]),
    "radnet_decode_clip_i32": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_float,
                                       c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "radnet_decode_clip_f64": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_float,
                                       c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "radnet_apply_regr": (c_int, [c_void_p, c_void_p, c_longlong, c_void_p, c_void_p]),
    "radnet_sort_nms_i32_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_int]),
    "radnet_sort_nms_i32": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_double, c_int,
                                    c_void_p, c_void_p, c_size_t, c_void_p]),
    "radnet_nms_f64_workspace_bytes": (c_size_t, [c_int, c_int]),
    "radnet_nms_f64": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_double, c_int, c_void_p, c_void_p,
                               c_void_p, c_size_t, c_void_p]),
    "radnet_roi_pool": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p,
                                c_int, c_int, c_void_p, c_void_p]),
    "radnet_roi_pool_form": (c_int, [c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "radnet_set_option": (c_int, [ctypes.c_char_p, c_longlong]),
    "radnet_get_option": (c_int, [ctypes.c_char_p, c_void_p]),
    "radnet_rpn_targets_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_int]),
    "radnet_rpn_targets_workspace_init": (c_int, [c_void_p, c_size_t, c_int, c_int, c_void_p]),
    "radnet_rpn_targets": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                                   c_void_p, c_double, c_void_p, c_double, c_int, c_double, c_void_p,
                                   c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "radnet_roi_targets": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_double,
                                   c_double, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                   c_void_p, c_void_p]),
    "radnet_roi_targets_batch": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_void_p,
                                         c_void_p, c_void_p, c_int, c_int, c_int, c_double, c_double,
                                         c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                         c_void_p, c_void_p]),
    "radnet_mt19937_seed": (c_int, [c_void_p, c_int, c_void_p, c_void_p]),
    "radnet_rpn_subsample_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "radnet_rpn_subsample": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                                     c_void_p, c_size_t, c_void_p]),
    "radnet_select_samples": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                                      c_void_p, c_void_p]),
    "radnet_rpn_losses_workspace_bytes": (c_size_t, [c_int]),
    "radnet_rpn_losses_workspace_init": (c_int, [c_void_p, c_size_t, c_int, c_void_p]),
    "radnet_rpn_losses": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p,
                                  c_void_p, c_size_t, c_void_p]),
    "radnet_class_losses": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p,
                                    c_void_p, c_void_p, c_void_p]),
    "radnet_match_detections_workspace_bytes": (c_size_t, [c_int, c_int]),
    "radnet_match_detections": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_double,
                                        c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "radnet_class_ap_workspace_bytes": (c_size_t, [c_int]),
    "radnet_class_ap": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                c_void_p, c_size_t, c_void_p]),
    "radnet_synth_panels": (c_int, [ctypes.c_ulonglong, c_longlong, c_longlong, c_int, c_int, c_int, c_int, c_int,
                                    c_void_p, c_void_p, c_void_p, c_void_p]),
    "radnet_iou_pairs": (c_int, [c_void_p, c_void_p, c_longlong, c_void_p, c_void_p]),
    "radnet_real_coordinates": (c_int, [c_void_p, c_longlong, c_double, c_void_p, c_void_p]),
    "radnet_cls_record_bytes": (c_size_t, [c_int]),
    "radnet_classify_decode": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_void_p,
                                       c_void_p, c_double, c_void_p, c_int, c_void_p, c_int, c_void_p]),
    "radnet_classify_nms": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_void_p,
                                    c_void_p, c_double, c_void_p, c_int, c_double, c_int, c_void_p, c_void_p,
                                    c_void_p, c_int, c_void_p]),
    "radnet_class_nms_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "radnet_class_nms": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_double, c_int, c_void_p,
                                 c_void_p, c_void_p, c_int, c_void_p, c_size_t, c_void_p]),
    "radnet_final_nms_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "radnet_final_nms": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_double, c_double, c_int,
                                 c_void_p, c_int, c_void_p, c_size_t, c_void_p]),
}


class RadnetError(RuntimeError):
    """A libradnet_b200 entry point returned a negative status."""

    def __init__(self, code, name, message):
        super().__init__("%s (%d): %s" % (name, code, message))
        self.code = code


_lib = None


def load():
    """Load the shared library (once).  Raises ImportError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise ImportError(
            "libradnet_b200.so is missing at %s - build it with "
            "`python -m rock_art_radnet_b200.build` (nvcc, sm_100a). There is no CPU fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)       # AttributeError here = ABI mismatch, also loud
        fn.restype = res
        fn.argtypes = args
    if lib.radnet_version() != ABI_VERSION:
        raise ImportError("libradnet_b200.so ABI version %d != %d - rebuild it with "
                          "`python -m rock_art_radnet_b200.build`" % (lib.radnet_version(), ABI_VERSION))
    _lib = lib
    return lib


def call(name, *args):
    """Invoke a status-returning entry point; raise RadnetError on failure."""
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        msg = lib.radnet_last_error_string().decode("utf-8", "replace")
        raise RadnetError(rc, lib.radnet_error_name(rc).decode(), msg)
    return rc


def det_record_bytes(max_boxes):
    return int(load().radnet_det_record_bytes(int(max_boxes)))


def set_option(name, value):
    """Process-wide tuning option of the library (include/radnet_b200.h, radnet_set_option)."""
    call("radnet_set_option", name.encode(), int(value))


def get_option(name):
    out = ctypes.c_longlong(0)
    call("radnet_get_option", name.encode(), ctypes.addressof(out))
    return int(out.value)


class option:
    """`with option("nms_cluster", 0): ...` - set an option for a block, restore it afterwards."""

    def __init__(self, name, value):
        self.name, self.value = name, value

    def __enter__(self):
        self.old = get_option(self.name)
        set_option(self.name, self.value)
        return self

    def __exit__(self, *exc):
        set_option(self.name, self.old)
        return False
