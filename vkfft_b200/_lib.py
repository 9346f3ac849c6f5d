"""Loader + ctypes prototypes for the C ABI in include/b200fft.h (libb200fft.so).

There is deliberately no fallback: if the CUDA library is missing the import fails loudly.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libb200fft.so")

MAX_DIMS = 4


class b200fft_desc(ctypes.Structure):
    _fields_ = [
        ("struct_size", ctypes.c_uint32), ("fft_dim", ctypes.c_uint32),
        ("size", ctypes.c_uint64 * MAX_DIMS), ("number_batches", ctypes.c_uint64),
        ("coordinate_features", ctypes.c_uint64),
        ("precision", ctypes.c_uint32), ("perform_r2c", ctypes.c_uint32), ("perform_dct", ctypes.c_uint32),
        ("perform_dst", ctypes.c_uint32), ("normalize", ctypes.c_uint32),
        ("disable_reorder_four_step", ctypes.c_uint32),
        ("make_forward_plan_only", ctypes.c_uint32), ("make_inverse_plan_only", ctypes.c_uint32),
        ("is_input_formatted", ctypes.c_uint32), ("is_output_formatted", ctypes.c_uint32),
        ("inverse_return_to_input", ctypes.c_uint32), ("user_temp_buffer", ctypes.c_uint32),
        ("buffer_stride", ctypes.c_uint64 * MAX_DIMS), ("input_stride", ctypes.c_uint64 * MAX_DIMS),
        ("output_stride", ctypes.c_uint64 * MAX_DIMS),
        ("omit_dimension", ctypes.c_uint32 * MAX_DIMS), ("buffer_size", ctypes.c_uint64),
        ("temp_buffer_size", ctypes.c_uint64),
        ("device", ctypes.c_int32), ("reserved0", ctypes.c_uint32), ("stream", ctypes.c_void_p),
        ("dist_world", ctypes.c_uint32), ("dist_rank", ctypes.c_uint32),
        ("perform_convolution", ctypes.c_uint32), ("kernel_convolution", ctypes.c_uint32), ("matrix_convolution", ctypes.c_uint32),
        ("symmetric_kernel", ctypes.c_uint32), ("number_kernels", ctypes.c_uint32), ("conjugate_convolution", ctypes.c_uint32),
        ("cross_power_spectrum_normalization", ctypes.c_uint32), ("reserved1", ctypes.c_uint32), ("reserved", ctypes.c_uint64 * 3),
        ("perform_zeropadding", ctypes.c_uint32 * 4), ("zeropad_left", ctypes.c_uint64 * 4), ("zeropad_right", ctypes.c_uint64 * 4),
        ("frequency_zeropadding", ctypes.c_uint32), ("reserved2", ctypes.c_uint32 * 3),
    ]


class b200fft_buffers(ctypes.Structure):
    _fields_ = [
        ("buffer", ctypes.c_void_p), ("temp_buffer", ctypes.c_void_p), ("input_buffer", ctypes.c_void_p),
        ("output_buffer", ctypes.c_void_p),
        ("buffer_offset", ctypes.c_uint64), ("temp_buffer_offset", ctypes.c_uint64),
        ("input_buffer_offset", ctypes.c_uint64), ("output_buffer_offset", ctypes.c_uint64),
        ("stream", ctypes.c_void_p),
        ("kernel", ctypes.c_void_p), ("kernel_offset", ctypes.c_uint64),
    ]


class b200fft_plan_info(ctypes.Structure):
    _fields_ = [
        ("num_passes_forward", ctypes.c_uint32), ("num_passes_inverse", ctypes.c_uint32),
        ("temp_bytes", ctypes.c_uint64), ("lut_bytes", ctypes.c_uint64), ("algorithmic_bytes", ctypes.c_uint64),
        ("flops", ctypes.c_double),
    ]


# every symbol include/b200fft.h declares
EXPORTS = [
    "b200fft_plan_create", "b200fft_exec", "b200fft_plan_destroy", "b200fft_plan_get_info",
    "b200fft_plan_describe", "b200fft_exec_host", "b200fft_host_alloc", "b200fft_host_free",
    "b200fft_error_string", "b200fft_version", "b200fft_kernel_count",
    "b200fft_window_granularity", "b200fft_window_create", "b200fft_window_export", "b200fft_window_import",
    "b200fft_window_base", "b200fft_window_local", "b200fft_window_barrier", "b200fft_window_status",
    "b200fft_window_destroy", "b200fft_plan_attach_window", "b200fft_plan_axis_uploads",
]

_lib = None


def load():
    """Return the loaded library (ctypes.CDLL) with prototypes set. Raises if the .so is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build the CUDA engine first (python -c 'import __graft_entry__ as g; g.build()' "
            "or `make -j8`). There is no CPU fallback.")
    L = ctypes.CDLL(LIB_PATH)
    vp = ctypes.c_void_p
    L.b200fft_plan_create.argtypes = [ctypes.POINTER(b200fft_desc), ctypes.POINTER(vp)]
    L.b200fft_plan_create.restype = ctypes.c_int
    L.b200fft_exec.argtypes = [vp, ctypes.c_int, ctypes.POINTER(b200fft_buffers)]
    L.b200fft_exec.restype = ctypes.c_int
    L.b200fft_plan_destroy.argtypes = [vp]
    L.b200fft_plan_destroy.restype = None
    L.b200fft_plan_get_info.argtypes = [vp, ctypes.POINTER(b200fft_plan_info)]
    L.b200fft_plan_get_info.restype = ctypes.c_int
    L.b200fft_plan_describe.argtypes = [vp, ctypes.c_int, ctypes.c_char_p, ctypes.c_size_t]
    L.b200fft_plan_describe.restype = ctypes.c_size_t
    L.b200fft_exec_host.argtypes = [vp, ctypes.c_int, vp, vp, ctypes.c_uint64, ctypes.c_uint64]
    L.b200fft_exec_host.restype = ctypes.c_int
    L.b200fft_host_alloc.argtypes = [ctypes.c_uint64]
    L.b200fft_host_alloc.restype = vp
    L.b200fft_host_free.argtypes = [vp]
    L.b200fft_host_free.restype = None
    L.b200fft_error_string.argtypes = [ctypes.c_int]
    L.b200fft_error_string.restype = ctypes.c_char_p
    L.b200fft_version.restype = ctypes.c_int
    L.b200fft_kernel_count.restype = ctypes.c_int
    L.b200fft_window_granularity.argtypes = [ctypes.c_int]
    L.b200fft_window_granularity.restype = ctypes.c_uint64
    L.b200fft_window_create.argtypes = [ctypes.c_int, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint64, ctypes.POINTER(vp)]
    L.b200fft_window_create.restype = ctypes.c_int
    L.b200fft_window_export.argtypes = [vp, ctypes.POINTER(ctypes.c_int)]
    L.b200fft_window_export.restype = ctypes.c_int
    L.b200fft_window_import.argtypes = [vp, ctypes.c_uint32, ctypes.POINTER(ctypes.c_int)]
    L.b200fft_window_import.restype = ctypes.c_int
    L.b200fft_window_base.argtypes = [vp]
    L.b200fft_window_base.restype = vp
    L.b200fft_window_local.argtypes = [vp]
    L.b200fft_window_local.restype = vp
    L.b200fft_window_barrier.argtypes = [vp, vp]
    L.b200fft_window_barrier.restype = ctypes.c_int
    L.b200fft_window_status.argtypes = [vp]
    L.b200fft_window_status.restype = ctypes.c_int
    L.b200fft_window_destroy.argtypes = [vp]
    L.b200fft_window_destroy.restype = None
    L.b200fft_plan_attach_window.argtypes = [vp, vp]
    L.b200fft_plan_attach_window.restype = ctypes.c_int
    _lib = L
    return L
