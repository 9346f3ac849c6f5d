"""ctypes front-end of the CPU kernel-body emulation (tests only; see cuda_emu.h)."""
import ctypes, os, subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
KIND_ROWS, KIND_ROWS_TOUT, KIND_COLS = 0, 1, 2
OP_TW, OP_SCALE = 1, 2


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "_build", "libb200fft_emu.so")
        srcs = [os.path.join(_HERE, f) for f in ("emu_driver.cpp", "cuda_emu.h")]
        csrc = os.path.join(_HERE, "..", "..", "vkfft_b200", "csrc")
        srcs += [os.path.join(csrc, f) for f in os.listdir(csrc)]
        if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
            subprocess.check_call([os.path.join(_HERE, "build.sh")])
        _LIB = ctypes.CDLL(so)
        _LIB.emu_run_pass.restype = ctypes.c_int
    return _LIB


def kernels():
    L = lib()
    out = []
    buf = (ctypes.c_int * 21)()
    for i in range(L.emu_kernel_count()):
        L.emu_kernel_info(i, buf)
        v = list(buf)
        out.append(dict(kind=v[0], prec=v[1], n=v[2], inv=v[3], ops=v[4], threads=v[5], q=v[6], tpl=v[7], v=v[8],
                        smem=v[9], ns=v[10], radices=v[11:11 + v[10]], variant=v[19], pipelined=v[20]))
    return out


def run_pass(kind, prec, n, inv, ops, inp, out, G, nb=(1, 1, 1), in_es=1, out_es=1, in_gs=None, out_gs=None,
             in_bs=(0, 0, 0), out_bs=(0, 0, 0), twM=0, tw_line0=0, scale=1.0, log=False, variant=0):
    L = lib()
    nbA = (ctypes.c_uint * 3)(*nb)
    ibs = (ctypes.c_longlong * 3)(*in_bs)
    obs = (ctypes.c_longlong * 3)(*out_bs)
    rep = (ctypes.c_double * 3)()
    rc = L.emu_run_pass(kind, prec, n, inv, ops, int(variant), inp.ctypes.data_as(ctypes.c_void_p),
                        out.ctypes.data_as(ctypes.c_void_p), ctypes.c_uint(G), nbA,
                        ctypes.c_longlong(in_es), ctypes.c_longlong(out_es), ctypes.c_longlong(in_gs),
                        ctypes.c_longlong(out_gs), ibs, obs, ctypes.c_ulonglong(twM), ctypes.c_uint(tw_line0),
                        ctypes.c_double(scale), int(log), rep)
    if rc != 0:
        raise RuntimeError(f"emu_run_pass rc={rc}")
    return dict(worst=rep[0], mean=rep[1], accesses=rep[2])


class Desc(ctypes.Structure):
    """ctypes mirror of b200fft_desc (include/b200fft.h)."""
    _fields_ = [
        ("struct_size", ctypes.c_uint32), ("fft_dim", ctypes.c_uint32),
        ("size", ctypes.c_uint64 * 4), ("number_batches", ctypes.c_uint64), ("coordinate_features", ctypes.c_uint64),
        ("precision", ctypes.c_uint32), ("perform_r2c", ctypes.c_uint32), ("perform_dct", ctypes.c_uint32),
        ("perform_dst", ctypes.c_uint32), ("normalize", ctypes.c_uint32), ("disable_reorder_four_step", ctypes.c_uint32),
        ("make_forward_plan_only", ctypes.c_uint32), ("make_inverse_plan_only", ctypes.c_uint32),
        ("is_input_formatted", ctypes.c_uint32), ("is_output_formatted", ctypes.c_uint32),
        ("inverse_return_to_input", ctypes.c_uint32), ("user_temp_buffer", ctypes.c_uint32),
        ("buffer_stride", ctypes.c_uint64 * 4), ("input_stride", ctypes.c_uint64 * 4), ("output_stride", ctypes.c_uint64 * 4),
        ("omit_dimension", ctypes.c_uint32 * 4), ("buffer_size", ctypes.c_uint64), ("temp_buffer_size", ctypes.c_uint64),
        ("device", ctypes.c_int32), ("reserved0", ctypes.c_uint32), ("stream", ctypes.c_void_p),
        ("dist_world", ctypes.c_uint32), ("dist_rank", ctypes.c_uint32),
        ("perform_convolution", ctypes.c_uint32), ("kernel_convolution", ctypes.c_uint32), ("matrix_convolution", ctypes.c_uint32),
        ("symmetric_kernel", ctypes.c_uint32), ("number_kernels", ctypes.c_uint32), ("conjugate_convolution", ctypes.c_uint32),
        ("cross_power_spectrum_normalization", ctypes.c_uint32), ("reserved1", ctypes.c_uint32), ("reserved", ctypes.c_uint64 * 3),
        ("perform_zeropadding", ctypes.c_uint32 * 4), ("zeropad_left", ctypes.c_uint64 * 4), ("zeropad_right", ctypes.c_uint64 * 4),
        ("frequency_zeropadding", ctypes.c_uint32), ("reserved2", ctypes.c_uint32 * 3),
    ]


def make_desc(shape_xyz, batches=1, prec=0, **kw):
    d = Desc()
    d.struct_size = ctypes.sizeof(Desc)
    d.fft_dim = len(shape_xyz)
    for i, s in enumerate(shape_xyz):
        d.size[i] = s
    d.number_batches = batches
    d.precision = prec
    for k, v in kw.items():
        if isinstance(v, (list, tuple)):
            for i, x in enumerate(v):
                getattr(d, k)[i] = x
        else:
            setattr(d, k, v)
    return d


def exec_plan(desc, inverse, buffer, inp=None, out=None, kernel=None):
    L = lib()
    L.emu_set_kernel(kernel.ctypes.data_as(ctypes.c_void_p) if kernel is not None else None)
    L.emu_exec_plan.restype = ctypes.c_int
    npass = ctypes.c_int(0)
    vp = lambda a: a.ctypes.data_as(ctypes.c_void_p) if a is not None else None
    rc = L.emu_exec_plan(ctypes.byref(desc), int(inverse), vp(buffer), vp(inp), vp(out), ctypes.byref(npass))
    return rc, npass.value


def exec_plan_pass(desc, inverse, buffer, temp, pass_index):
    """one launch of a plan on caller-owned buffer/temp; returns (rc, npasses, sync_before flags)"""
    L = lib()
    L.emu_exec_plan_pass.restype = ctypes.c_int
    npass = ctypes.c_int(0)
    sync = (ctypes.c_int * 16)()
    vp = lambda a: a.ctypes.data_as(ctypes.c_void_p) if a is not None else None
    rc = L.emu_exec_plan_pass(ctypes.byref(desc), int(inverse), vp(buffer), vp(temp), int(pass_index), ctypes.byref(npass), sync)
    return rc, npass.value, [bool(sync[i]) for i in range(npass.value)]


def describe(desc, inverse=-1):
    L = lib()
    buf = ctypes.create_string_buffer(16384)
    rc = L.emu_describe(ctypes.byref(desc), int(inverse), buf, len(buf))
    return rc, buf.value.decode()
