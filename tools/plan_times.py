#!/usr/bin/env python3
"""Per-launch times of a plan (b200fft_debug_exec_timed): python tools/plan_times.py  -> the BASELINE configs."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vkfft_b200 as vk
from vkfft_b200 import _lib

L = _lib.load()
L.b200fft_debug_exec_timed.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
CASES = [
    ("2D DCT-II FP32 8192^2 x2", dict(FFTdim=2, size=[8192, 8192], numberBatches=2, performDCT=2), 8192 * 8192 * 2, torch.float32),
    ("3D C2C FP64 256^3 x8", dict(FFTdim=3, size=[256, 256, 256], numberBatches=8, doublePrecision=1), 256 ** 3 * 8 * 2, torch.float64),
    ("2D R2C FP32 4096^2 x16", dict(FFTdim=2, size=[4096, 4096], numberBatches=16, performR2C=1), 4098 * 4096 * 16, torch.float32),
    ("1D C2C FP32 N=2187 batch 2^16", dict(FFTdim=1, size=[2187], numberBatches=1 << 16), 2187 * 65536 * 2, torch.float32),
    ("1D C2C FP32 N=1000 batch 2^18", dict(FFTdim=1, size=[1000], numberBatches=1 << 18), 1000 * (1 << 18) * 2, torch.float32),
    ("1D C2C FP32 N=8192 batch 2^15", dict(FFTdim=1, size=[8192], numberBatches=1 << 15), (1 << 28) * 2, torch.float32),
    ("1D C2C FP32 N=2048 batch 2^17", dict(FFTdim=1, size=[2048], numberBatches=1 << 17), (1 << 28) * 2, torch.float32),
]
for name, cfg, nscal, dt in CASES:
    buf = torch.zeros(nscal, dtype=dt, device="cuda").uniform_(-1, 1)
    app = vk.VkFFTApplication()
    rc = vk.initializeVkFFT(app, vk.VkFFTConfiguration(device=0, **cfg))
    assert rc == 0, (name, rc)
    info = vk.planInfo(app)
    for inv in (-1, 1):
        b = _lib.b200fft_buffers()
        b.buffer = buf.data_ptr()
        ms, kind, n = (ctypes.c_float * 32)(), (ctypes.c_int * 32)(), ctypes.c_int(0)
        best = None
        for _ in range(4):
            rc = L.b200fft_debug_exec_timed(app._plan, inv, ctypes.byref(b), ms, kind, 32, ctypes.byref(n))
            assert rc == 0
            cur = [ms[i] for i in range(n.value)]
            best = cur if best is None else [min(a, c) for a, c in zip(best, cur)]
        notes = (info["forward"] if inv == -1 else info["inverse"]).strip().split("\n")
        print(f"== {name} {'forward' if inv == -1 else 'inverse'}: total {sum(best):.3f} ms")
        for t, note in zip(best, notes):
            print(f"   {t * 1e3:8.1f} us  {note}")
    vk.deleteVkFFT(app)
    del buf
