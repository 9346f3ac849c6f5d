#!/usr/bin/env python3
"""Time every registered specialised kernel (all variants) on a 2 GiB synthetic pass: python tools/ktune.py [prec=0] [filter_n]"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vkfft_b200 import _lib

prec = int(sys.argv[1]) if len(sys.argv) > 1 else 0
only_n = set(int(x) for x in sys.argv[2].split(",")) if len(sys.argv) > 2 else None
L = _lib.load()
L.b200fft_debug_time_kernel.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint32,
                                        ctypes.c_int, ctypes.POINTER(ctypes.c_float), ctypes.c_char_p, ctypes.c_int]
esz = 16 if prec else 8
total = (1 << 31) // esz
a = torch.zeros(total * esz // 4, dtype=torch.float32, device="cuda").uniform_(-1, 1)
b = torch.zeros_like(a)
info = (ctypes.c_int * 10)()
rows = []
for i in range(L.b200fft_kernel_count()):
    L.b200fft_debug_kernel_info(i, info)
    kind, p, n, inv, ops, var, thr, q, tpl, smem = list(info)
    if kind > 2 or p != prec or inv != 0: continue
    if ops & ~1: continue   # fused real-transform / Bluestein flavours need their own tables: not timed here
    if kind == 2 and ops == 0 and not os.environ.get("KTUNE_COLS_PLAIN"): continue   # default: four-step flavour of COLS
    if kind == 2 and ops != 0 and os.environ.get("KTUNE_COLS_PLAIN"): continue
    if only_n and n not in only_n: continue
    ms = ctypes.c_float(0)
    name = ctypes.create_string_buffer(256)
    rc = L.b200fft_debug_time_kernel(i, a.data_ptr(), b.data_ptr() if kind == 1 else a.data_ptr(), total, int(os.environ.get("KTUNE_OTHER", "1024")), 5,
                                     ctypes.byref(ms), name, 256)
    gbs = 2 * total * esz / (ms.value * 1e-3) / 1e9 if ms.value > 0 else 0
    rows.append((kind, n, var, ms.value, gbs, thr, q, smem, name.value.decode(), rc))
for r in sorted(rows):
    print(f"kind={r[0]} n={r[1]:6d} v={r[2]} {r[3]*1e3:8.1f} us {r[4]:7.0f} GB/s thr={r[5]:4d} q={r[6]:2d} smem={r[7]:6d} rc={r[9]} {r[8]}")
