#!/usr/bin/env python3
"""Lengths with prime factors 17..127: Rader stage in the runtime-scheduled kernel against Bluestein (B200FFT_RADER_MAX_PRIME=13),
ms per forward+inverse pair of ~512 MiB, and the reference's CUDA backend."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import vkfft_b200 as vk
import vkfft_oracle as orc

def timed(n, batch, buf, env):
    old = os.environ.get("B200FFT_RADER_MAX_PRIME")
    if env is None: os.environ.pop("B200FFT_RADER_MAX_PRIME", None)
    else: os.environ["B200FFT_RADER_MAX_PRIME"] = str(env)
    app = vk.VkFFTApplication()
    rc = vk.initializeVkFFT(app, vk.VkFFTConfiguration(FFTdim=1, size=[n], numberBatches=batch, device=0, normalize=1))
    if old is None: os.environ.pop("B200FFT_RADER_MAX_PRIME", None)
    else: os.environ["B200FFT_RADER_MAX_PRIME"] = old
    if rc != 0: return None, 0
    lp = vk.VkFFTLaunchParams(buffer=buf)
    np_ = vk.planInfo(app)["num_passes_forward"]
    for _ in range(2): vk.VkFFTAppend(app, -1, lp); vk.VkFFTAppend(app, 1, lp)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(4): vk.VkFFTAppend(app, -1, lp); vk.VkFFTAppend(app, 1, lp)
    b.record(); torch.cuda.synchronize()
    vk.deleteVkFFT(app)
    return a.elapsed_time(b) / 4, np_

for n in (34, 51, 102, 136, 323, 17 * 81, 19 * 100, 23 * 64 * 3, 529, 12167, 94, 37 * 41, 47 * 16, 61 * 27, 2032, 127 * 9, 101 * 10, 113):
    batch = max(1, (1 << 26) // n)
    buf = torch.zeros(batch * n, dtype=torch.complex64, device="cuda")
    torch.view_as_real(buf).uniform_(-1, 1)
    r, pr = timed(n, batch, buf, 127)
    bl, pb = timed(n, batch, buf, 13)
    ref = None
    if orc.ref_available():
        L = orc.ref_lib(); d = orc.ref_desc((n,), batch, False); h = ctypes.c_void_p()
        if L.vkref_open(ctypes.byref(d), ctypes.byref(h)) == 0:
            e, w = ctypes.c_double(), ctypes.c_double()
            buf.uniform_(-1e-3, 1e-3) if False else None
            if L.vkref_bench_pairs(h, buf.data_ptr(), 1, 3, ctypes.byref(e), ctypes.byref(w)) == 0: ref = e.value
            L.vkref_close(h)
    print(f"N={n:6d} batch {batch:8d}: Rader stage {r:8.3f} ms ({pr} launches)   Bluestein {bl:8.3f} ms ({pb} launches)   reference {ref if ref is None else round(ref,3)}", flush=True)
    del buf
