#!/usr/bin/env python3
"""One-launch Bluestein (stockham.cuh RMODE 11) against the two-launch plan, the runtime-scheduled kernel (smooth lengths
without a curated kernel) and the reference's CUDA backend: ms per forward+inverse pair of ~512 MiB of complex64."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import vkfft_b200 as vk
import vkfft_oracle as orc

def timed(n, batch, buf, env, dbl=False):
    saved = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    app = vk.VkFFTApplication()
    rc = vk.initializeVkFFT(app, vk.VkFFTConfiguration(FFTdim=1, size=[n], numberBatches=batch, device=0, normalize=1, doublePrecision=int(dbl)))
    for k, v in saved.items():
        if v is None: os.environ.pop(k, None)
        else: os.environ[k] = v
    if rc != 0: return None, 0
    lp = vk.VkFFTLaunchParams(buffer=buf)
    np_ = vk.planInfo(app)["num_passes_forward"]
    for _ in range(2): vk.VkFFTAppend(app, -1, lp); vk.VkFFTAppend(app, 1, lp)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5): vk.VkFFTAppend(app, -1, lp); vk.VkFFTAppend(app, 1, lp)
    b.record(); torch.cuda.synchronize()
    vk.deleteVkFFT(app)
    return a.elapsed_time(b) / 5, np_

def reference(n, batch, buf, dbl=False):
    if not orc.ref_available(): return None
    L = orc.ref_lib(); d = orc.ref_desc((n,), batch, dbl); h = ctypes.c_void_p()
    ref = None
    if L.vkref_open(ctypes.byref(d), ctypes.byref(h)) == 0:
        e, w = ctypes.c_double(), ctypes.c_double()
        if L.vkref_bench_pairs(h, buf.data_ptr(), 2, 5, ctypes.byref(e), ctypes.byref(w)) == 0: ref = e.value
        L.vkref_close(h)
    return ref

f = lambda v: "    -   " if v is None else f"{v:8.3f}"
print("non-smooth lengths: one launch | two launches | reference")
for n in (17 * 2, 37, 51, 94, 113, 127, 251, 323, 509, 529, 761, 1019, 1517, 2032, 2039, 3001, 4093):
    batch = max(1, (1 << 26) // n)
    buf = torch.zeros(batch * n, dtype=torch.complex64, device="cuda")
    torch.view_as_real(buf).uniform_(-1, 1)
    one, p1 = timed(n, batch, buf, {})
    two, p2 = timed(n, batch, buf, {"B200FFT_NO_FUSED_BLUESTEIN": "1"})
    ref = reference(n, batch, buf)
    print(f"N={n:6d} batch {batch:8d}: {f(one)} ms ({p1})   {f(two)} ms ({p2})   reference {f(ref)}", flush=True)
    del buf
print("smooth lengths without an ahead-of-time kernel: plan-time instantiated template | runtime-scheduled kernel (first/last stage from registers) | the same with separate copy phases (round 1) | reference")
for n in (34, 66, 154, 182, 286, 338, 770, 1001, 1100, 1430, 1694, 2002, 2310, 2730, 3003, 3146, 4004):
    batch = max(1, (1 << 26) // n)
    buf = torch.zeros(batch * n, dtype=torch.complex64, device="cuda")
    torch.view_as_real(buf).uniform_(-1, 1)
    jit, p1 = timed(n, batch, buf, {})
    gen, p2 = timed(n, batch, buf, {"B200FFT_NO_JIT": "1", "B200FFT_RADER_MAX_PRIME": "127"} if n == 34 else {"B200FFT_NO_JIT": "1"})
    old, p3 = timed(n, batch, buf, {"B200FFT_NO_JIT": "1", "B200FFT_GENERIC_STAGED": "1", **({"B200FFT_RADER_MAX_PRIME": "127"} if n == 34 else {})})
    ref = reference(n, batch, buf)
    print(f"N={n:6d} batch {batch:8d}: {f(jit)} ms ({p1})   {f(gen)} ms ({p2})   {f(old)} ms ({p3})   reference {f(ref)}", flush=True)
    del buf
print("2-D / R2C with such lengths: plan-time templates | runtime-scheduled kernel | reference")
for shape, r2c in (((1100, 1100), False), ((770, 1430), False), ((2002, 154), True)):
    nx, ny = shape
    batch = max(1, (1 << 25) // (nx * ny))
    for env, tag in (({}, "templates"), ({"B200FFT_NO_JIT": "1"}, "runtime-scheduled")):
        saved = {k: os.environ.get(k) for k in env}
        os.environ.update(env)
        app = vk.VkFFTApplication()
        if r2c:
            buf = torch.zeros(batch * ny * (nx + 2), dtype=torch.float32, device="cuda").uniform_(-1, 1)
        else:
            buf = torch.zeros(batch * ny * nx, dtype=torch.complex64, device="cuda"); torch.view_as_real(buf).uniform_(-1, 1)
        rc = vk.initializeVkFFT(app, vk.VkFFTConfiguration(FFTdim=2, size=[nx, ny], numberBatches=batch, device=0, normalize=1, performR2C=int(r2c)))
        for k, v in saved.items():
            if v is None: os.environ.pop(k, None)
            else: os.environ[k] = v
        if rc != 0:
            print(shape, tag, "rc", rc); continue
        lp = vk.VkFFTLaunchParams(buffer=buf)
        for _ in range(2): vk.VkFFTAppend(app, -1, lp); vk.VkFFTAppend(app, 1, lp)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5): vk.VkFFTAppend(app, -1, lp); vk.VkFFTAppend(app, 1, lp)
        b.record(); torch.cuda.synchronize()
        print(f"{'R2C' if r2c else 'C2C'} {nx} x {ny} batch {batch}: {tag} {a.elapsed_time(b) / 5:8.3f} ms ({vk.planInfo(app)['num_passes_forward']} launches)", flush=True)
        vk.deleteVkFFT(app)
        del buf
print("FP64, non-smooth: one launch | two launches | reference")
for n in (127, 509, 1019, 2039):
    batch = max(1, (1 << 25) // n)
    buf = torch.zeros(batch * n, dtype=torch.complex128, device="cuda")
    torch.view_as_real(buf).uniform_(-1, 1)
    one, p1 = timed(n, batch, buf, {}, True)
    two, p2 = timed(n, batch, buf, {"B200FFT_NO_FUSED_BLUESTEIN": "1"}, True)
    ref = reference(n, batch, buf, True)
    print(f"N={n:6d} batch {batch:8d}: {f(one)} ms ({p1})   {f(two)} ms ({p2})   reference {f(ref)}", flush=True)
    del buf
