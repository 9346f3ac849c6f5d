#!/usr/bin/env python3
"""BASELINE.json configs 3-5 (single GPU part): time the engine and the unmodified reference (oracle/_ref) on the same
GPU, report ms per forward+inverse pair and the fraction of the HBM roofline (algorithmic bytes / time / measured peak)."""
import ctypes, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import vkfft_b200 as vk
import vkfft_oracle as orc

PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0

CASES = [
    # name, size_xyz, batch, double, kwargs(engine), kwargs(ref), real?
    ("3D C2C FP64 256^3 x8", (256, 256, 256), 8, True, {}, {}, False),
    ("3D C2C FP64 512^3", (512, 512, 512), 1, True, {}, {}, False),
    ("3D C2C FP32 512^3 x2", (512, 512, 512), 2, False, {}, {}, False),
    ("2D R2C FP32 4096^2 x16", (4096, 4096), 16, False, dict(performR2C=1), dict(perform_r2c=1), True),
    ("2D DCT-II FP32 8192^2 x2", (8192, 8192), 2, False, dict(performDCT=2), dict(perform_dct=2), True),
    ("2D C2C FP32 4096^2 x8", (4096, 4096), 8, False, {}, {}, False),
    ("1D C2C FP32 2^26 x4", (1 << 26,), 4, False, {}, {}, False),
    ("1D C2C FP32 N=1000 batch 2^18", (1000,), 1 << 18, False, {}, {}, False),
    ("1D C2C FP32 N=2187 batch 2^16", (2187,), 1 << 16, False, {}, {}, False),
    ("1D C2C FP32 N=509 (Bluestein) batch 2^18", (509,), 1 << 18, False, {}, {}, False),
    ("1D C2C FP32 N=1088 (17 x 64, prime-radix kernel) batch 2^17", (1088,), 1 << 17, False, {}, {}, False),
    ("1D C2C FP32 N=2032 (16 x 127, Rader stage in the runtime-scheduled kernel) batch 2^16", (2032,), 1 << 16, False, {}, {}, False),
]


def time_pairs(fn, reps):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


out = []
for name, size, batch, dbl, kw, rkw, real in CASES:
    pts = batch
    for s in size:
        pts *= s
    esz = (8 if dbl else 4) * (1 if real else 2)
    if kw.get("performR2C"):
        alloc = batch * (size[0] // 2 + 1) * 2
        for s in size[1:]:
            alloc *= s
    else:
        alloc = pts * (1 if real else 2)
    dt = torch.float64 if dbl else torch.float32
    buf = torch.zeros(alloc, dtype=dt, device="cuda").uniform_(-1, 1)
    row = {"case": name}
    app = vk.VkFFTApplication()
    rc = vk.initializeVkFFT(app, vk.VkFFTConfiguration(FFTdim=len(size), size=list(size), numberBatches=batch, device=0,
                                                       doublePrecision=int(dbl), normalize=1, **kw))
    if rc != 0:
        row["engine"] = vk.getVkFFTErrorString(rc)
    else:
        info = vk.planInfo(app)
        lp = vk.VkFFTLaunchParams(buffer=buf)
        ms = time_pairs(lambda: (vk.VkFFTAppend(app, -1, lp), vk.VkFFTAppend(app, 1, lp)), 3)
        alg = 2 * info["algorithmic_bytes"]
        row.update(engine_ms_pair=round(ms, 3), passes=info["num_passes_forward"], frac_of_peak=round(alg / (ms * 1e-3) / 1e9 / PEAK, 3))
        vk.deleteVkFFT(app)
    if orc.ref_available():
        L = orc.ref_lib()
        d = orc.ref_desc(size, batch, dbl, **rkw)
        h = ctypes.c_void_p()
        rc = L.vkref_open(ctypes.byref(d), ctypes.byref(h))
        if rc == 0:
            e, w = ctypes.c_double(), ctypes.c_double()
            buf.uniform_(-1e-3, 1e-3)
            rc = L.vkref_bench_pairs(h, buf.data_ptr(), 1, 3, ctypes.byref(e), ctypes.byref(w))
            row["reference_ms_pair"] = round(e.value, 3) if rc == 0 else f"error {rc}"
            L.vkref_close(h)
        else:
            row["reference_ms_pair"] = f"init error {rc}"
    print(json.dumps(row), flush=True)
    out.append(row)
    del buf
    torch.cuda.empty_cache()
