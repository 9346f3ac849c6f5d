#!/usr/bin/env python3
"""Small ragged cases of every kernel family, meant to run under compute-sanitizer (memcheck / racecheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vkfft_b200 as vk

CASES = [
    dict(FFTdim=1, size=[4096], numberBatches=3),                       # single pass
    dict(FFTdim=1, size=[512], numberBatches=13),                       # several lines per CTA, ragged
    dict(FFTdim=1, size=[16384], numberBatches=3),                      # TMA-fed persistent kernel
    dict(FFTdim=1, size=[1 << 17], numberBatches=3),                    # Four-Step 2 launches
    dict(FFTdim=1, size=[1 << 23], numberBatches=1),                    # Four-Step 3 launches
    dict(FFTdim=3, size=[40, 24, 12], numberBatches=2),                 # strided axes, non-pow2 (generic kernel)
    dict(FFTdim=2, size=[64, 8192], numberBatches=1),                   # strided Four-Step
    dict(FFTdim=1, size=[1000], numberBatches=7),                       # curated non-pow2 specialised kernel
    dict(FFTdim=1, size=[509], numberBatches=5),                        # Bluestein (specialised)
    dict(FFTdim=1, size=[4391], numberBatches=2),                       # long Bluestein
    dict(FFTdim=1, size=[1088], numberBatches=3),                       # Rader stage
    dict(FFTdim=2, size=[4096, 6], numberBatches=3, performR2C=1),      # fused R2C
    dict(FFTdim=1, size=[1 << 18], numberBatches=2, performR2C=1),      # long R2C
    dict(FFTdim=2, size=[256, 128], numberBatches=3, performDCT=2),     # fused DCT-II
    dict(FFTdim=2, size=[62, 8192], numberBatches=1, performDCT=3),     # long strided DCT-III
    dict(FFTdim=2, size=[34, 20], numberBatches=3, performDCT=4),       # generic DCT-IV
    dict(FFTdim=1, size=[64], numberBatches=5, performDST=2),
    dict(FFTdim=3, size=[64, 32, 16], numberBatches=1, doublePrecision=1),
]
for c in CASES:
    dbl = c.get("doublePrecision", 0)
    n = c["numberBatches"]
    for s in c["size"]:
        n *= s
    real = any(k in c for k in ("performDCT", "performDST"))
    if c.get("performR2C"):
        n = c["numberBatches"] * (c["size"][0] // 2 + 1) * 2
        for s in c["size"][1:]:
            n *= s
        real = True
    dt = torch.float64 if dbl else torch.float32
    buf = torch.zeros(n * (1 if real else 2), dtype=dt, device="cuda").uniform_(-1, 1)
    app = vk.VkFFTApplication()
    rc = vk.initializeVkFFT(app, vk.VkFFTConfiguration(device=0, **c))
    assert rc == 0, (c, vk.getVkFFTErrorString(rc))
    for inv in (-1, 1):
        assert vk.VkFFTAppend(app, inv, vk.VkFFTLaunchParams(buffer=buf)) == 0
    torch.cuda.synchronize()
    vk.deleteVkFFT(app)
    print("ok", c, flush=True)
