#!/usr/bin/env python3
"""Time the fused Four-Step for explicit settings: python tools/fused_exp.py "log2n:group:ctas:flags" ...   (0 = default;
group = CTAs per group, ctas = cap on resident CTAs, flags: 1 no discard, 2 ignore dependencies)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vkfft_b200 as vk
os.environ["B200FFT_FUSED4"] = "1"

PEAK = 6575.4e9
pts = 1 << 28
buf = torch.empty(pts, dtype=torch.complex64, device="cuda")
torch.view_as_real(buf).uniform_(-1, 1)
lp = vk.VkFFTLaunchParams(buffer=buf)
for spec in sys.argv[1:]:
    f = [int(x) for x in spec.split(":")] + [0] * 6
    logn, group, ctas, flags = f[:4]
    env = {"B200FFT_FUSED_GROUP": group, "B200FFT_FUSED_CTAS": ctas, "B200FFT_FUSED_FLAGS": flags}
    for k, v in env.items():
        if v: os.environ[k] = str(v)
        else: os.environ.pop(k, None)
    n = 1 << logn
    app = vk.VkFFTApplication()
    assert vk.initializeVkFFT(app, vk.VkFFTConfiguration(FFTdim=1, size=[n], numberBatches=pts // n, device=0)) == 0
    note = vk.planInfo(app)["forward"].split("\n")[0]
    note = note[note.index("groups of"):note.index("]")] if "groups of" in note else "NOT FUSED"
    for _ in range(2): vk.VkFFTAppend(app, -1, lp)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5): vk.VkFFTAppend(app, -1, lp)
    b.record(); torch.cuda.synchronize()
    t = a.elapsed_time(b) / 5
    print(f"{spec:28s} {t*1e3:8.1f} us  frac {2*pts*8/1e6/t/PEAK*1e9:.3f}  [{note}]", flush=True)
    vk.deleteVkFFT(app)
    torch.view_as_real(buf).uniform_(-1, 1)
