#!/usr/bin/env python3
"""Fused Four-Step (csrc/fused4.cuh) on the GPU: bit-exactness against the two-launch plan, error against torch.fft,
and time per 2 GiB forward transform for a grid of ring settings.  python tools/fused_sweep.py [quick]"""
import os, sys, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vkfft_b200 as vk
os.environ["B200FFT_FUSED4"] = "1"

PEAK = 6575.4e9
pts = 1 << 28
quick = len(sys.argv) > 1 and sys.argv[1] == "quick"


def plan(n, batch, **env):
    old = {k: os.environ.get(k) for k in env}
    for k, v in env.items():
        if v is None: os.environ.pop(k, None)
        else: os.environ[k] = str(v)
    app = vk.VkFFTApplication()
    rc = vk.initializeVkFFT(app, vk.VkFFTConfiguration(FFTdim=1, size=[n], numberBatches=batch, device=0))
    for k, v in old.items():
        if v is None: os.environ.pop(k, None)
        else: os.environ[k] = v
    assert rc == 0, rc
    return app


def timed(fn, reps=6):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


buf = torch.empty(pts, dtype=torch.complex64, device="cuda")
for logn in range(15, 23):
    n = 1 << logn
    # ---- correctness: ring smaller than the number of units, both directions
    batch = max(24, (96 << 20) // (n * 8))
    x = torch.view_as_complex(torch.empty(batch * n, 2, device="cuda").uniform_(-1, 1)).reshape(batch, n)
    a_f = plan(n, batch)
    a_u = plan(n, batch, B200FFT_NO_FUSED4=1)
    desc = vk.planInfo(a_f)["forward"].split("\n")[0]
    ok = True
    for inv in (-1, 1):
        y1 = x.clone(); y2 = x.clone()
        vk.VkFFTAppend(a_f, inv, vk.VkFFTLaunchParams(buffer=y1))
        vk.VkFFTAppend(a_u, inv, vk.VkFFTLaunchParams(buffer=y2))
        torch.cuda.synchronize()
        ref = torch.fft.fft(x.to(torch.complex128), dim=1) if inv == -1 else torch.fft.ifft(x.to(torch.complex128), dim=1) * n
        e1 = (torch.linalg.norm(y1.to(torch.complex128) - ref) / torch.linalg.norm(ref)).item()
        same = torch.equal(torch.view_as_real(y1), torch.view_as_real(y2))
        ok &= same and e1 < 1e-6
        print(f"N=2^{logn} batch {batch} inv {inv:2d}: fused vs two-launch bit-identical={same}  rel err vs fft64 {e1:.2e}", flush=True)
    vk.deleteVkFFT(a_f); vk.deleteVkFFT(a_u)
    del x, y1, y2, ref
    print("   ", desc[:230])
    # ---- timing on the 2 GiB buffer
    lp = vk.VkFFTLaunchParams(buffer=buf)
    torch.view_as_real(buf).uniform_(-1, 1)
    a_u = plan(n, pts // n, B200FFT_NO_FUSED4=1)
    t_u = timed(lambda: vk.VkFFTAppend(a_u, -1, lp))
    vk.deleteVkFFT(a_u)
    line = f"N=2^{logn}: two launches {t_u*1e3:7.1f} us ({4*pts*8/1e6/t_u/PEAK*1e9:.2f})"
    best = None
    seq_kb = n * 8 // 1024
    for group in ((0,) if quick else (0, 8, 16, 32, 64, 128)):
        torch.view_as_real(buf).uniform_(-1, 1)
        env = dict(B200FFT_FUSED_GROUP=group) if group else {}
        a_f = plan(n, pts // n, **env)
        note = vk.planInfo(a_f)["forward"].split("\n")[0]
        note = note[note.index("groups of"):note.index("]")] if "groups of" in note else note[-80:]
        t = timed(lambda: vk.VkFFTAppend(a_f, -1, lp))
        vk.deleteVkFFT(a_f)
        print(f"      group {group:3d}: {t*1e3:7.1f} us  frac {2*pts*8/1e6/t/PEAK*1e9:.3f}   [{note}]", flush=True)
        if best is None or t < best[0]: best = (t, group, 0)
    if best is None:
        print(line + " | fused: no configuration ran"); continue
    print(line + f" | fused best {best[0]*1e3:7.1f} us frac {2*pts*8/1e6/best[0]/PEAK*1e9:.3f} (group {best[1]})  ok={ok}", flush=True)
