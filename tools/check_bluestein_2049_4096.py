#!/usr/bin/env python3
"""Non-smooth lengths in (2048, 4096]: the two specialised Bluestein launches (padded length 8192) against the oracle, and their
time against the route they replaced (Rader stages in the runtime-scheduled kernel, forced with B200FFT_RADER_MAX_PRIME=127)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import torch
import vkfft_b200 as vk
import vkfft_oracle as orc

def plan(n, batch, env):
    saved = {k: os.environ.get(k) for k in env}; os.environ.update(env)
    app = vk.VkFFTApplication()
    rc = vk.initializeVkFFT(app, vk.VkFFTConfiguration(FFTdim=1, size=[n], numberBatches=batch, device=0, normalize=1))
    for k, v in saved.items():
        if v is None: os.environ.pop(k, None)
        else: os.environ[k] = v
    return app if rc == 0 else None

for n in (2050, 2731, 3001, 3526, 4094):
    x = orc.random_input((3, n), np.complex64, seed=n)
    t = torch.from_numpy(x).cuda()
    app = plan(n, 3, {})
    desc = vk.planInfo(app)["forward"].count("bluestein")
    vk.VkFFTAppend(app, -1, vk.VkFFTLaunchParams(buffer=t)); torch.cuda.synchronize()
    e = orc.error_metrics(t.cpu().numpy(), orc.c2c(x, 1))["l2_rel"]
    vk.VkFFTAppend(app, 1, vk.VkFFTLaunchParams(buffer=t)); torch.cuda.synchronize()
    e2 = orc.error_metrics(t.cpu().numpy(), x)["l2_rel"]
    vk.deleteVkFFT(app)
    batch = (1 << 26) // n
    buf = torch.zeros(batch * n, dtype=torch.complex64, device="cuda"); torch.view_as_real(buf).uniform_(-1, 1)
    ms = {}
    for tag, env in (("bluestein", {}), ("rader", {"B200FFT_RADER_MAX_PRIME": "127"})):
        a = plan(n, batch, env)
        if a is None: ms[tag] = None; continue
        lp = vk.VkFFTLaunchParams(buffer=buf)
        vk.VkFFTAppend(a, -1, lp); vk.VkFFTAppend(a, 1, lp)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3): vk.VkFFTAppend(a, -1, lp); vk.VkFFTAppend(a, 1, lp)
        e1.record(); torch.cuda.synchronize()
        ms[tag] = e0.elapsed_time(e1) / 3
        vk.deleteVkFFT(a)
    del buf
    print(f"N={n}: bluestein launches {desc}, l2_rel fwd {e:.2e} round trip {e2:.2e} ({'ok' if e < 1e-6 and e2 < 2e-6 else 'FAIL'}); ms per pair of 512 MiB: two Bluestein launches {ms['bluestein']:.3f}, Rader stages {ms['rader'] if ms['rader'] is None else round(ms['rader'], 3)}", flush=True)
