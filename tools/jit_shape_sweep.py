#!/usr/bin/env python3
"""Shape sweep for the plan-time instantiated contiguous-line kernels (jit.cpp): for a few lengths time the rule's shape against
alternatives (threads per line, lines per CTA, register budget, radix order) through B200FFT_JIT_SHAPE, ms per forward+inverse
pair of ~512 MiB.  Used to check / refine the shape rule; prints the best candidates per length."""
import itertools, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import vkfft_b200 as vk


def factor(n):
    best = None
    cands = [16, 15, 14, 13, 12, 11, 10, 9, 8, 7, 6, 5, 4, 3, 2]
    def rec(v, start, acc):
        nonlocal best
        if v == 1:
            if best is None or len(acc) < len(best) or (len(acc) == len(best) and max(acc) < max(best)):
                best = list(acc)
            return
        if best is not None and len(acc) >= len(best):
            return
        for i in range(start, len(cands)):
            if v % cands[i] == 0:
                acc.append(cands[i]); rec(v // cands[i], i, acc); acc.pop()
    rec(n, 0, [])
    return best


def timed(n, batch, buf, shape):
    if shape is None:
        os.environ.pop("B200FFT_JIT_SHAPE", None)
    else:
        os.environ["B200FFT_JIT_SHAPE"] = shape
    app = vk.VkFFTApplication()
    rc = vk.initializeVkFFT(app, vk.VkFFTConfiguration(FFTdim=1, size=[n], numberBatches=batch, device=0, normalize=1))
    os.environ.pop("B200FFT_JIT_SHAPE", None)
    if rc != 0:
        return None
    if "JIT_" not in vk.planInfo(app)["forward"]:
        vk.deleteVkFFT(app)
        return None
    lp = vk.VkFFTLaunchParams(buffer=buf)
    vk.VkFFTAppend(app, -1, lp); vk.VkFFTAppend(app, 1, lp)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3):
        vk.VkFFTAppend(app, -1, lp); vk.VkFFTAppend(app, 1, lp)
    b.record(); torch.cuda.synchronize()
    vk.deleteVkFFT(app)
    return a.elapsed_time(b) / 3


lengths = [int(x) for x in sys.argv[1:]] or [1100, 2002, 4004]
for n in lengths:
    r = factor(n)
    batch = max(1, (1 << 26) // n)
    buf = torch.zeros(batch * n, dtype=torch.complex64, device="cuda")
    torch.view_as_real(buf).uniform_(-1, 1)
    base = timed(n, batch, buf, None)
    rmax = max(r)
    tpl0 = n // rmax
    while tpl0 > 256:
        tpl0 = (tpl0 + 1) // 2
    res = []
    orders = {tuple(r), tuple(sorted(r))}
    seen = set()
    for order in orders:
        for tpl in sorted({tpl0, (tpl0 + 1) // 2}):
            for q in (1, 2):
                if tpl * q > 512 or q * n * 8 > 64 * 1024:
                    continue
                for regs in (64, 96, 128):
                    key = (order, tpl, q, regs)
                    if key in seen:
                        continue
                    seen.add(key)
                    ms = timed(n, batch, buf, f"{tpl},{q},{regs}," + ",".join(map(str, order)))
                    if ms is not None:
                        res.append((ms, key))
    res.sort()
    print(f"N={n} rule {r} tpl {tpl0}: {base:.3f} ms;  best: " + "; ".join(f"{ms:.3f} {k}" for ms, k in res[:5]) + f";  worst {res[-1][0]:.3f}" if res else f"N={n}: nothing", flush=True)
    del buf
