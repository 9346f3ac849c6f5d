import os, sys, time, json
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'emu'))
import numpy as np, emu
seed, count = int(sys.argv[1]), int(sys.argv[2])
rng = np.random.default_rng(seed)
bad = []; t0 = time.time(); done = 0
def rel(a, b): return float(np.linalg.norm(a - b) / np.linalg.norm(b))
for it in range(count):
    if it % 2 == 0:
        # distributed plans: every rank played on two host arrays
        world = int(rng.choice([2, 3, 4, 5, 6, 8]))
        f = [int(rng.choice([2, 3, 4, 5, 6, 8, 9, 10, 12, 15, 16, 24, 32, 64])) for _ in range(int(rng.integers(2, 4)))]
        n = int(np.prod(f)) * world * (world if rng.integers(0, 2) else 1)
        if n > 200000 or n < 64: continue
        inv = int(rng.choice([-1, 1])); prec = int(rng.integers(0, 2))
        cdt = np.complex64 if prec == 0 else np.complex128
        x = (rng.uniform(-1, 1, n) + 1j * rng.uniform(-1, 1, n)).astype(cdt)
        buf, tmp = x.copy(), np.zeros(n, cdt)
        descs = [emu.make_desc((n,), 1, prec, user_temp_buffer=1, dist_world=world, dist_rank=r) for r in range(world)]
        rc, npass, sync = emu.exec_plan_pass(descs[0], inv, buf, tmp, -1)
        if rc == 3002: continue
        if rc: bad.append(("dist", n, world, rc)); continue
        segs, cur = [], []
        for p in range(npass):
            if sync[p] and cur: segs.append(cur); cur = []
            cur.append(p)
        segs.append(cur)
        ok = True
        for seg in segs:
            order = list(rng.permutation(world))
            for r in order:
                for p in seg:
                    rc, _, _ = emu.exec_plan_pass(descs[int(r)], inv, buf, tmp, p)
                    ok = ok and rc == 0
        ref = np.fft.fft(x.astype(np.complex128)) if inv == -1 else np.fft.ifft(x.astype(np.complex128)) * n
        e = rel(buf, ref)
        if not ok or not e < (3e-6 if prec == 0 else 1e-12): bad.append(("dist", n, world, inv, prec, ok, e))
        done += 1
    else:
        nd = int(rng.integers(1, 4))
        pool = [8, 16, 32, 64, 128, 256, 12, 20, 48, 100, 7, 512]
        shape = tuple(int(rng.choice(pool)) for _ in range(nd))
        if np.prod(shape) > 70000: continue
        C = int(rng.integers(1, 4)); B = int(rng.integers(1, 3)); prec = int(rng.integers(0, 2)); r2c = int(rng.integers(0, 2)) if shape[0] % 2 == 0 else 0
        M = int(rng.choice([0, 0, 2, 3])); sym = int(rng.integers(0, 2)); NK = int(rng.choice([1, 1, 2]))
        if M: C = M
        if NK > 1: B = 1
        conj = int(rng.choice([0, 0, 1, 2])); norm = int(rng.integers(0, 2))
        rdt, cdt = (np.float32, np.complex64) if prec == 0 else (np.float64, np.complex128)
        axes = tuple(range(-nd, 0)); npshape = tuple(reversed(shape))
        kplanes = (M * (M + 1) // 2 if sym else M * M) if M else C
        if r2c:
            x = rng.uniform(-1, 1, (B, C) + npshape).astype(rdt)
            kk = rng.uniform(-1, 1, (NK, kplanes) + npshape).astype(rdt)
            X = np.fft.rfftn(x.astype(np.float64), axes=axes); Kf = np.fft.rfftn(kk.astype(np.float64), axes=axes)
            buf = np.zeros((max(B, NK), C) + npshape[:-1] + (shape[0] + 2,), rdt); buf[:B, ..., :shape[0]] = x
        else:
            x = (rng.uniform(-1, 1, (B, C) + npshape) + 1j * rng.uniform(-1, 1, (B, C) + npshape)).astype(cdt)
            kk = (rng.uniform(-1, 1, (NK, kplanes) + npshape) + 1j * rng.uniform(-1, 1, (NK, kplanes) + npshape)).astype(cdt)
            X = np.fft.fftn(x.astype(np.complex128), axes=axes); Kf = np.fft.fftn(kk.astype(np.complex128), axes=axes)
            buf = np.zeros((max(B, NK), C) + npshape, cdt); buf[:B] = x
        K = Kf.astype(cdt)
        Ku = np.conj(K.astype(np.complex128)) if conj == 2 else K.astype(np.complex128)
        Xu = np.conj(X) if conj == 1 else X
        outs = np.zeros((max(B, NK), C) + X.shape[2:], np.complex128)
        for o in range(max(B, NK)):
            kb = o if NK > 1 else 0; xb = 0 if NK > 1 else o
            if M:
                for r in range(M):
                    for c in range(M):
                        if sym:
                            a, b2 = min(r, c), max(r, c); idx = a * M - a * (a - 1) // 2 + (b2 - a)
                        else:
                            idx = r * M + c
                        outs[o, r] += Ku[kb, idx] * Xu[xb, c]
            else:
                for c in range(C): outs[o, c] = Ku[kb, c] * Xu[xb, c]
        ref = (np.fft.irfftn(outs, s=npshape, axes=axes) if r2c else np.fft.ifftn(outs, axes=axes)) * (1 if norm else np.prod(shape))
        d = emu.make_desc(shape, B, prec, coordinate_features=C, perform_convolution=1, perform_r2c=r2c, matrix_convolution=M, symmetric_kernel=sym,
                          number_kernels=NK, conjugate_convolution=conj, normalize=norm)
        rc, npass = emu.exec_plan(d, -1, buf, kernel=K)
        got = buf[..., :shape[0]] if r2c else buf
        e = rel(got, ref) if not rc else None
        if rc or not e < (4e-6 if prec == 0 else 1e-11): bad.append(("conv", shape, B, C, M, sym, NK, conj, norm, prec, r2c, rc, e, npass))
        done += 1
print(json.dumps({"seed": seed, "done": done, "bad": [str(b) for b in bad[:20]], "nbad": len(bad), "sec": round(time.time() - t0, 1)}))
