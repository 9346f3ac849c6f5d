"""random layouts: padded bufferStride, omitDimension, coordinateFeatures, out-of-place formatted buffers (C2C), R2C with isInputFormatted"""
import os, sys, time, json
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'emu'))
import numpy as np, emu
seed, count = int(sys.argv[1]), int(sys.argv[2])
rng = np.random.default_rng(seed)
SIZES = [2, 3, 4, 5, 6, 7, 8, 9, 10, 12, 13, 15, 16, 17, 18, 20, 24, 25, 27, 30, 32, 34, 36, 40, 45, 48, 49, 50, 60, 64, 72, 81, 96, 100, 121, 125, 128, 19, 23, 29, 37, 53]
bad = []; t0 = time.time(); done = 0
def rel(a, b):
    nb = np.linalg.norm(b); return float(np.linalg.norm(a - b) / (nb if nb else 1))
for it in range(count):
    nd = int(rng.integers(1, 4))
    shape = [int(rng.choice(SIZES)) for _ in range(nd)]
    b = int(rng.integers(1, 3)); C = int(rng.integers(1, 3)); prec = int(rng.integers(0, 2)); inv = int(rng.choice([-1, 1]))
    cdt = np.complex64 if prec == 0 else np.complex128; rdt = np.float32 if prec == 0 else np.float64
    tol = 3e-6 if prec == 0 else 1e-12
    case = str(rng.choice(["pad", "omit", "oop", "r2c_in"]))
    tag = (case, tuple(shape), b, C, prec, inv)
    try:
        if case == "pad":
            # padded strides: stride[0] >= size[0] etc
            pads = [int(rng.integers(0, 4)) for _ in range(nd)]
            st = []; cur = 1
            for a in range(nd):
                cur = cur * shape[a] + pads[a]; st.append(cur)
            tot = st[-1] * b * C
            full = (rng.uniform(-1, 1, tot) + 1j * rng.uniform(-1, 1, tot)).astype(cdt)
            buf = full.copy()
            rc, _ = emu.exec_plan(emu.make_desc(tuple(shape), b, prec, coordinate_features=C, buffer_stride=st), inv, buf)
            # reference: gather logical elements
            idx = np.zeros([b * C] + list(reversed(shape)), dtype=np.int64)
            grids = np.meshgrid(*[np.arange(s) for s in reversed(shape)], indexing='ij')
            off = np.zeros_like(grids[0])
            for a in range(nd):
                off = off + grids[nd - 1 - a] * (1 if a == 0 else st[a - 1])
            for bb in range(b * C): idx[bb] = off + bb * st[-1]
            x = full[idx]; axes = tuple(range(1, nd + 1))
            ref = np.fft.fftn(x.astype(np.complex128), axes=axes) if inv == -1 else np.fft.ifftn(x.astype(np.complex128), axes=axes) * np.prod(shape)
            e = rel(buf[idx], ref)
            mask = np.ones(tot, bool); mask[idx.ravel()] = False
            untouched = np.array_equal(buf[mask], full[mask])
            if rc or not e < tol or not untouched: bad.append((tag, st, rc, e, untouched))
        elif case == "omit":
            if nd < 2: continue
            omit = [int(rng.integers(0, 2)) for _ in range(nd)]
            if all(omit): omit[0] = 0
            npshape = (b * C,) + tuple(reversed(shape))
            x = (rng.uniform(-1, 1, npshape) + 1j * rng.uniform(-1, 1, npshape)).astype(cdt)
            buf = x.copy()
            rc, _ = emu.exec_plan(emu.make_desc(tuple(shape), b, prec, coordinate_features=C, omit_dimension=omit), inv, buf)
            axes = tuple(nd - a for a in range(nd) if not omit[a])
            n = np.prod([shape[a] for a in range(nd) if not omit[a]])
            ref = np.fft.fftn(x.astype(np.complex128), axes=axes) if inv == -1 else np.fft.ifftn(x.astype(np.complex128), axes=axes) * n
            if rc or not rel(buf, ref) < tol: bad.append((tag, omit, rc, rel(buf, ref)))
        elif case == "oop":
            npshape = (b * C,) + tuple(reversed(shape))
            x = (rng.uniform(-1, 1, npshape) + 1j * rng.uniform(-1, 1, npshape)).astype(cdt)
            fin = int(rng.integers(0, 2)); fout = 1 - fin if rng.integers(0, 2) else 1
            tin, tbuf, tout = x.copy(), np.zeros_like(x), np.zeros_like(x)
            if not fin and inv == -1: tbuf = x.copy()
            if not fout and inv == 1: tbuf = x.copy()
            if inv == 1 and fout: tout = x.copy()
            rc, _ = emu.exec_plan(emu.make_desc(tuple(shape), b, prec, coordinate_features=C, is_input_formatted=fin, is_output_formatted=fout), inv, tbuf, inp=tin, out=tout)
            axes = tuple(range(1, nd + 1))
            ref = np.fft.fftn(x.astype(np.complex128), axes=axes) if inv == -1 else np.fft.ifftn(x.astype(np.complex128), axes=axes) * np.prod(shape)
            res = (tout if fout else tbuf) if inv == -1 else tbuf
            if rc or not rel(res, ref) < tol: bad.append((tag, fin, fout, rc, rel(res, ref)))
        else:
            nx = shape[0]; H = nx // 2 + 1
            npshape = (b * C,) + tuple(reversed(shape))
            x = rng.uniform(-1, 1, npshape).astype(rdt)
            out = np.zeros(npshape[:-1] + (H,), cdt)
            d = emu.make_desc(tuple(shape), b, prec, coordinate_features=C, perform_r2c=1, is_input_formatted=1, inverse_return_to_input=1)
            rc, _ = emu.exec_plan(d, -1, out, inp=x.copy())
            if rc == 3003: continue
            axes = tuple(range(1, nd + 1))
            e1 = rel(out, np.fft.rfftn(x.astype(np.float64), axes=axes)) if not rc else None
            back = np.zeros_like(x)
            rc2, _ = emu.exec_plan(d, 1, out, inp=back)
            e2 = rel(back, x.astype(np.float64) * np.prod(shape)) if not rc2 else None
            if rc or rc2 or not e1 < tol or not e2 < tol: bad.append((tag, rc, rc2, e1, e2))
        done += 1
    except Exception as e:
        bad.append((tag, 'exc', repr(e)))
print(json.dumps({"seed": seed, "done": done, "bad": [str(x) for x in bad[:30]], "nbad": len(bad), "sec": round(time.time() - t0, 1)}))
