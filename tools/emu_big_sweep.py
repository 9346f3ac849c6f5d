import os, sys, time, json
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'emu'))
import numpy as np, emu
part, parts = int(sys.argv[1]), int(sys.argv[2])
rng = np.random.default_rng(5)
sizes = sorted(set(list(range(4301, 20000, 211)) + list(range(20011, 140000, 2741)) + [4391, 8191, 10007, 16411, 32771, 65537, 20011, 17 * 1024, 19 * 2048, 127 * 256, 23 * 23 * 23, 131 * 64, 3 ** 9, 5 ** 6, 7 ** 5, 11 ** 4, 13 ** 4,
               6 ** 6, 10 ** 5, 12 ** 4 * 5, 2 ** 17, 2 ** 17 + 2 ** 16, 3 * 2 ** 15, 5 * 2 ** 14, 7 * 2 ** 13, 9 * 2 ** 13, 15 * 2 ** 12, 1000 * 128, 360 * 360, 1080 * 96]))
bad = []; t0 = time.time(); done = 0
for n in sizes[part::parts]:
    prec = 0
    x = (rng.uniform(-1, 1, (1, n)) + 1j * rng.uniform(-1, 1, (1, n))).astype(np.complex64)
    inv = -1 if n % 2 else 1
    buf = x.copy()
    try:
        rc, npass = emu.exec_plan(emu.make_desc((n,), 1, 0), inv, buf)
    except Exception as e:
        bad.append((n, 'exc', repr(e))); continue
    ref = np.fft.fft(x.astype(np.complex128), axis=-1) if inv == -1 else np.fft.ifft(x.astype(np.complex128), axis=-1) * n
    err = float(np.linalg.norm(buf - ref) / np.linalg.norm(ref))
    if rc or not err < 3e-6: bad.append((n, inv, rc, err, npass))
    done += 1
print(json.dumps({"part": part, "done": done, "bad": bad, "sec": round(time.time() - t0, 1)}))
