import sys, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/oracle'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, torch
import vkfft_b200 as vk, vkfft_oracle as orc
def run(buf, cfg, inv):
    t = torch.from_numpy(np.ascontiguousarray(buf)).cuda()
    app = vk.VkFFTApplication(); rc = vk.initializeVkFFT(app, cfg); assert rc == 0, rc
    print(vk.planInfo(app)['forward' if inv==-1 else 'inverse'])
    rc = vk.VkFFTAppend(app, inv, vk.VkFFTLaunchParams(buffer=t)); torch.cuda.synchronize(); assert rc == 0, rc
    out = t.cpu().numpy(); vk.deleteVkFFT(app); return out
for shape, batch in [((1<<17,),3), ((1<<20,),1), ((1<<20,),3)]:
    nx, H = shape[0], shape[0]//2+1
    x = orc.random_input((batch,)+tuple(reversed(shape)), np.float32, seed=sum(shape))
    buf = np.zeros(x.shape[:-1]+(2*H,), np.float32); buf[..., :nx] = x
    cfg = vk.VkFFTConfiguration(FFTdim=1, size=list(shape), numberBatches=batch, device=0, performR2C=1)
    y = run(buf, cfg, -1)
    ref = orc.r2c(x, 1)
    got = y.view(np.complex64)
    print(shape, batch, 'fwd err', orc.error_metrics(got, ref)['l2_rel'], 'per-batch', [float(orc.error_metrics(got[b], ref[b])['l2_rel']) for b in range(batch)])
    z = run(y, cfg, 1)
    print('inv err', orc.error_metrics(z[..., :nx], x.astype(np.float64)*nx)['l2_rel'])
