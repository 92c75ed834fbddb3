import sys, os, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import bench
from audian_b200 import _lib, device
_lib.init(0)
sos, esos = bench.designs()
C, n = bench.CHANNELS, bench.FRAMES
NFFT, HOP, RATE = bench.NFFT, bench.HOP, bench.RATE
nspec = n//HOP
x_dev = device.synth(0, n, C, RATE, bench.SEED)
def buffers(kind):
    shapes = [(n, C), (n, C), (nspec, C, NFFT//2 + 1), (n, C)]
    if kind == 'register':
        arrs = [np.empty(s) for s in shapes]
        for a in arrs: _lib.host_register(a)
        keep = None
    else:
        keep = [torch.empty(s, dtype=torch.float64).pin_memory() for s in shapes]
        arrs = [t.numpy() for t in keep]
    arrs[0][:] = x_dev.cpu().numpy()
    return arrs, keep
for kind in ('register', 'hostalloc', 'register', 'hostalloc'):
    (hx, hf, hs, he), keep = buffers(kind)
    def step():
        _lib.chain(sos, hx, hf, RATE, 0, spec=hs, nfft=NFFT, hop=HOP, esos=esos, env=he, clamp_negative=True)
    for _ in range(2): step()
    torch.cuda.synchronize()
    ts = []
    for _ in range(6):
        t0 = time.perf_counter(); step(); ts.append(time.perf_counter() - t0)
    print(kind, 'ms per step: min %.2f  median %.2f' % (min(ts)*1e3, sorted(ts)[len(ts)//2]*1e3), flush=True)
    if kind == 'register':
        for a in (hx, hf, hs, he): _lib.host_unregister(a)
    del hx, hf, hs, he, keep
