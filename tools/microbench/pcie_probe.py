import sys, time, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import xcube_resampling_b200 as xrs
from xcube_resampling_b200 import _dev, rectify as xrect, synthetic as syn
torch.cuda.set_device(0)
def T(fn, n=3):
    fn(); torch.cuda.synchronize()
    t0=time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize()
    return (time.perf_counter()-t0)/n*1e3
h2d_host = torch.empty((21,4091,4865), dtype=torch.float32, pin_memory=True)
d_in = torch.empty((21,4091,4865), dtype=torch.float32, device="cuda")
d_out = torch.empty((21,5013,7992), dtype=torch.float32, device="cuda")
d2h_host = torch.empty((21,5013,7992), dtype=torch.float32, pin_memory=True)
gb_in, gb_out = h2d_host.numel()*4/1e9, d2h_host.numel()*4/1e9
t=T(lambda: d_in.copy_(h2d_host, non_blocking=True)); print(f"H2D {gb_in:.2f} GB {t:.1f} ms {gb_in/t*1e3:.1f} GB/s")
t=T(lambda: d2h_host.copy_(d_out, non_blocking=True)); print(f"D2H {gb_out:.2f} GB {t:.1f} ms {gb_out/t*1e3:.1f} GB/s")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def both():
    with torch.cuda.stream(s1): d_in.copy_(h2d_host, non_blocking=True)
    with torch.cuda.stream(s2): d2h_host.copy_(d_out, non_blocking=True)
t=T(both); print(f"both concurrently {t:.1f} ms")
# pitched H2D
d_p = torch.empty((21,4091,4896), dtype=torch.float32, device="cuda")
t=T(lambda: d_p[..., :4865].copy_(h2d_host, non_blocking=True)); print(f"H2D pitched view {t:.1f} ms")
# pipeline alone with a trivial process
vals = h2d_host.numpy()
def pipe():
    p = _dev.BandPipeline(vals, (5013,7992), np.float32)
    p.run(lambda src, out: None)
t=T(pipe); print(f"BandPipeline no-op process {t:.1f} ms")
t0=time.perf_counter(); p=_dev.BandPipeline(vals, (5013,7992), np.float32); torch.cuda.synchronize(); print("construct ms", (time.perf_counter()-t0)*1e3)
t0=time.perf_counter(); p.run(lambda s,o: None); print("run ms", (time.perf_counter()-t0)*1e3)
