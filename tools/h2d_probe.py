"""How fast / how asynchronous is a DMA out of a page-locked-in-place numpy array vs a torch pinned block?"""
import time
import numpy as np
import torch
from mmvae_b200 import ops
from mmvae_b200.synth import synth_csr
import scipy.sparse as sp
import pandas as pd
from mmvae_b200.feed import StagedCSRBatches

n = 3_730_000
src = np.random.default_rng(0).integers(0, 60000, size=4 * n, dtype=np.int32)
dst = torch.empty(n, dtype=torch.int32, device="cuda")
pin = torch.empty(n, dtype=torch.int32, pin_memory=True)
st = torch.cuda.Stream()
t = time.perf_counter(); ops.host_register(src); print("register %.1f MB: %.2f ms" % (src.nbytes / 1e6, (time.perf_counter() - t) * 1e3))
for name, fn in (("registered numpy", lambda k: ops.h2d_async(dst, src[k * n:(k + 1) * n], 4 * n, stream=st)),
                 ("torch pinned", lambda k: dst.copy_(pin, non_blocking=True))):
    for rep in range(2):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(st):
            e0.record()
            t = time.perf_counter()
            for k in range(4):
                fn(k)
            host = (time.perf_counter() - t) * 1e3 / 4
            e1.record()
        torch.cuda.synchronize()
        print(f"{name:18s} host {host:.3f} ms/call   device {e0.elapsed_time(e1) / 4:.3f} ms/copy  "
              f"{4 * n / (e0.elapsed_time(e1) / 4 * 1e-3) / 1e9:.1f} GB/s")
t = time.perf_counter(); ops.host_unregister(src); print("unregister: %.2f ms" % ((time.perf_counter() - t) * 1e3))

B, G = 1024, 60664
crow, col, val = synth_csr(4 * B, G, 0.06, 1)
chunk = sp.csr_matrix((val, col, crow), shape=(4 * B, G))
frame = pd.DataFrame({"cell": np.arange(4 * B)})
def source():
    while True:
        yield chunk, frame
for kw in (dict(workers=0, ahead=4, pin_chunks=True), dict(workers=6)):
    feed = StagedCSRBatches(source(), B, device="cuda", **kw)
    it = iter(feed)
    for _ in range(8):
        next(it)
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(40):
        next(it)
    torch.cuda.synchronize()
    print(kw, "feed alone: %.3f ms/batch" % ((time.perf_counter() - t) * 1e3 / 40))
    feed.close()
