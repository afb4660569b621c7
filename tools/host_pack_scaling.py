"""Host packer scaling on this box: aggregate batches/s of cmmvae_host_slice_rows with T threads, each thread
cycling over its own pageable 4-batch chunk (the bench's e2e host side), writing into its own pinned blocks."""
import sys, time, threading
import numpy as np
import torch
from mmvae_b200 import ops
from mmvae_b200.synth import synth_csr

B, G = 1024, 60664
crow, col, val = synth_csr(4 * B, G, 0.06, 1)
nmax = int(max(crow[(k + 1) * B] - crow[k * B] for k in range(4)))
pin = torch.cuda.is_available()


def worker(tid, n_iter, out):
    c, i, v = crow.copy(), col.copy(), val.copy()      # own pageable chunk (120 MB)
    oc = torch.empty(B + 1, dtype=torch.int32, pin_memory=pin).numpy()
    oi = torch.empty(nmax + 8, dtype=torch.int16, pin_memory=pin).numpy().view(np.uint16)
    ov = torch.empty(nmax + 8, dtype=torch.float32, pin_memory=pin).numpy()
    bar.wait()
    t0 = time.perf_counter()
    for k in range(n_iter):
        lo = (k % 4) * B
        ops.host_slice_rows(c, i, v, lo, lo + B, G, oc, oi, ov)
    out[tid] = (time.perf_counter() - t0) / n_iter


for T in [int(a) for a in sys.argv[1:]] or [1, 2, 4, 8]:
    bar = threading.Barrier(T)
    out = [0.0] * T
    th = [threading.Thread(target=worker, args=(t, 24, out)) for t in range(T)]
    [t.start() for t in th]
    [t.join() for t in th]
    per = sum(out) / T
    nnz = int(crow[-1]) / 4
    print(f"T={T:3d}  {per * 1e3:6.2f} ms/batch/thread  {T / per:8.0f} batches/s  "
          f"{T / per * nnz * 14 / 1e9:7.1f} GB/s (14 B/nnz)", flush=True)
