"""Where does the bf16 gradient error of the encoder side come from?  One config-2 step vs the oracle with
selected GEMMs of the middle chain switched to the exact fp32 path.  Diagnostic (gpurun), not a test."""
import os
import sys

import numpy as np
import pandas as pd
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
from helpers import csr_batch, rel_l2  # noqa: E402
from mmvae_b200 import layers as L  # noqa: E402
from oracle import cmmvae_oracle as O  # noqa: E402
from test_parity_fullshape_gpu import dropout_masks, oracle_spec, state_of  # noqa: E402

O.FAST_CSR = True
G_SCALE = int(os.environ.get("DIAG_G", "60530"))
B = int(os.environ.get("DIAG_B", "1024"))


def run(variant):
    L.set_precision("bf16")
    model, species, _ = bench.build_model(2)
    d = bench.Dims(2)
    G = d.G_HUMAN
    spec = oracle_spec(species, d.H1, d.H2, d.HV, d.Z)
    P = state_of(model)
    model.cuda().train()
    model.configure_optimizers()
    eng = model.engine()
    orig_tc = eng._tc
    if variant == "mid_fp32":
        eng._tc = lambda *dims: orig_tc(*dims) and max(dims) > 2048
    elif variant == "dq_fp32":
        eng._tc = lambda *dims: orig_tc(*dims) and not (len(dims) == 2 and dims == (eng.Hv, 2 * eng.Z))
    crow, col, val = bench.synth_csr(B, G, 0.05, seed=900)
    eps = torch.randn(B, d.Z, generator=torch.Generator().manual_seed(0))
    cpu_m, gpu_m = dropout_masks("human", (d.H1, d.H2), B, 0.1, 40)
    ref = O.train_step(spec, P, {}, "human", crow, col, val, eps, 1.0, dropout_masks=cpu_m)
    L.inject_noise(eps.cuda())
    L.inject_dropout_masks(gpu_m)
    model.training_step((csr_batch(crow, col, val, G), pd.DataFrame({"cell": np.arange(B)}), "human"), 0)
    torch.cuda.synchronize()
    params = dict(model.named_parameters())
    out = {}
    for k, g in ref["grads"].items():
        out[k] = rel_l2(params[f"module.{k}"].grad.detach().cpu().numpy(), g.numpy())
    print(f"== {variant}")
    for k, v in out.items():
        if k.endswith("weight") or "bn" in k:
            print(f"   {k:60s} {v:.4f}")
    del model, eng
    torch.cuda.empty_cache()


for v in sys.argv[1:] or ["base", "dq_fp32", "mid_fp32"]:
    run(v)
