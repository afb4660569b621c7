"""Repeat the plain configs[1] run in several modes and report loss spikes (looking for rare garbage reads)."""
import os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tools.cond_bench as cb
from mmvae_b200 import layers as L, ops
from mmvae_b200.synth import synth_csr

batches = [tuple(torch.from_numpy(a).cuda() for a in synth_csr(cb.B, cb.G, 0.06, 10 + i)) for i in range(4)]
eps = torch.randn(cb.B, cb.Z, device="cuda")


def run(mode, reps=3, steps=68):
    L.set_precision("fp32" if mode == "fp32" else "bf16")
    ops.set_pdl(mode != "nopdl")
    for rep in range(reps):
        model = cb.build(False)
        model.cuda().train()
        model.configure_optimizers()
        eng = model.engine()
        if mode == "nodrop":
            for plans in list(eng.enc_plan.values()) + [eng.vaeenc_plan]:
                for lp in plans:
                    lp.p_drop = 0.0
        if mode == "nospmmtc":
            eng.spmm_tc = False
        traj = []
        for t in range(steps):
            crow, col, val = batches[0 if mode == "onebatch" else t % 4]
            rec = eng.train_step("human", crow, col, val, int(col.numel()), 1.0, eps=None if mode == "freshnoise" else eps)
            if mode == "sync":
                torch.cuda.synchronize()
            s = eng.scalars(rec)
            traj.append((s["loss"], s["kl_loss"], s["grad_norms/expert_human"]))
        big = [(i, f"{x[0]:.3g}", f"{x[1]:.3g}") for i, x in enumerate(traj) if x[0] > 1.2e7]
        print(mode, rep, "spikes:", big[:8], flush=True)
    ops.set_pdl(True)


for mode in (sys.argv[1:] or ["base", "onebatch", "nopdl", "fp32", "nodrop", "nospmmtc", "freshnoise"]):
    run(mode)
