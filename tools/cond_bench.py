"""Cost of the conditional layers at the size configs/model/human_only.yaml ships: 8 batch keys x
Linear(128,128)+LayerNorm per metadata value (donor_id: 4 644 values), parallel selection, concat layer 1024->128,
on top of the BASELINE configs[1] human expert, 1024 cells per step.  Prints one JSON line."""
import json, os, sys, tempfile, time
import numpy as np, pandas as pd, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmvae_b200 import layers as L
from mmvae_b200.synth import synth_csr

G, H1, H2, HV, Z, B = 60664, 1024, 512, 256, 128, 1024
CONDS = {"assay": 8, "dataset_id": 272, "dev_stage": 60, "disease": 30, "donor_id": 4644, "sex": 3, "tissue_general": 40}


def build(conditional: bool):
    from mmvae_b200.config import AutogradConfig, GradientClipConfig
    from mmvae_b200.models import CMMVAEModel
    from mmvae_b200.modules import CLVAE, CMMVAE
    from mmvae_b200.modules.base import ConcatBlockConfig, Expert, Experts, FCBlockConfig, KLAnnealingFn
    relu = torch.nn.ReLU
    torch.manual_seed(0)
    kw = {}
    if conditional:
        tmp = tempfile.mkdtemp()
        os.makedirs(os.path.join(tmp, "human")); os.makedirs(os.path.join(tmp, "shared"))
        for c, n in CONDS.items():
            pd.DataFrame([f"{c}_{i}" for i in range(n)]).to_csv(
                os.path.join(tmp, "human" if c == "dataset_id" else "shared", f"unique_expression_{c}.csv"),
                header=False, index=False)
        kw = dict(conditional_config=FCBlockConfig(layers=[Z], use_layer_norm=True, activation_fn=None),
                  concat_config=ConcatBlockConfig(activation_fn=relu), conditionals_directory=tmp,
                  conditionals=list(CONDS) + ["species"], selection_order=["parallel"])
    experts = Experts([Expert("human", FCBlockConfig([G, H1, H2], dropout_rate=0.1, use_batch_norm=True, activation_fn=relu),
                              FCBlockConfig([H2, H1, G], activation_fn=relu))])
    vae = CLVAE(FCBlockConfig([H2, HV], use_batch_norm=True, activation_fn=relu, return_hidden=True),
                FCBlockConfig([Z, HV, H2], activation_fn=relu), latent_dim=Z, **kw)
    clip = lambda: GradientClipConfig(val=10, algorithm="norm")  # noqa: E731
    return CMMVAEModel(CMMVAE(vae, experts, []), autograd_config=AutogradConfig(clip(), clip(), clip()),
                       kl_annealing_fn=KLAnnealingFn(1.0))


def run(conditional, precision="bf16", steps=60):
    L.set_precision(precision)
    model = build(conditional)
    model.cuda().train()
    model.configure_optimizers()
    eng = model.engine()
    assert eng is not None and (eng.cond is not None) == conditional
    rng = np.random.default_rng(0)
    batches = [tuple(torch.from_numpy(a).cuda() for a in synth_csr(B, G, 0.06, 10 + i)) for i in range(4)]
    metas = [pd.DataFrame({c: [f"{c}_{i}" for i in rng.integers(0, n, B)] for c, n in CONDS.items()}) for _ in range(4)]
    eps = torch.randn(B, Z, device="cuda")
    losses, plan_ms = [], 0.0

    def step(t):
        nonlocal plan_ms
        crow, col, val = batches[t % 4]
        if conditional and t % 8 == 0:      # (timed separately on a sample of steps: train_step repeats it)
            t0 = time.perf_counter(); eng.cond.make_plan(metas[t % 4], "human", B); plan_ms += 8 * (time.perf_counter() - t0)
        return eng.train_step("human", crow, col, val, int(col.numel()), 1.0, eps=eps,
                              metadata=metas[t % 4] if conditional else None)
    for t in range(8):
        step(t)
    torch.cuda.synchronize()
    plan_ms = 0.0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for t in range(steps):
        rec = step(8 + t)
    e1.record()
    torch.cuda.synchronize()
    s = eng.scalars(rec)
    out = {"ms_per_step": e0.elapsed_time(e1) / steps, "loss": s["loss"], "recon": s["recon_loss"]}
    if conditional:
        out.update(plan_ms=plan_ms * 1e3 / steps, slots=eng.cond.n_slots, present=eng.cond.plan["n_present"],
                   tiles=eng.cond.plan["n_tiles"], bank_mb=eng.cond.p.numel() * 4 / 1e6)
        eng.timers = {}
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        pl = eng.cond.plan
        z = torch.randn(B, Z, device="cuda")
        d = torch.randn(B, 8 * Z, device="cuda")
        ns = torch.zeros(1, dtype=torch.float64, device="cuda")
        # device time of the three phases: R calls each captured into a CUDA graph (stream launches of these short
        # kernels are bound by the ~50 us the host needs per ctypes call, which says nothing about the kernels)
        from mmvae_b200 import ops
        R = 20
        phases = {"fwd_ms": lambda: eng.cond.forward(z, eng.ws, True), "bwd_ms": lambda: eng.cond.backward(d, eng.ws),
                  "norm_adam_ms": lambda: (eng.cond.add_norm_sq(ns), eng.cond.clip_adam(ns, 10.0))}
        ops.set_pdl(False)
        st = torch.cuda.Stream()
        for name, fn in phases.items():
            fn()
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=st):
                for _ in range(R):
                    fn()
            g.replay()
            torch.cuda.synchronize()
            ev[0].record(); g.replay(); ev[1].record()
            torch.cuda.synchronize()
            out[name] = ev[0].elapsed_time(ev[1]) / R
        ops.set_pdl(True)
    return out


if __name__ == "__main__":
    res = {"plain": run(False), "conditional": run(True)}
    if os.environ.get("COND_FP32"):
        res["conditional_fp32"] = run(True, "fp32", steps=10)
    print(json.dumps(res))
