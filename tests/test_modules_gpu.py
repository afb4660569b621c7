"""Module-level API on the GPU: the reference's shape tests (tests/test_components.py) against the
B200 modules, and the generic autograd path (module forward -> elbo -> backward) against the oracle."""
import numpy as np
import pandas as pd
import pytest
import torch
import torch.nn as nn

from helpers import GoldenCase, build_b200_model, csr_batch, rel_l2
from oracle import cmmvae_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _fp32():
    from mmvae_b200 import layers as L
    L.set_precision("fp32")
    yield
    L.set_precision("bf16")


def test_fc_block_forward_shapes():
    from mmvae_b200.modules.base import FCBlock, FCBlockConfig
    block = FCBlock(FCBlockConfig(layers=[10, 20, 30], dropout_rate=0.5)).cuda()
    assert block(torch.randn(5, 10).cuda()).shape == (5, 30)
    block = FCBlock(FCBlockConfig(layers=[10, 20, 30], return_hidden=True, activation_fn=nn.ReLU)).cuda()
    out, hidden = block(torch.randn(5, 10).cuda())
    assert out.shape == (5, 30) and [h.shape for h in hidden] == [(5, 20), (5, 30)]


def test_conditional_layer_forward(tmp_path):
    from mmvae_b200.modules.base import ConditionalLayer, FCBlockConfig
    csv = tmp_path / "unique_assays.csv"
    vals = ["10x 5' v1", "10x 3' v3", "microwell-seq", "10x 5' transcription profiling"]
    pd.DataFrame(vals).to_csv(csv, header=False, index=False)
    layer = ConditionalLayer("assay", str(csv), FCBlockConfig(layers=[10])).cuda()
    x = torch.randn(5, 10).cuda()
    meta = pd.DataFrame({"assay": [vals[0], vals[1], vals[2], vals[2], vals[3]]})
    out = layer(x, meta)
    assert out.shape == x.shape
    # row routing is exact: rows 2 and 3 went through the same block
    ref = layer.conditions["microwell-seq"](x[2:4])
    assert torch.allclose(out[2:4], ref, atol=1e-6)


def test_encoder_and_expert_shapes():
    from mmvae_b200.modules.base import Encoder, Expert, FCBlockConfig
    enc = Encoder(latent_dim=5, fc_block_config=FCBlockConfig(layers=[10])).cuda()
    q_m, q_v, latent, hidden = enc(torch.randn(5, 10).cuda())
    assert q_m.shape == (5, 5) and q_v.shape == (5, 5) and latent.shape == (5, 5) and isinstance(hidden, list)
    assert bool((q_v > 0).all())
    expert = Expert("expert1", FCBlockConfig(layers=[10], return_hidden=[True], activation_fn=nn.ReLU),
                    FCBlockConfig(layers=[10])).cuda()
    encoded, hidden = expert.encode(torch.randn(5, 10).cuda())
    assert encoded.shape == (5, 10) and all(h.shape == (5, 10) for h in hidden)
    assert expert.decode(encoded).shape == (5, 10)


@pytest.mark.parametrize("name", ["core_human"])
def test_autograd_path_matches_oracle(name, tmp_path):
    """CMMVAE.forward -> BaseVAE.elbo -> loss.backward() (the reference's own call sequence,
    cmmvae_model.py:157-187) on the B200 modules, against the oracle's gradients."""
    from mmvae_b200 import layers as L
    gc = GoldenCase(name)
    model = build_b200_model(gc, tmp_path)
    model.load_state_dict({f"module.{k}": v for k, v in gc.state("init").items()})
    model.cuda().train()
    s = gc.step(0)
    x = csr_batch(s["crow"], s["col"], s["val"], gc.genes["human"])
    L.inject_noise(s["eps"].cuda())
    qz, pz, z, xhats, hidden = model.module(x, pd.DataFrame({"a": np.arange(gc.dims["B"])}), "human")
    assert xhats["human"].shape == (gc.dims["B"], gc.genes["human"]) and len(hidden) == 1
    ld = model.module.vae.elbo(qz, pz, x, xhats["human"], s["kl_weight"])
    assert set(ld) == {"loss", "recon_loss", "kl_loss", "kl_weight"}
    ld["loss"].backward()
    ref = s["logs"]
    assert float(ld["loss"]) == pytest.approx(ref["loss/training/human"], rel=2e-5)
    assert float(ld["kl_loss"]) == pytest.approx(ref["kl_loss/training/human"], rel=2e-5)
    assert float(qz.mean.mean()) == pytest.approx(ref["Mean/training/human"], rel=1e-4)
    assert float(qz.variance.mean()) == pytest.approx(ref["Variance/training/human"], rel=1e-4)
    params = dict(model.named_parameters())
    for k, g in s["grads"].items():
        mine = params[f"module.{k}"].grad
        assert mine is not None, k
        if k.endswith(".lin.bias") and k.replace(".lin.bias", ".bn.weight") in gc.state("init"):
            continue
        assert rel_l2(mine.cpu().numpy(), g) < 3e-4, (k, rel_l2(mine.cpu().numpy(), g))
    # the species that was not used has no gradient at all
    assert all(p.grad is None for n, p in params.items() if ".experts.mouse." in n)


def test_eval_forward_and_latents(tmp_path):
    gc = GoldenCase("core_human")
    model = build_b200_model(gc, tmp_path)
    model.load_state_dict({f"module.{k}": v for k, v in gc.state("final").items()})
    model.cuda().eval()
    s = gc.step("val")
    from mmvae_b200 import layers as L
    x = csr_batch(s["crow"], s["col"], s["val"], gc.genes["human"])
    L.inject_noise(s["eps"].cuda())
    with torch.no_grad():
        qz, pz, z, xhats, hidden = model.module(x, pd.DataFrame({"a": np.arange(gc.dims["B"])}), "human")
    assert rel_l2(z.cpu().numpy(), s["z"]) < 2e-5
    assert rel_l2(xhats["human"].cpu().numpy(), s["xhat"]) < 2e-5
    L.inject_noise(s["eps"].cuda())
    emb = model.predict_step((x, pd.DataFrame({"a": np.arange(gc.dims["B"])}), "human"), 0)
    zz, meta = emb["z"]
    assert rel_l2(zz.cpu().numpy(), s["z"]) < 2e-5 and (meta["species"] == "human").all()
