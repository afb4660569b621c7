"""Pin the oracle: it must reproduce every vector the UNMODIFIED reference produced
(tests/golden/*.npz, made by tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from helpers import GoldenCase, rel_l2, untag
from oracle import cmmvae_oracle as O

CASES = ["core_human", "two_species_adv", "human_conditional"]


def bias_feeds_batchnorm(name, state):
    return name.endswith(".lin.bias") and name.replace(".lin.bias", ".bn.weight") in state


@pytest.mark.parametrize("name", CASES)
def test_oracle_reproduces_reference_training(name):
    torch.set_num_threads(2)
    gc = GoldenCase(name)
    spec, P, opt = gc.spec(), gc.state("init"), {}
    for t in range(gc.n_steps):
        s = gc.step(t)
        out = O.train_step(spec, P, opt, s["species"], s["crow"], s["col"], s["val"], s["eps"], s["kl_weight"],
                           labels=s["labels"] if gc.with_adv else None, cond=gc.cond(s, 4242 + t))
        ref_logs = untag(s["logs"], s["species"])
        assert set(ref_logs) == set(out["logs"]), (sorted(ref_logs), sorted(out["logs"]))
        for k, v in ref_logs.items():
            assert out["logs"][k] == pytest.approx(v, rel=2e-5, abs=1e-6), (t, k)
        assert rel_l2(out["z"].numpy(), s["z"]) < 1e-5
        for k, g in s["grads"].items():
            # adversary grads were captured at the discriminator update (its own clip call)
            og = out["grads"][f"discriminator/{k}" if k.startswith("adversarials.") else k].numpy()
            # a Linear bias feeding BatchNorm has an analytically zero gradient: both sides hold
            # only rounding noise there, so those are compared absolutely
            assert rel_l2(og, g) < 1e-4 or np.abs(og - g).max() < 1e-4, (t, k)
    final = gc.state("final")
    for k, v in final.items():
        if k.endswith("num_batches_tracked"):
            assert int(P[k]) == int(v), k
        elif bias_feeds_batchnorm(k, final):
            # d(loss)/d(bias) is analytically 0 there (BatchNorm removes the mean): the reference's
            # gradient is rounding noise that Adam normalises into O(lr) steps -- a random walk no
            # implementation can reproduce, with no effect on any output.  Bounded, not matched.
            assert np.abs(P[k].numpy() - v.numpy()).max() <= 2 * gc.n_steps * 5e-3 * 1.01 + 1e-6, k
        elif k.endswith("bn.running_mean"):
            # the batch mean contains that random-walking bias, scaled by momentum 0.01
            atol = 0.01 * gc.n_steps * 2 * gc.n_steps * 5e-3
            assert np.abs(P[k].numpy() - v.numpy()).max() <= atol + 1e-6, k
        else:
            assert rel_l2(P[k].numpy(), v.numpy()) < 2e-5, k


@pytest.mark.parametrize("name", CASES)
def test_oracle_reproduces_reference_validation(name):
    gc = GoldenCase(name)
    s = gc.step("val")
    out = O.eval_step(gc.spec(), gc.state("final"), s["species"], s["crow"], s["col"], s["val"], s["eps"],
                      kl_weight=untag(s["logs"], s["species"], "validation")["kl_weight"], cond=gc.cond(s, 777))
    ref = untag(s["logs"], s["species"], "validation")
    for k in ("loss", "recon_loss", "kl_loss"):
        assert out["logs"][k] == pytest.approx(ref[k], rel=1e-5)
    assert ref["val_loss"] == pytest.approx(ref["loss"])
    assert rel_l2(out["z"].numpy(), s["z"]) < 1e-5
    assert rel_l2(out["xhat"].numpy(), s["xhat"]) < 1e-5


def test_exact_zero_gradient_for_absent_genes():
    """W1 columns of genes absent from the whole batch get exactly 0 gradient."""
    gc = GoldenCase("core_human")
    s = gc.step(0)
    present = np.zeros(gc.genes["human"], dtype=bool)
    present[s["col"]] = True
    g = s["grads"]["experts.human.encoder.fc_layers.0.lin.weight"]
    assert (~present).sum() > 0
    assert np.all(g[:, ~present] == 0.0)
    spec, P = gc.spec(), gc.state("init")
    out = O.train_step(spec, P, {}, s["species"], s["crow"], s["col"], s["val"], s["eps"], s["kl_weight"])
    og = out["grads"]["experts.human.encoder.fc_layers.0.lin.weight"].numpy()
    assert np.all(og[:, ~present] == 0.0)


def test_other_species_untouched():
    """the species string selects the expert; the other expert is not stepped."""
    gc = GoldenCase("two_species_adv")
    spec, P, opt = gc.spec(), gc.state("init"), {}
    before = {k: v.clone() for k, v in P.items() if k.startswith("experts.mouse.")}
    s = gc.step(0)
    assert s["species"] == "human"
    O.train_step(spec, P, opt, "human", s["crow"], s["col"], s["val"], s["eps"], s["kl_weight"], labels=s["labels"])
    for k, v in before.items():
        assert torch.equal(P[k], v), k
    assert "experts/mouse" not in opt


def test_oracle_fast_csr_path_agrees():
    """the full-size CPU-timing path (torch sparse addmm, as the reference executes it) equals the
    explicit per-nonzero restatement, forward and weight gradient"""
    crow, col, val = O.synth_csr(16, 300, 0.1, seed=4)
    gen = torch.Generator().manual_seed(4)
    W = torch.randn(32, 300, generator=gen).requires_grad_()
    b = torch.randn(32, generator=gen)
    y0 = O.csr_linear(crow, col, val, W, b)
    g0, = torch.autograd.grad(y0.square().sum(), W)
    O.FAST_CSR = True
    try:
        y1 = O.csr_linear(crow, col, val, W, b)
        g1, = torch.autograd.grad(y1.square().sum(), W)
    finally:
        O.FAST_CSR = False
    # fp32 summation order differs between the two: compare against the tensor's scale
    assert (y0 - y1).abs().max() <= 1e-5 * y0.abs().max()
    assert (g0 - g1).abs().max() <= 1e-5 * g0.abs().max()


def test_output_discriminator_restatement_matches_torch_modules():
    """the oracle's output-discriminator step against the reference's own network definition executed by torch:
    nn.Sequential(Linear, Sigmoid, Linear, Sigmoid, Linear, Sigmoid) + binary_cross_entropy(mean) + Adam(lr 1e-3)
    (runners/meta_discriminators.py:33-49,103,131-134; the runner itself needs cluster data and a checkpoint, so
    the pin is the architecture + loss + optimizer it names, run here for three steps on random reconstructions)"""
    import torch.nn as nn
    torch.manual_seed(3)
    G, B = 300, 40
    net = nn.Sequential(nn.Linear(G, 128), nn.Sigmoid(), nn.Linear(128, 64), nn.Sigmoid(), nn.Linear(64, 1), nn.Sigmoid())
    Pd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    optim = torch.optim.Adam(net.parameters(), lr=0.001)
    opt = O.OptState()
    for t in range(3):
        xhat = torch.relu(torch.randn(B, G, generator=torch.Generator().manual_seed(t)))
        label = float(t % 2)
        net.zero_grad()
        loss = torch.nn.functional.binary_cross_entropy(net(xhat), torch.full((B, 1), label), reduction="mean")
        loss.backward()
        want_grads = {k: p.grad.clone() for k, p in net.named_parameters()}
        optim.step()
        got = O.output_discriminator_step(Pd, opt, xhat, label)
        assert got["loss"] == pytest.approx(float(loss), rel=1e-6)
        for k, g in want_grads.items():
            assert torch.allclose(got["grads"][k], g, rtol=1e-5, atol=1e-8), k
        for k, v in net.state_dict().items():
            assert torch.allclose(Pd[k], v, rtol=1e-5, atol=1e-7), (t, k)
