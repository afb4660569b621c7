"""CPU-side checks: the C-ABI library builds for sm_100a, loads and exports every symbol that
include/cmmvae_b200.h declares (no compute calls without a GPU); host logic of the module mirror
(config validation, log tagging, KL annealing, optimizer map, YAML class_path loading).  The shape /
key tests follow the reference's own tests (tests/test_components.py, tests/test_tag_log_dict.py)."""
import ctypes
import os

import pandas as pd
import pytest
import torch
import torch.nn as nn

import mmvae_b200.compat as compat
from mmvae_b200 import _lib
from mmvae_b200.models import CMMVAEModel, tag_log_dict
from mmvae_b200.models.cmmvae_model import convert_to_flat_list_and_map
from mmvae_b200.modules import CLVAE, CMMVAE
from mmvae_b200.modules.base import (ConditionalLayer, Encoder, Expert, Experts, FCBlock, FCBlockConfig,
                                     KLAnnealingFn, LinearKLAnnealingFn)
from mmvae_b200.modules.base.components import collect_species_files, is_iterable

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ---------------------------------------------------------------------------------------- C ABI
def test_library_builds_and_exports_every_declared_symbol():
    path = _lib.build_library()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    names = _lib.exported_symbols_in_header()
    assert len(names) >= 20 and "cmmvae_decoder_mse_fused" in names and "cmmvae_csr_linear_fwd" in names
    for n in names:
        assert hasattr(lib, n), n
    assert lib.cmmvae_abi_version() == 2
    lib.cmmvae_last_error.restype = ctypes.c_char_p
    assert isinstance(lib.cmmvae_last_error(), bytes)


def test_every_exported_entry_point_is_declared_in_the_header():
    """the other direction: the library exports no `cmmvae_*` function that include/cmmvae_b200.h does not declare
    (the header is the contract a reference-side binding is written against)"""
    import shutil
    import subprocess
    nm = shutil.which("nm")
    if nm is None:
        pytest.skip("nm not available")
    out = subprocess.run([nm, "-D", "--defined-only", _lib.build_library()], capture_output=True, text=True, check=True)
    exported = {line.split()[-1] for line in out.stdout.splitlines() if line.split() and line.split()[-1].startswith("cmmvae_")}
    declared = set(_lib.exported_symbols_in_header())
    assert exported - declared == set(), sorted(exported - declared)
    assert declared - exported == set(), sorted(declared - exported)


def test_argument_validation_happens_before_any_launch():
    lib = _lib.load()
    lib.cmmvae_last_error.restype = ctypes.c_char_p
    rc = lib.cmmvae_csr_linear_fwd(None, None, None, 0, 10, 8, None, 0, None, None, None)
    assert rc == -1 and b"bad shape" in lib.cmmvae_last_error()
    rc = lib.cmmvae_gemm_bf16_tc(ctypes.c_void_p(16), 7, 0, ctypes.c_void_p(32), 8, 0, 4, 4, 4, None, 0, 0,
                                 ctypes.c_void_p(64), None, 4, None, None)
    assert rc == -1 and b"multiples of 8" in lib.cmmvae_last_error()


def test_sass_contains_tcgen05_and_tma():
    import subprocess
    out = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "UTCHMMA" in out and "UTMALDG" in out and "LDTM" in out


def test_forward_on_cpu_tensors_fails_loudly():
    block = FCBlock(FCBlockConfig(layers=[10, 20, 30], dropout_rate=0.5))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        block(torch.randn(5, 10))


# ------------------------------------------------------------------- reference unit tests, host side
@pytest.mark.parametrize("obj,expect", [([1, 2, 3], True), ("string", True), (123, False), ({"k": "v"}, True),
                                        (None, False)])
def test_is_iterable(obj, expect):
    assert is_iterable(obj) is expect


def test_fc_block_config_and_block_properties():
    cfg = FCBlockConfig(layers=[10, 20, 30], dropout_rate=0.5)
    block = FCBlock(cfg)
    assert cfg.n_layers == 2 and block.input_dim == 10 and block.output_dim == 30
    assert FCBlock(FCBlockConfig(layers=[10, 20, 30], return_hidden=False)).can_bypass
    assert not FCBlock(FCBlockConfig(layers=[10, 20, 30], return_hidden=True)).can_bypass
    cfg = FCBlockConfig(layers=[10, 20, 30], dropout_rate=[0.5, 0.3], use_batch_norm=True)
    assert cfg.layers == [10, 20, 30] and cfg.dropout_rate == [0.5, 0.3] and cfg.use_batch_norm == [True, True]
    assert all(issubclass(a, nn.ReLU) for a in FCBlockConfig(layers=[10, 20], activation_fn=nn.ReLU).activation_fn)
    assert FCBlockConfig(layers=[7]).layers == [7, 7]
    names = [n for n, _ in FCBlock(FCBlockConfig([4, 5], use_batch_norm=True, use_layer_norm=True,
                                                 activation_fn=nn.ReLU, dropout_rate=0.1)).fc_layers[0].named_children()]
    assert names == ["lin", "bn", "ln", "af", "dr"]


@pytest.mark.parametrize("kwargs", [dict(layers=[-10, 20]), dict(layers=[10, 20, 30], dropout_rate=[0.5]),
                                    dict(layers=(10, 20)), dict(layers=[10, 20], use_batch_norm=[1]),
                                    dict(layers=[10, 20], activation_fn=[int])])
def test_fc_block_config_validation_errors(kwargs):
    with pytest.raises(ValueError):
        FCBlockConfig(**kwargs)


def test_conditional_layer_and_species_files(tmp_path):
    csv = tmp_path / "unique_assays.csv"
    pd.DataFrame(["10x 5' v1", "10x 3' v3", "microwell-seq", "a.b"]).to_csv(csv, header=False, index=False)
    layer = ConditionalLayer(batch_key="assay", conditions_path=str(csv), fc_block_config=FCBlockConfig(layers=[10]))
    assert layer.batch_key == "assay" and len(layer.conditions) == 4 and "a_b" in layer.conditions
    (tmp_path / "d" / "shared").mkdir(parents=True)
    (tmp_path / "d" / "human").mkdir()
    (tmp_path / "d" / "shared" / "unique_expression_assay.csv").write_text("x\n")
    (tmp_path / "d" / "human" / "unique_expression_assay.csv").write_text("x\n")
    (tmp_path / "d" / "human" / "unique_expression_sex.csv").write_text("x\n")
    found = collect_species_files(str(tmp_path / "d"), ["assay", "sex"])
    assert set(found["shared"]) == {"assay"} and set(found["human"]) == {"sex"}


def test_encoder_expert_experts_construction():
    enc = Encoder(latent_dim=5, fc_block_config=FCBlockConfig(layers=[10]))
    assert enc.n_layers == 1 and enc.var_eps == 1e-4
    cfg = FCBlockConfig(layers=[10, 20])
    e1, e2 = Expert("expert1", cfg, cfg), Expert("expert2", cfg, cfg)
    assert e1.id == "expert1" and e1.encoder is not None and e1.decoder is not None
    with pytest.raises(NotImplementedError):
        e1()
    experts = Experts([e1, e2])
    assert len(experts) == 2 and "expert1" in experts and experts.labels == {"expert1": 0, "expert2": 1}


def test_tag_log_dict():
    d = {"loss": torch.tensor(1.0), "accuracy": torch.tensor(0.9)}
    assert tag_log_dict(d) == d
    assert set(tag_log_dict(d, tags=["modelA", "experiment1"], sep="_", key_pos="first")) == {
        "loss_modelA_experiment1", "accuracy_modelA_experiment1"}
    assert set(tag_log_dict(d, tags=["modelA", "experiment1"], sep="_", key_pos="last")) == {
        "modelA_experiment1_loss", "modelA_experiment1_accuracy"}
    assert tag_log_dict({}) == {}
    with pytest.raises(ValueError):
        tag_log_dict({"loss": 1.0}, key_pos="invalid")


def test_kl_annealing_matches_reference_schedule():
    fn = LinearKLAnnealingFn(min_kl_weight=0.1, max_kl_weight=1.0, warmup_steps=1, climax_steps=4)
    seen = []
    for _ in range(7):
        seen.append(fn.kl_weight)
        fn.step()
    assert seen == pytest.approx([0.1, 0.1, 0.325, 0.55, 0.775, 1.0, 1.0])
    const = KLAnnealingFn(0.5)
    const.step()
    assert const.kl_weight == 0.5


def test_flat_list_and_map():
    flat = []
    m = convert_to_flat_list_and_map({"experts": {"human": "a", "mouse": "b"}, "vae": "c", "adversarials": {1: "d"}}, flat)
    assert flat == ["a", "b", "c", "d"]
    assert m == {"experts": {"human": 0, "mouse": 1}, "vae": 2, "adversarials": {1: 3}}


def test_yaml_class_path_tree_instantiates_under_cmmvae_names():
    """the reference's jsonargparse trees (configs/model/*.yaml) load against this package"""
    compat.install_as_cmmvae()
    import cmmvae.models  # noqa: F401  (alias)
    tree = {
        "class_path": "cmmvae.models.CMMVAEModel",
        "init_args": {
            "kl_annealing_fn": {"class_path": "cmmvae.modules.base.KLAnnealingFn", "init_args": {"kl_weight": 1.0}},
            "adv_weight": 0,
            "autograd_config": {"class_path": "cmmvae.config.AutogradConfig", "init_args": {
                "vae_gradient_clip": {"class_path": "cmmvae.config.GradientClipConfig",
                                      "init_args": {"val": 10, "algorithm": "norm"}}}},
            "module": {"class_path": "cmmvae.modules.CMMVAE", "init_args": {
                "vae": {"class_path": "cmmvae.modules.CLVAE", "init_args": {
                    "latent_dim": 16,
                    "encoder_config": {"class_path": "cmmvae.modules.base.FCBlockConfig", "init_args": {
                        "layers": [32, 24], "use_batch_norm": True, "activation_fn": "torch.nn.ReLU",
                        "return_hidden": True}},
                    "decoder_config": {"class_path": "cmmvae.modules.base.FCBlockConfig", "init_args": {
                        "layers": [16, 24, 32], "activation_fn": "torch.nn.ReLU"}}}},
                "experts": {"class_path": "cmmvae.modules.base.Experts", "init_args": {"experts": [
                    {"class_path": "cmmvae.modules.base.Expert", "init_args": {
                        "id": "human",
                        "encoder_config": {"class_path": "cmmvae.modules.base.FCBlockConfig", "init_args": {
                            "layers": [200, 64, 32], "dropout_rate": [0.1, 0.1], "use_batch_norm": True,
                            "activation_fn": "torch.nn.ReLU"}},
                        "decoder_config": {"class_path": "cmmvae.modules.base.FCBlockConfig", "init_args": {
                            "layers": [32, 64, 200], "activation_fn": "torch.nn.ReLU"}}}}]}},
                "adversarials": None}}}}
    model = compat.instantiate(tree)
    assert isinstance(model, CMMVAEModel) and model.adv_weight == 1.0   # 0 -> 1.0, reference quirk
    assert len(model.module.adversarials) == 0                            # attribute always exists
    keys = set(model.state_dict())
    assert "module.experts.human.encoder.fc_layers.0.lin.weight" in keys
    assert "module.vae.encoder.mean_encoder.weight" in keys
    assert "module.experts.human.encoder.fc_layers.1.bn.running_mean" in keys
    assert tuple(model.state_dict()["module.experts.human.decoder.fc_layers.1.lin.weight"].shape) == (200, 64)
    # He init: zero biases, fan_out-scaled weights
    w = model.module.experts["human"].encoder.fc_layers[0].lin
    assert float(w.bias.abs().max()) == 0.0
    assert float(w.weight.std()) == pytest.approx((2.0 / 64) ** 0.5, rel=0.1)


def test_reference_state_dict_loads_by_name():
    """golden init state (from the unmodified reference) loads strictly: identical names and shapes"""
    from helpers import GoldenCase, build_b200_model
    import tempfile
    gc = GoldenCase("two_species_adv")
    with tempfile.TemporaryDirectory() as tmp:
        model = build_b200_model(gc, tmp)
    res = model.load_state_dict({f"module.{k}": v for k, v in gc.state("init").items()}, strict=True)
    assert not res.missing_keys and not res.unexpected_keys


def test_bench_reference_arm_contract_on_cpu():
    """`bench.py --impl reference` (the CPU arm the driver runs beside the GPU arm) needs no GPU and prints one
    JSON line with the contract's keys; a bounded sample keeps it to seconds here."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--batch", "32",
                          "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.strip().splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "cells/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["vs_baseline"] is None and d["scaling"] == "weak"
    have_ref = os.path.exists(os.path.join(root, "baseline", "_ref", "cmmvae", "models", "cmmvae_model.py"))
    assert d["cpu_baseline"]["kind"] == ("reference" if have_ref else "port")   # the unmodified reference when installed
    assert d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["steps"] == d["cpu_baseline"]["steps_run"] == 1        # the line reports the steps it actually ran
    assert d["e2e"] == {"value": d["value"], "unit": "cells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_label_lookup_matches_per_cell_dict_walk():
    """adversary labels (cmmvae_model.py:103-115): the factorised lookup must give exactly the ids the
    reference's per-cell dict walk gives (int64, row index in the human csv) and raise KeyError on a value the
    csv does not hold."""
    import types
    import numpy as np
    from mmvae_b200.models.cmmvae_model import CMMVAEModel
    from mmvae_b200.modules.base.components import Adversarial
    saved = dict(Adversarial.labels)
    try:
        Adversarial.labels.clear()
        Adversarial.labels.update({"assay": {f"assay_{i}": i for i in range(8)},
                                   "dataset_id": {f"ds_{i}": i for i in range(272)}})
        rng = np.random.default_rng(5)
        meta = pd.DataFrame({"assay": [f"assay_{i}" for i in rng.integers(0, 8, 777)],
                             "dataset_id": [f"ds_{i}" for i in rng.integers(0, 272, 777)],
                             "unrelated": np.arange(777)})
        me = types.SimpleNamespace(_label_ring={}, _label_slot=0)
        got = CMMVAEModel._labels(me, meta, "cpu")
        assert list(got) == ["assay", "dataset_id"]
        for c, table in Adversarial.labels.items():
            want = torch.tensor([table[v] for v in meta[c].values], dtype=torch.int64)
            assert got[c].dtype == torch.int64 and torch.equal(got[c], want)
        meta.loc[3, "assay"] = "never_seen"
        with pytest.raises(KeyError):
            CMMVAEModel._labels(me, meta, "cpu")
    finally:
        Adversarial.labels.clear()
        Adversarial.labels.update(saved)


def test_linear_kl_schedule_reproduces_the_reference_run():
    """the KL weights the unmodified reference fed into steps 0..3 of the `two_species_adv` golden run
    (LinearKLAnnealingFn(0.1, 1.0, warmup 1, climax 4), one step() per training step)"""
    from helpers import GoldenCase
    from mmvae_b200.modules.base import KLAnnealingFn, LinearKLAnnealingFn
    gc = GoldenCase("two_species_adv")
    fn = LinearKLAnnealingFn(min_kl_weight=0.1, max_kl_weight=1.0, warmup_steps=1, climax_steps=4)
    for t in range(gc.n_steps):
        assert fn.kl_weight == pytest.approx(gc.step(t)["kl_weight"], rel=1e-12)
        fn.step()
    for _ in range(20):
        fn.step()
    assert fn.kl_weight == 1.0 and (fn.m, fn.b) == (pytest.approx(0.225), 0.1)
    const = KLAnnealingFn(0.5)
    const.step()
    assert const.kl_weight == 0.5
    const.kl_weight = 0.25          # assignable, as in the reference
    assert const.kl_weight == 0.25
    warm = LinearKLAnnealingFn(0.1, 1.0, warmup_steps=3, climax_steps=4)
    warm.kl_weight = 0.7            # an assigned value survives until the warm-up is over
    warm.step(); warm.step()
    assert warm.kl_weight == 0.7
    warm.step()
    assert warm.kl_weight == pytest.approx(0.1)


def test_clip_config_objects_behave_like_the_reference_records():
    from mmvae_b200.config import AutogradConfig, GradientClipConfig
    c = GradientClipConfig(val=10, algorithm="norm")
    assert tuple(c) == (10, "norm") and (c.val, c.algorithm) == (10, "norm") and bool(c)
    assert tuple(GradientClipConfig()) == (None, None) and bool(GradientClipConfig())
    with pytest.raises(ValueError):
        GradientClipConfig(1.0, "l2")
    a = AutogradConfig(c)
    assert a.adversarial_gradient_clip is c and a.vae_gradient_clip is None and a.expert_gradient_clip is None
    a = AutogradConfig(expert_gradient_clip=c)
    assert a.expert_gradient_clip is c and a.adversarial_gradient_clip is None


def test_custom_op_layer_is_registered():
    """the C-ABI entry points are reachable as torch.ops.cmmvae.* (CUDA kernels only: a CPU tensor is
    refused by the dispatcher, there is nothing to fall back to)"""
    import mmvae_b200.torch_ops as T
    for name in T.OPS:
        assert hasattr(torch.ops.cmmvae, name), name
    with pytest.raises(NotImplementedError, match="CPU"):
        torch.ops.cmmvae.cast_bf16(torch.randn(4), torch.empty(4, dtype=torch.bfloat16))


def test_conditional_host_plan_groups_rows_by_value():
    """CondBank.host_plan (pure host): every (batch key, cell) pair lands in exactly one tile of the slot of ITS value
    -- what ConditionalLayer.forward's dict of row lists does (components.py:388-411) --, tiles hold <= 32 rows,
    single-tile slots are flagged, '.' in a value is formatted like the reference (components.py:355-365), unknown
    values raise KeyError like the ModuleDict lookup, and a non-shared key without species raises RuntimeError"""
    import numpy as np
    from mmvae_b200.conditional import CondBank, ROWS
    bank = CondBank.__new__(CondBank)
    bank.names = ["assay", "donor", "species"]
    bank.kind = {"assay": "shared", "donor": "per_species", "species": "block"}
    bank.block_slot = {("species", "human"): 0, ("species", "mouse"): 1}
    bank.tables = {("assay", None): ({f"a_{i}": i for i in range(3)}, np.arange(2, 5, dtype=np.int32)),
                   ("donor", "human"): ({f"d{i}": i for i in range(50)}, np.arange(5, 55, dtype=np.int32)),
                   ("donor", "mouse"): ({f"d{i}": i for i in range(7)}, np.arange(55, 62, dtype=np.int32))}
    B = 150
    rng = np.random.default_rng(3)
    assay = [f"a.{i}" if i == 1 else f"a_{i}" for i in rng.integers(0, 3, B)]      # 'a.1' must resolve to key 'a_1'
    donor = [f"d{i}" for i in rng.integers(0, 50, B)]
    meta = pd.DataFrame({"assay": assay, "donor": pd.Categorical(donor)})
    tiles, rows, present, ranges, multi = bank.host_plan(meta, "human", B)
    want = {0: {0: set(range(B))}}                                       # key index -> slot -> rows
    want[0] = {}
    for b in range(B):
        want[0].setdefault(2 + int(assay[b].replace(".", "_")[2:]), set()).add(b)
    want[1] = {}
    for b in range(B):
        want[1].setdefault(5 + int(donor[b][1:]), set()).add(b)
    want[2] = {0: set(range(B))}
    got, n_tiles_of = {0: {}, 1: {}, 2: {}}, {}
    for slot, start, count, w in tiles:
        c, sole = int(w) & 0xFFFF, int(w) >> 16
        assert 0 < count <= ROWS
        got[c].setdefault(int(slot), []).extend(rows[start:start + count].tolist())
        n_tiles_of[int(slot)] = n_tiles_of.get(int(slot), 0) + 1
        assert sole in (0, 1)
    for c in (0, 1, 2):
        assert {s: set(r) for s, r in got[c].items()} == want[c]
        assert all(len(r) == len(set(r)) for r in got[c].values())
    for slot, start, count, w in tiles:
        assert (int(w) >> 16) == int(n_tiles_of[int(slot)] == 1)
    assert sorted(present.tolist()) == sorted(n_tiles_of) and set(multi.tolist()) == {s for s, n in n_tiles_of.items() if n > 1}
    for c, (lo, n) in enumerate(ranges):                                 # contiguous tile range per key
        assert all((int(t[3]) & 0xFFFF) == c for t in tiles[lo:lo + n]) and n == sum((int(t[3]) & 0xFFFF) == c for t in tiles)
    with pytest.raises(KeyError):
        bank.host_plan(pd.DataFrame({"assay": ["zzz"] * B, "donor": donor}), "human", B)
    with pytest.raises(RuntimeError, match="species"):
        bank.host_plan(meta, None, B)
    with pytest.raises(KeyError):
        bank.host_plan(meta, "rat", B)
