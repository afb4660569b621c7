"""The reference's OWN unit tests (tests/test_components.py, tests/test_tag_log_dict.py of zdebruine/MMVAE), run
unmodified against this repo's modules: ``mmvae_b200.compat.install_as_cmmvae()`` makes ``import cmmvae...``
resolve to the B200 implementation, and the default device is ``cuda`` so that the forward passes those tests run
(FCBlock, ConditionalLayer, Encoder, Expert) execute on the sm_100a kernels -- the package has no CPU path.
The test files are copied into baseline/_ref/reference_tests by tools/install_reference.sh (git-ignored; they travel
to the GPU box with the snapshot); without them the test is skipped."""
import importlib.util
import inspect
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_TESTS = os.path.join(ROOT, "baseline", "_ref", "reference_tests")


def _load(name):
    path = os.path.join(REF_TESTS, name + ".py")
    if not os.path.exists(path):
        pytest.skip("reference unit tests not installed (tools/install_reference.sh)")
    spec = importlib.util.spec_from_file_location("reference_" + name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("name", ["test_components", "test_tag_log_dict"])
def test_reference_unit_tests_pass_on_cuda(name, tmp_path, monkeypatch):
    import mmvae_b200.compat as compat
    for m in [m for m in sys.modules if m == "cmmvae" or m.startswith("cmmvae.")]:
        del sys.modules[m]
    compat.install_as_cmmvae()
    monkeypatch.chdir(REF_TESTS)         # two tests read src/cmmvae/data/... relative to the reference's repo root
    torch.set_default_device("cuda")
    try:
        mod = _load(name)
        tests = [(n, f) for n, f in inspect.getmembers(mod, inspect.isfunction) if n.startswith("test_")]
        assert len(tests) >= (25 if name == "test_components" else 3), len(tests)
        failed = []
        for n, f in tests:
            try:
                f()
            except Exception as e:  # noqa: BLE001
                failed.append((n, f"{type(e).__name__}: {e}"))
        assert not failed, failed
    finally:
        torch.set_default_device("cpu")
