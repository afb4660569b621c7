#!/usr/bin/env bash
# Install the UNMODIFIED reference (zdebruine/MMVAE, package `cmmvae`) into baseline/_ref (git-ignored; it travels
# to the GPU box with the gpurun snapshot).  Used only by `bench.py --impl reference` / `torch_cuda_baseline` and
# by tests/golden/make_golden.py -- never by the product package.
#
# /root/reference is read-only and its setup.cfg says `packages = find:` + `package_dir = =src` WITHOUT
# `[options.packages.find] where = src`, so a plain `pip install /root/reference` builds a wheel with no modules
# (outcome recorded in DESIGN.md).  The install therefore runs from a copy under /tmp whose only change is that
# missing packaging stanza; no reference source file is touched.
set -euo pipefail
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
SRC="${1:-/root/reference}"
TMP="$(mktemp -d /tmp/cmmvae_ref.XXXXXX)"
cp -r "$SRC"/. "$TMP"/
printf '\n[options.packages.find]\nwhere = src\n' >> "$TMP/setup.cfg"
rm -rf "$ROOT/baseline/_ref"
python -m pip install --no-index --no-build-isolation --find-links /opt/wheelhouse --no-deps \
    --target "$ROOT/baseline/_ref" "$TMP" >/dev/null
# the reference's own unit tests travel with the install (git-ignored like the rest of baseline/_ref):
# tests/test_reference_unit_tests_gpu.py runs them, unmodified, against this repo's modules on the GPU
mkdir -p "$ROOT/baseline/_ref/reference_tests"
cp "$SRC/tests/test_components.py" "$SRC/tests/test_tag_log_dict.py" "$ROOT/baseline/_ref/reference_tests/"
# (two of those tests read a label list relative to the reference's repository root)
mkdir -p "$ROOT/baseline/_ref/reference_tests/src/cmmvae/data/conditional_layers"
cp "$SRC/src/cmmvae/data/conditional_layers/unique_assays.csv" "$ROOT/baseline/_ref/reference_tests/src/cmmvae/data/conditional_layers/"
rm -rf "$TMP"
python - <<PY
import sys; sys.path.insert(0, "$ROOT/baseline/_ref")
import cmmvae.modules, cmmvae.config, cmmvae.constants
print("installed", cmmvae.__file__)
PY
