#!/usr/bin/env bash
# Install the UNMODIFIED reference (yang-song/score_sde_pytorch) under baseline/_ref so that
# `bench.py --impl reference` can time the reference's own CPU path on the GPU box's host cores.
#
# The reference is a flat directory of Python modules with no setup.py / pyproject.toml, so
# `pip install --target baseline/_ref /root/reference` has nothing to build ("neither setup.py nor
# pyproject.toml found"); the install is a plain copy of its sources.  baseline/_ref is git-ignored (no reference
# source enters the history) but not gpurun-ignored, so it travels to the GPU box with the snapshot.
#
# The reference's `op/` package JIT-compiles two CUDA extensions at import (op/upfirdn2d.py:10-16,
# op/fused_act.py:11-17); they are built here once into baseline/_ref_ext (TORCH_EXTENSIONS_DIR) so the box
# only has to re-link/verify them.  On CPU tensors the reference never calls them (op/upfirdn2d.py:146-149).
set -euo pipefail
REPO="$(cd "$(dirname "$0")/.." && pwd)"
SRC="${1:-/root/reference}"
DST="$REPO/baseline/_ref"
if [ ! -d "$SRC" ]; then echo "install_ref: $SRC not found (nothing to do on a GPU box: baseline/_ref travels prebuilt)"; exit 0; fi
mkdir -p "$DST"
# sources only: no notebooks / images / VCS data
( cd "$SRC" && tar --exclude=.git --exclude=assets --exclude='*.ipynb' -cf - . ) | ( cd "$DST" && tar -xf - )
echo "install_ref: copied $(find "$DST" -name '*.py' | wc -l) python files to $DST"
# pre-build the JIT extensions for sm_100 (no GPU needed)
export TORCH_EXTENSIONS_DIR="$REPO/baseline/_ref_ext" TORCH_CUDA_ARCH_LIST="10.0"
python - <<'EOF'
import os, sys, types, time
repo = os.path.dirname(os.path.dirname(os.path.abspath(os.environ['TORCH_EXTENSIONS_DIR'])))
sys.path.insert(0, repo)
import bench
t0 = time.time()
bench.import_reference()
print(f'install_ref: reference imports (extensions built/cached in {time.time() - t0:.0f} s)')
EOF
