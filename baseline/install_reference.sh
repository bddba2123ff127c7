#!/bin/sh
# Offline install of the UNMODIFIED reference (dmme 0.5.2) into the git-ignored baseline/_ref, for `bench.py --impl reference`
# and the on-GPU-box cpu_baseline (kind "reference").  The checkout is read-only and setuptools writes build files into the
# source tree, so the install runs from a throw-away copy; --no-deps because pytorch_lightning / torchmetrics / lmdb are
# neither in the image nor in the wheelhouse (oracle/ref_shim.py registers inert stand-ins for exactly those imports).
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
SRC="${DMME_REFERENCE_ROOT:-/root/reference}"
[ -d "$SRC/src/dmme" ] || { echo "no reference checkout at $SRC"; exit 0; }
TMP="$(mktemp -d)"
cp -r "$SRC" "$TMP/ref"
rm -rf "$HERE/_ref"
python -m pip install --quiet --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse --target "$HERE/_ref" "$TMP/ref"
rm -rf "$TMP"
echo "installed: $(ls "$HERE/_ref")"
