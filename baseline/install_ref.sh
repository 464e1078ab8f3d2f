#!/bin/sh
# Installs the UNMODIFIED reference (teboli/polyblur, /root/reference) into baseline/_ref so that
# `bench.py --impl reference`, its cpu_baseline leg and its library_gpu leg can import it.  Build container only:
# /root/reference does not exist on the GPU box; baseline/_ref is git-ignored but travels with the snapshot.
# The source tree is read-only and setuptools writes build/ + egg-info next to setup.py, hence the /tmp copy;
# --no-deps because the pinned requirements (torch 1.13, scikit-image ...) are not in the offline wheelhouse --
# the reference runs on this image's torch / numpy, and bench.py stubs `skimage.img_as_float32` (identity).
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${POLYBLUR_REFERENCE:-/root/reference}"
[ -d "$REF/polyblur" ] || { echo "no reference at $REF"; exit 0; }
TMP="$(mktemp -d)"
cp -r "$REF" "$TMP/ref"
rm -rf "$HERE/_ref"
python -m pip install --quiet --no-index --no-build-isolation --find-links /opt/wheelhouse --no-deps \
    --target "$HERE/_ref" "$TMP/ref"
rm -rf "$TMP"
echo "installed the reference into $HERE/_ref"
