#!/bin/bash
# Install the UNMODIFIED reference (Roth-Lab/phyclone 0.5.1) into baseline/_ref for the CPU arm of bench.py
# (`bench.py --impl reference`, cpu_baseline.kind = "reference").  Build container only: /root/reference does not
# exist on the GPU box, baseline/_ref is git-ignored but travels there with the gpurun snapshot.
# pip needs a writable source tree (it builds a wheel), so it installs from a copy; --no-deps because the one
# missing dependency, rustworkx (Rust, no wheel, no network), is served at run time by oracle/refshim.
set -euo pipefail
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
[ -d /root/reference ] || { echo "no /root/reference here: nothing to install"; exit 0; }
TMP="$(mktemp -d)"
cp -r /root/reference "$TMP/ref"
rm -rf "$ROOT/baseline/_ref"
mkdir -p "$ROOT/baseline"
python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
  --target "$ROOT/baseline/_ref" "$TMP/ref"
rm -rf "$TMP"
echo "installed: $(ls "$ROOT/baseline/_ref")"
