#!/usr/bin/env bash
# TEST / BENCH INFRASTRUCTURE — makes the UNMODIFIED reference package importable on the GPU box.
#
# The reference (hokiepete/dynlab) is pure Python; "building" it is placing its package where
# the snapshot that travels to the GPU box can see it.  This recipe copies
# /root/reference/src/dynlab byte for byte into oracle/_ref/dynlab (git-ignored: the reference's
# sources never enter this repository's history; NOT gpurun-ignored: the copy ships with the
# snapshot exactly like our own built .so).  `pip install --target` is not used because the
# reference's build backend (hatchling) is not in the offline wheelhouse; the result would be
# the same four files.  oracle/ref_loader.py imports the copy after shimming its two missing
# third-party imports (contourpy, pathos).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
SRC="${DYNLAB_REFERENCE_SRC:-/root/reference/src/dynlab}"
DST="$HERE/_ref"
if [ ! -d "$SRC" ]; then
  echo "build_ref: $SRC not present (GPU box?) - keeping whatever is in $DST" >&2
  exit 0
fi
rm -rf "$DST/dynlab"
mkdir -p "$DST"
cp -r "$SRC" "$DST/dynlab"
find "$DST" -name __pycache__ -type d -prune -exec rm -rf {} +
( cd "$SRC/.." && find dynlab -name '*.py' -type f | sort | xargs sha256sum ) > "$DST/SHA256SUMS"
( cd "$DST" && sha256sum -c --quiet SHA256SUMS )
echo "build_ref: $(wc -l < "$DST/SHA256SUMS") reference files -> $DST/dynlab (unmodified, checksums verified)"
