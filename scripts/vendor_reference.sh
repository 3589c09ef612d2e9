#!/usr/bin/env bash
# Copies the UNMODIFIED reference model files into the git-ignored baseline/_ref/ so that they travel to the GPU box with
# the gpurun snapshot (the box has no /root/reference). Nothing under baseline/_ref/ is product source or enters history
# (.gitignore lists it); tests and bench legs that need it skip / say "unavailable" when it is absent.
#
#   scripts/vendor_reference.sh [/path/to/reference]
#
# Also invoked by __graft_entry__.build() when /root/reference is mounted.
set -euo pipefail
SRC="${1:-${TITOK_REFERENCE_ROOT:-/root/reference}}"
ROOT="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
DST="$ROOT/baseline/_ref"
if [ ! -d "$SRC/model" ]; then
  echo "vendor_reference: no reference at $SRC (nothing copied)" >&2
  exit 0
fi
rm -rf "$DST"
mkdir -p "$DST"
# only what the tokenizer path and its consumer need: model/ (titok, base, quantizer; losses for the discriminator
# wrapper), train_utils/ (CodebookLogger), configs/ (tiny.yaml)
cp -r "$SRC/model" "$SRC/train_utils" "$SRC/configs" "$DST/"
find "$DST" -name '__pycache__' -type d -prune -exec rm -rf {} +
( cd "$SRC" && find model train_utils configs -type f -name '*.py' -o -type f -name '*.yaml' | sort | xargs sha256sum ) > "$DST/MANIFEST.sha256"
echo "vendor_reference: $(wc -l < "$DST/MANIFEST.sha256") files -> $DST"
