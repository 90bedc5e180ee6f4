#!/bin/bash
# oracle/make_ref.sh — stage the UNMODIFIED reference modules of the hot path for the reference arms of bench.py.
#
# The reference (eminorhan/tae) is pure Python with no packaging (no setup.py / pyproject), so it cannot be pip-installed
# into baseline/_ref as the base contract describes; its hot path is two files.  This recipe copies them, byte for byte,
# from where they lie under /root/reference into oracle/_ref/ — which is git-ignored (reference sources never enter this
# repository's history) but NOT gpurun-ignored, so the copy travels to the GPU box, where /root/reference does not exist.
#   oracle/_ref/tae.py         <- /root/reference/tae.py         (model, factories)
#   oracle/_ref/util/misc.py   <- /root/reference/util/misc.py   (add_weight_decay, adjust_learning_rate, NativeScaler)
# Users: bench.py `--impl reference` (CPU arm, kind "reference") and bench.py's `gpu_reference` legs (the same modules on
# the same B200: bf16 autocast eager / torch.compile, fp16 + GradScaler as shipped).  Test infrastructure only — nothing
# under tae_b200/ imports it.  A SHA-256 manifest is written next to the copies so a stale or edited copy is detectable.
set -e
REF=${1:-/root/reference}
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/_ref"
if [ ! -f "$REF/tae.py" ] || [ ! -f "$REF/util/misc.py" ]; then
  echo "make_ref: $REF/tae.py or util/misc.py not found (GPU box: the prebuilt oracle/_ref is used as is)" >&2
  exit 0
fi
mkdir -p "$OUT/util"
cp "$REF/tae.py" "$OUT/tae.py"
cp "$REF/util/misc.py" "$OUT/util/misc.py"
: > "$OUT/util/__init__.py"
( cd "$OUT" && sha256sum tae.py util/misc.py > MANIFEST.sha256 )
( cd "$REF" && sha256sum tae.py util/misc.py ) | diff -q - "$OUT/MANIFEST.sha256" > /dev/null
echo "make_ref: staged $(wc -l < "$OUT/tae.py") + $(wc -l < "$OUT/util/misc.py") lines under $OUT (sha256 verified)"
