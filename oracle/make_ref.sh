#!/bin/sh
# Populate oracle/_ref/ with the UNMODIFIED reference (pure Python: nothing to compile) from /root/reference, where it
# lies in the build container.  oracle/_ref/ is git-ignored (the reference's sources never enter this repository's
# history) but not gpurun-ignored, so it travels to the GPU box, where /root/reference does not exist.  Used by:
#   * bench.py --impl reference / cpu_baseline (kind "reference"): create_algorithm("mixed-tile-greedy").run + wq:684-687
#   * tests/test_gpu_cli.py: the reference's own `wq` and sweep programs run over this package's drop-in modules
# Test infrastructure only: nothing under quantization_analysis_b200/ imports it.
set -e
SRC=${1:-/root/reference}
DST="$(cd "$(dirname "$0")" && pwd)/_ref"
if [ ! -d "$SRC/compression_algorithms" ]; then
    echo "make_ref: $SRC not present; keeping whatever $DST holds" >&2
    exit 0
fi
rm -rf "$DST"
mkdir -p "$DST/scripts"
cp "$SRC/quantization_formats.py" "$SRC/hf_model_utils.py" "$SRC/wq" "$DST/"
cp -r "$SRC/compression_algorithms" "$DST/compression_algorithms"
cp "$SRC/scripts/sweep_mixed_tile_threshold.py" "$DST/scripts/"
find "$DST" -name __pycache__ -type d -exec rm -rf {} + 2>/dev/null || true
(cd "$SRC" && find quantization_formats.py hf_model_utils.py wq compression_algorithms scripts/sweep_mixed_tile_threshold.py -type f ! -name '*.pyc' | sort | xargs sha256sum) > "$DST/SHA256SUMS"
echo "make_ref: copied the reference into $DST"
