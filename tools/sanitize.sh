#!/bin/bash
# compute-sanitizer passes over one reduced-shape launch of every kernel template (race / sync / memory
# checking of the hand-rolled mbarrier / TMEM / cluster protocols).  Runs on a B200:
#   gpurun --timeout 1500 -- 'bash tools/sanitize.sh'
# Logs land in gpurun_out/sanitize_<tool>.log; tools/sanitize_summarize.py condenses them into profiles/.
# Each tool gets its own time box (racecheck / initcheck serialise heavily); a pass that runs out of time is reported as such.
set -u
OUT=${OUT:-gpurun_out}
mkdir -p "$OUT"
SAN=${SAN:-/usr/local/cuda/bin/compute-sanitizer}
# small-shape tests that between them launch every kernel template: generic conv (TMA 2D / im2col / gather, staged and direct
# epilogues, k-block widths), stem (v3 + multi-frame + generic), temporal-halo, 3x3 halo, bottleneck tail (both modes), CTA-pair
# (plain / residual / half-width tail), fused pools, max/avg pools, preprocess, segment, head, tf32
SEL=${SEL:-'conv_matches_fp32_reference or producers_agree or half_width_k or narrow_channels or stem_conv_folded or stem_fused_temporal or conv3_fused_temporal or bottleneck_tail or cta_pair or maxpool_is_exact or same_padding or avgpool or preprocess_small or segment_golden or add_magnitude'}
FILES=${FILES:-tests/test_gpu_kernels.py}
run() {
  tool=$1; limit=$2; shift 2
  echo "=== $tool (limit ${limit}s) $(date -u +%H:%M:%S)"
  timeout "$limit" "$SAN" --tool "$tool" "$@" --error-exitcode 66 --log-file "$OUT/sanitize_$tool.raw" \
    python -m pytest $FILES -x -q -m gpu -k "$SEL" -p no:cacheprovider > "$OUT/sanitize_$tool.pytest" 2>&1
  rc=$?
  { echo "tool=$tool rc=$rc (0 clean, 66 sanitizer errors, 124 time box exceeded)"; tail -n 3 "$OUT/sanitize_$tool.pytest";
    grep -c "^========= " "$OUT/sanitize_$tool.raw" 2>/dev/null | sed 's/^/raw report lines: /';
    grep -E "ERROR SUMMARY|Invalid|Race reported|hazard|Barrier error|Uninitialized" "$OUT/sanitize_$tool.raw" | sort | uniq -c | sort -rn | head -n 40; } > "$OUT/sanitize_$tool.log"
  cat "$OUT/sanitize_$tool.log"
}
run memcheck ${T_MEM:-420}
run synccheck ${T_SYNC:-300}
run racecheck ${T_RACE:-420} --racecheck-report analysis
run initcheck ${T_INIT:-300}
