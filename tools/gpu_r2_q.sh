#!/bin/bash
# compute-sanitizer over every kernel family (round 2 additions included)
mkdir -p gpurun_out
out=gpurun_out/r2_compute_sanitizer.txt
echo "# compute-sanitizer on tools/sanitize.py (round 2: + geometry-2 scan single / multi-pass with cp.async feed, per-CTA scratch regions, device sink, batched queries, warp-per-subject end cells in passes), one B200" > $out
for tool in memcheck racecheck synccheck; do
  echo "== $tool" >> $out
  timeout 1500 compute-sanitizer --tool $tool python tools/sanitize.py 2>&1 | grep -E "ok|ERROR SUMMARY|RACECHECK SUMMARY|Error|hazard|=========  " | head -30 >> $out
done
cat $out
