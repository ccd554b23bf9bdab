#!/bin/bash
# compute-sanitizer over the chain's kernels (run on the GPU box): memcheck, racecheck, initcheck, synccheck over
# __graft_entry__.smoke() and one short run of modes 0 r (33 stations, overlap mode), 2 m and 3 s.
# Writes gpurun_out/sanitizer_<tag>.txt: one line per (tool, workload) with the error count, then the raw tails.
#   tools/sanitize.sh [tag]
tag=${1:-r2}
out=gpurun_out/sanitizer_$tag.txt
mkdir -p gpurun_out
CS=${CS:-/usr/local/cuda/bin/compute-sanitizer}
: > $out.raw
echo "# compute-sanitizer $($CS --version 2>/dev/null | head -1)  ($(date -u +%FT%TZ), $(nvidia-smi --query-gpu=name --format=csv,noheader | head -1))" > $out
echo "# tool, workload, exit status, ERROR SUMMARY" >> $out
run() {  # tool, label, command...
  local tool=$1 label=$2; shift 2
  local log=$(mktemp)
  timeout 900 $CS --tool $tool --error-exitcode 99 --print-limit 20 "$@" > $log 2>&1
  local rc=$?
  local summ=$(grep -E "ERROR SUMMARY|RACECHECK SUMMARY" $log | tail -1 | sed 's/^=* *//')
  echo "$tool, $label, rc=$rc, ${summ:-no summary line}" >> $out
  { echo "=== $tool / $label (rc=$rc)"; grep -v "^$" $log | tail -25; } >> $out.raw
  rm -f $log
}
for tool in memcheck racecheck initcheck synccheck; do
  run $tool "smoke (0 r, 1 station, 12 blocks)" python -c "import __graft_entry__ as g; g.smoke()"
  run $tool "mode 0 r, 33 stations, 8 blocks, overlap" python tools/sanitize_run.py 0 r 33 8 1
  run $tool "mode 2 m, 3 stations, 3 blocks" python tools/sanitize_run.py 2 m 3 3 0
  run $tool "mode 3 s, 3 stations, 3 blocks" python tools/sanitize_run.py 3 s 3 3 0
done
cat $out.raw >> $out
rm -f $out.raw
cat $out | head -20
