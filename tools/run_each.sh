#!/bin/bash
# Run every selected GPU test in its own process under a timeout (a hung kernel then costs one test, not the call):
#   tools/run_each.sh <log> <per-test seconds> <max timeouts> <pytest selection ...>
log=$1; per=$2; maxto=$3; shift 3
ids=$(python -m pytest "$@" --collect-only -q -m gpu 2>/dev/null | grep "::")
nto=0
: > "$log"
for id in $ids; do
  out=$(timeout "$per" python -m pytest "$id" -q -s -x 2>&1); rc=$?
  echo "=== $id rc=$rc" >> "$log"
  echo "$out" | grep -vE "^$|passed|warnings summary|^\.$" | tail -25 >> "$log"
  if [ $rc -eq 124 ]; then nto=$((nto+1)); echo "TIMEOUT $id" >> "$log"; fi
  if [ $nto -ge $maxto ]; then echo "too many timeouts, stopping" >> "$log"; break; fi
done
grep -c "rc=0" "$log" | sed 's/^/passed: /'; grep "rc=[1-9]" "$log"
