#!/bin/sh
# ab_variants.sh : run bench.py once per kernel-variant library in exp/ (built with -D macros) and print value + roofline.frac
# usage (on the GPU box): sh tools/ab_variants.sh "<bench args>" exp/lib_*.so
ARGS="$1"; shift
for lib in "$@"; do
  out=$(DLA_B200_LIB="$PWD/$lib" timeout 300 python bench.py $ARGS --skip-cpu 2>/dev/null | tail -1)
  echo "$lib $(echo "$out" | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("value %.1f  e2e %.1f  frac %.3f  ms/step %.1f" % (d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["ms_per_step"]))' 2>/dev/null || echo FAILED)"
done
