#!/bin/bash
# developer A/B helper: tools/quick_bench.py (ntt lde) with each library variant given as argument
for lib in "$@"; do echo "== $lib"; TF21_LIB=$PWD/$lib python tools/quick_bench.py ntt lde 2>&1 | grep -v launches; done
