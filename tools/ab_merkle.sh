for lib in "$@"; do echo "== $lib"; TF21_LIB=$PWD/$lib python tools/quick_bench.py merkle 2>&1 | grep merkle; done
