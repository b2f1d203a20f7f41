#!/bin/bash
# developer A/B helper: tools/ab_libs.sh "ENV=.. lib.so" ...  (each argument: optional VAR=VAL words, then the library path)
for spec in "$@"; do
  lib="${spec##* }"; envs="${spec% *}"; [ "$envs" == "$spec" ] && envs=""
  echo "== $spec"
  env $envs TF21_LIB=$PWD/$lib python - <<'PY'
import importlib, os, sys, torch
sys.path.insert(0, os.getcwd())
tf = importlib.import_module("twenty-first_b200")
dev = tf.device; dev.init(0)
x = torch.randint(0, 2**62, (256 << 20,), dtype=torch.int64, device="cuda:0")
def t(fn, it=7):
    for _ in range(3): fn()
    torch.cuda.synchronize(); best = 1e9
    for _ in range(it):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); best = min(best, a.elapsed_time(b))
    return best
ms = t(lambda: dev.ntt_(x, 1 << 20, 1, False)); msi = t(lambda: dev.ntt_(x, 1 << 20, 1, True))
dev.profile_enable(True)
for _ in range(3): dev.ntt_(x, 1 << 20, 1, False)
torch.cuda.synchronize()
agg = {}
for name, m in dev.profile_read(): agg.setdefault(name, []).append(m)
dev.profile_enable(False)
print(f"ntt 2^20 x256 fwd {ms:.3f} ms ({16*256*2**20/ms/1e6/6448.1*100:.1f}%) inv {msi:.3f} ms", {k: round(sum(v)/len(v), 4) for k, v in agg.items()})
if os.environ.get("AB_MERKLE"):
    leafs = torch.randint(0, 2**62, (5 << 24,), dtype=torch.int64, device="cuda:0"); nodes = torch.zeros(10 << 24, dtype=torch.int64, device="cuda:0")
    mm = t(lambda: dev.merkle_build(leafs, nodes), 3); print(f"merkle 2^24 {mm:.3f} ms ({2**24/mm/1e6:.2f} G leaves/s)")
PY
done
