"""Per-kernel device times of the configs[3] LDE via the library's own event profiler (developer helper)."""
import importlib, os, sys, collections
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
tf = importlib.import_module("twenty-first_b200")
dev = tf.device
dev.init(0)
li, lo = 22, 26
vals = torch.randint(0, 2**62, (3 << li,), dtype=torch.int64, device="cuda:0")
out = torch.zeros(3 << lo, dtype=torch.int64, device="cuda:0")
g = tf.BFieldElement.generator()
for _ in range(2):
    dev.coset_lde(vals, 3, g, 1 << lo, g, out)
torch.cuda.synchronize()
dev.profile_enable(True)
for _ in range(3):
    dev.coset_lde(vals, 3, g, 1 << lo, g, out)
torch.cuda.synchronize()
prof = dev.profile_read()
dev.profile_enable(False)
n = len(prof) // 3
for name, ms in prof[:n]:
    print(f"{name:40s} {ms:.3f} ms")
print("total", sum(ms for _, ms in prof) / 3)
