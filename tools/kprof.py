"""Per-kernel device times of a few NTT shapes via the library's own event profiler (developer helper)."""
import importlib, os, sys, collections
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
tf = importlib.import_module("twenty-first_b200")
dev = tf.device
dev.init(0)
shapes = [(10, 65536, 1), (16, 1024, 1), (20, 256, 1), (24, 8, 1), (22, 4, 3)]
for log2n, batch, w in shapes:
    x = torch.randint(0, 2**62, ((batch * w) << log2n,), dtype=torch.int64, device="cuda:0")
    for inv in (False, True):
        for _ in range(2):
            dev.ntt_(x, 1 << log2n, w, inv)
        torch.cuda.synchronize()
        dev.profile_enable(True)
        for _ in range(3):
            dev.ntt_(x, 1 << log2n, w, inv)
        torch.cuda.synchronize()
        prof = dev.profile_read()
        dev.profile_enable(False)
        agg = collections.OrderedDict()
        for name, ms in prof:
            agg[name] = agg.get(name, 0.0) + ms / 3
        tot = sum(agg.values())
        print(f"2^{log2n} x{batch} w{w} {'inv' if inv else 'fwd'}: total {tot:.3f} ms | " + " | ".join(f"{k} {v:.3f}" for k, v in agg.items()))
    del x
if os.environ.get("AB_MERKLE"):
    leafs = torch.randint(0, 2**62, (5 << 24,), dtype=torch.int64, device="cuda:0"); nodes = torch.zeros(10 << 24, dtype=torch.int64, device="cuda:0")
    for _ in range(2): dev.merkle_build(leafs, nodes)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); dev.merkle_build(leafs, nodes); b.record(); torch.cuda.synchronize()
    print(f"merkle 2^24 {a.elapsed_time(b):.3f} ms")
