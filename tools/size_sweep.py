"""Developer helper: NTT throughput for every size (fixed total words), BFE and XFE."""
import importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
tf = importlib.import_module("twenty-first_b200")
dev = tf.device
dev.init(0)
HBM = 6448.1e9
total_log = int(os.environ.get("TOTAL_LOG", "27"))
for w in (1, 3):
    for log2n in [int(v) for v in os.environ["SWEEP_SIZES"].split(",")] if os.environ.get("SWEEP_SIZES") else range(3, 27):
        words = 1 << total_log
        n = 1 << log2n
        batch = max(1, words // (n * w))
        x = torch.randint(0, 2**62, (n * w * batch,), dtype=torch.int64, device="cuda:0")
        for _ in range(2):
            dev.ntt_(x, n, w, False)
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(4):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); dev.ntt_(x, n, w, False); b.record(); torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b))
        bytes_ = 16 * n * w * batch
        print(f"w={w} 2^{log2n:2d} x{batch:8d}: {best:8.3f} ms  {bytes_/best/1e6:7.0f} GB/s = {bytes_/best*1e3/HBM*100:5.1f}%", flush=True)
        del x
