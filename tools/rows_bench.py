"""Developer timing helper: Tip5::hash_varlen over table rows, row-major (tf21_tip5_hash_rows_dev) against column-major
(tf21_tip5_hash_columns_dev), next to plain hash_10 -- permutations per second."""
import importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
tf = importlib.import_module("twenty-first_b200")
dev = tf.device
dev.init(0)
cuda = torch.device("cuda:0")


def t(fn, it=5):
    for _ in range(2): fn()
    torch.cuda.synchronize(); best = 1e9
    for _ in range(it):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); best = min(best, a.elapsed_time(b))
    return best


n = 1 << 22
inp = torch.randint(0, 2**62, (10 * n,), dtype=torch.int64, device=cuda)
out = torch.zeros(5 * n, dtype=torch.int64, device=cuda)
ms = t(lambda: dev.tip5_hash_10(inp, out))
print(f"hash_10 x2^22: {ms:.3f} ms  {n/ms/1e6:.3f} G permutations/s")
st = torch.randint(0, 2**62, (16 * n,), dtype=torch.int64, device=cuda)
ms = t(lambda: dev.tip5_permute_(st))
print(f"permutation x2^22: {ms:.3f} ms  {n/ms/1e6:.3f} G permutations/s")
del inp, st
for n_rows, n_cols in ((1 << 20, 100), (1 << 20, 30), (1 << 18, 400), (1 << 22, 9)):
    data = torch.randint(0, 2**62, (n_rows * n_cols,), dtype=torch.int64, device=cuda)
    out = torch.zeros(5 * n_rows, dtype=torch.int64, device=cuda)
    perms = n_rows * ((n_cols + 1 + 9) // 10)
    ms_c = t(lambda: dev.tip5_hash_columns(data, n_rows, n_cols, out), 3)
    ms_r = t(lambda: dev.tip5_hash_rows(data, n_cols, out), 3)
    print(f"{n_rows} rows x {n_cols} cols: column-major {ms_c:.3f} ms = {perms/ms_c/1e6:.3f} G perm/s ({8*n_rows*n_cols/ms_c/1e6:.0f} GB/s read) | "
          f"row-major {ms_r:.3f} ms = {perms/ms_r/1e6:.3f} G perm/s ({8*n_rows*n_cols/ms_r/1e6:.0f} GB/s read)")
    del data, out
