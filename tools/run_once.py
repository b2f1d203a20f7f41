"""Run a few launches of one workload for ncu (developer helper)."""
import importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
tf = importlib.import_module("twenty-first_b200")
dev = tf.device
dev.init(0)
cuda = torch.device("cuda:0")
what = sys.argv[1] if len(sys.argv) > 1 else "ntt20"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
if what.startswith("ntt:"):  # ntt:<log2n>:<batch>[:<width>]
    parts = what.split(":")
    l2, batch = int(parts[1]), int(parts[2])
    w = int(parts[3]) if len(parts) > 3 else 1
    x = torch.randint(0, 2**63 - 1, ((batch * w) << l2,), dtype=torch.int64, device=cuda)
    for _ in range(reps):
        dev.ntt_(x, 1 << l2, w, False)
elif what == "ntt20":
    cols = int(os.environ.get("COLS", "64"))
    x = torch.randint(0, 2**63 - 1, (cols << 20,), dtype=torch.int64, device=cuda)
    for _ in range(reps):
        dev.ntt_(x, 1 << 20, 1, False)
elif what == "ntt16":
    x = torch.randint(0, 2**63 - 1, (1024 << 16,), dtype=torch.int64, device=cuda)
    for _ in range(reps):
        dev.ntt_(x, 1 << 16, 1, False)
elif what == "lde26":
    vals = torch.randint(0, 2**63 - 1, (3 << 22,), dtype=torch.int64, device=cuda)
    out = torch.zeros(3 << 26, dtype=torch.int64, device=cuda)
    g = tf.BFieldElement.generator()
    for _ in range(reps):
        dev.coset_lde(vals, 3, g, 1 << 26, g, out)
elif what == "ntt10":
    x = torch.randint(0, 2**63 - 1, (16384 << 10,), dtype=torch.int64, device=cuda)
    for _ in range(reps):
        dev.ntt_(x, 1 << 10, 1, False)
elif what == "merkle":
    h = int(os.environ.get("H", "22"))
    leafs = torch.randint(0, 2**63 - 1, (5 << h,), dtype=torch.int64, device=cuda)
    nodes = torch.zeros(10 << h, dtype=torch.int64, device=cuda)
    for _ in range(reps):
        dev.merkle_build(leafs, nodes)
elif what == "lde":
    vals = torch.randint(0, 2**63 - 1, (3 << 18,), dtype=torch.int64, device=cuda)
    out = torch.zeros(3 << 22, dtype=torch.int64, device=cuda)
    g = tf.BFieldElement.generator()
    for _ in range(reps):
        dev.coset_lde(vals, 3, g, 1 << 22, g, out)
torch.cuda.synchronize()
