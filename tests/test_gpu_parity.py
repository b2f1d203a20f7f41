"""GPU parity tests proper: the CUDA path, called through the C ABI (libtf21.so), against the CPU
oracle on identical seeded inputs, against the reference's golden vectors, and -- at full
BASELINE.json sizes -- through size-independent properties.  Bit-exact everywhere (integer work)."""
import importlib
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

P = 0xFFFFFFFF00000001


@pytest.fixture(scope="module")
def tf():
    mod = importlib.import_module("twenty-first_b200")
    mod.check(mod.lib.tf21_init(0))
    return mod


def _h(xs):
    return [int(x, 16) if isinstance(x, str) else int(x) for x in xs]


def rnd(seed, count):
    import oracle as o

    return o.splitmix64_words(seed, count)


def adversarial(n, w=1):
    """all-zero, all p-1, unit impulses, lanes 0xffffffff00000000"""
    total = n * w
    sets = [np.zeros(total, dtype=np.uint64), np.full(total, P - 1, dtype=np.uint64),
            np.full(total, 0xFFFFFFFF00000000, dtype=np.uint64)]
    for pos in {0, total - 1, total // 2}:
        e = np.zeros(total, dtype=np.uint64)
        e[pos] = 0xFFFFFFFF  # raw one
        sets.append(e)
    return sets


# ---- field primitives -------------------------------------------------------------------------------
def test_device_field_primitives(tf):
    import torch

    cuda = torch.device("cuda:0")
    edge = [0, 1, 2, 0xFFFFFFFF, 0x100000000, 0xFFFFFFFF00000000, P - 1, P - 2, P - 0xFFFFFFFF, 1 << 63,
            0x3FFFFFFFC, (1 << 32) - 2]
    a_list, b_list = [], []
    for x in edge:
        for y in edge:
            a_list.append(x)
            b_list.append(y)
    ra, rb = rnd(11, 20000), rnd(12, 20000)
    a = np.concatenate([np.array(a_list, dtype=np.uint64), ra])
    b = np.concatenate([np.array(b_list, dtype=np.uint64), rb])
    da = torch.from_numpy(a.view(np.int64)).to(cuda)
    db = torch.from_numpy(b.view(np.int64)).to(cuda)
    out = torch.zeros_like(da)

    def run(op, xa=da, xb=db):
        tf.check(tf.lib.tf21_selftest_field_dev(op, xa.data_ptr(), xb.data_ptr(), out.data_ptr(), xa.numel(), None))
        torch.cuda.synchronize()
        return [int(v) for v in out.cpu().numpy().view(np.uint64)]

    ai, bi = [int(v) for v in a], [int(v) for v in b]
    assert run(0) == [(x + y) % P for x, y in zip(ai, bi)]
    assert run(1) == [(x - y) % P for x, y in zip(ai, bi)]
    assert run(2) == [(x * y) % P for x, y in zip(ai, bi)]
    # arbitrary (non-canonical) u64 operands for mul / canon / weak add / reduce96
    wa = np.concatenate([np.array([(1 << 64) - 1, (1 << 64) - 2, P, P + 1, (1 << 64) - (1 << 32)], dtype=np.uint64),
                         np.random.default_rng(5).integers(0, 1 << 64, 20000, dtype=np.uint64)])
    wb = np.random.default_rng(6).integers(0, 1 << 64, wa.size, dtype=np.uint64)
    wb[:5] = np.array([(1 << 64) - 1, P, 3, (1 << 64) - 1, (1 << 64) - (1 << 32)], dtype=np.uint64)
    dwa = torch.from_numpy(wa.view(np.int64)).to(cuda)
    dwb = torch.from_numpy(wb.view(np.int64)).to(cuda)
    out = torch.zeros_like(dwa)
    wai, wbi = [int(v) for v in wa], [int(v) for v in wb]
    assert run(2, dwa, dwb) == [(x * y) % P for x, y in zip(wai, wbi)]
    assert run(3, dwa, dwb) == [x % P for x in wai]
    wbc = np.array([v % P for v in wbi], dtype=np.uint64)
    dwbc = torch.from_numpy(wbc.view(np.int64)).to(cuda)
    assert run(4, dwa, dwbc) == [(x + int(y)) % P for x, y in zip(wai, wbc)]
    assert run(5, dwa, dwb) == [(x + ((y & 0xFFFFFFFF) << 64)) % P for x, y in zip(wai, wbi)]
    # second-generation lazy primitives (any u64 first operand; second operand <= p)
    wbp = wbc.copy()
    wbp[:3] = np.array([P, P - 1, 0], dtype=np.uint64)  # t = p itself is allowed by gl_addl
    dwbp = torch.from_numpy(wbp.view(np.int64)).to(cuda)
    assert run(6, dwa, dwbp) == [(x + int(y)) % P for x, y in zip(wai, wbp)]
    assert run(7, dwa, dwb) == [x % P for x in wai]
    assert run(8, dwa, dwbc) == [(x - int(y)) % P for x, y in zip(wai, wbc)]
    assert run(9, dwa, dwbp) == [(x - int(y)) % P for x, y in zip(wai, wbp)]
    for s in list(range(3, 96, 3)) + [1, 31, 65, 95]:
        assert run(100 + s, dwa, dwb) == [(x << s) % P for x in wai], f"shift {s}"


# ---- NTT --------------------------------------------------------------------------------------------
def test_ntt_reference_kats(tf, oracle, kats):
    for name in ("bfield_basic", "bfield_max", "bfield_len32"):
        k = kats["ntt"][name]
        x = oracle.to_raw(k["input_values"])
        orig = x.copy()
        tf.ntt(x)
        assert [int(v) for v in oracle.to_values(x)] == k["expected_values"], name
        tf.intt(x)
        assert np.array_equal(x, orig)
    k = kats["ntt"]["xfield_basic"]
    x = oracle.to_raw(np.array(k["input_values"], dtype=np.uint64))
    orig = x.copy()
    tf.ntt(x)
    assert oracle.to_values(x).tolist() == k["expected_values"]
    tf.intt(x)
    assert np.array_equal(x, orig)


@pytest.mark.parametrize("log2n", list(range(0, 25)))
def test_bfe_ntt_matches_oracle(tf, oracle, log2n):
    n = 1 << log2n
    inputs = [rnd(100 + log2n, n)] + (adversarial(n) if log2n <= 12 or log2n == 20 else [])
    for x in inputs:
        want = x.copy()
        assert oracle.ntt(want, 1) == 0
        got = x.copy()
        tf.ntt(got)
        assert np.array_equal(got, want)
        assert (got < np.uint64(P)).all()
        want_i = x.copy()
        oracle.intt(want_i, 1)
        got_i = x.copy()
        tf.intt(got_i)
        assert np.array_equal(got_i, want_i)
        tf.intt(got)
        assert np.array_equal(got, x)


def sprinkled(seed, count):
    """words whose high half is 0xffffffff in SOME lanes of a warp: the optimistic canonicalisation of the 1024-point
    passes (csrc/ntt_fast.cuh dft_opt_stage) takes its slow path in warps where only a few lanes need it"""
    rng = np.random.default_rng(seed)
    a = rnd(seed, count)
    a[rng.random(count) < 1 / 29] = np.uint64(P - 1)
    b = np.zeros(count, dtype=np.uint64)
    b[rng.random(count) < 1 / 61] = np.uint64(P - 1)
    c = (rnd(seed + 1, count) & np.uint64(0xFFFFFFFF)) | np.uint64(0xFFFFFFFE00000000)
    c[rng.random(count) < 1 / 7] = np.uint64(P - 1)
    return [a, b, c]


@pytest.mark.parametrize("log2n,w", [(10, 1), (10, 3), (11, 1), (12, 1), (16, 1), (20, 1), (20, 3), (21, 1)])
def test_ntt_words_near_p_in_some_lanes(tf, oracle, log2n, w):
    n = 1 << log2n
    for x in sprinkled(900 + log2n + w, n * w):
        want = x.copy()
        assert oracle.ntt(want, w) == 0
        got = x.copy().reshape(n, w) if w > 1 else x.copy()
        tf.ntt(got)
        assert np.array_equal(got.reshape(-1), want)
        assert (got.reshape(-1) < np.uint64(P)).all()
        want_i = x.copy()
        oracle.intt(want_i, w)
        got_i = x.copy().reshape(n, w) if w > 1 else x.copy()
        tf.intt(got_i)
        assert np.array_equal(got_i.reshape(-1), want_i)


@pytest.mark.parametrize("log2n", list(range(0, 23)))
def test_xfe_ntt_matches_oracle(tf, oracle, log2n):
    n = 1 << log2n
    inputs = [rnd(200 + log2n, 3 * n)] + (adversarial(n, 3) if log2n <= 10 else [])
    for flat in inputs:
        want = flat.copy()
        oracle.ntt(want, 3)
        got = flat.copy().reshape(n, 3)
        tf.ntt(got)
        assert np.array_equal(got.reshape(-1), want)
        want_i = flat.copy()
        oracle.intt(want_i, 3)
        got_i = flat.copy().reshape(n, 3)
        tf.intt(got_i)
        assert np.array_equal(got_i.reshape(-1), want_i)


def test_ntt_edge_cases_and_errors(tf):
    # ntt.rs:471-498 (empty, length one, 0-1-0) and the panics of ntt.rs:135-137
    api = importlib.import_module("twenty-first_b200.api")
    e = np.zeros(0, dtype=np.uint64)
    tf.ntt(e)
    tf.intt(e)
    one = np.array([12345], dtype=np.uint64)
    tf.ntt(one)
    assert int(one[0]) == 12345
    tf.ntt(e)
    for bad in (3, 12, 1000):
        with pytest.raises(tf.Tf21Error) as ei:
            tf.ntt(np.zeros(bad, dtype=np.uint64))
        assert ei.value.code == tf.E_LEN_NOT_POW2
    assert tf.lib.tf21_ntt(None, 1 << 33, 1, 1) == tf.E_LEN_TOO_LARGE
    assert tf.lib.tf21_ntt(None, 8, 2, 1) == tf.E_BAD_ARG
    assert api is not None


@pytest.mark.parametrize("log2n,width,batch", [(3, 1, 5), (10, 1, 33), (10, 3, 7), (12, 1, 9), (16, 1, 4), (20, 1, 3), (11, 3, 5), (13, 3, 3), (17, 1, 3), (21, 3, 2)])
def test_batched_ntt_matches_oracle(tf, oracle, log2n, width, batch):
    api = importlib.import_module("twenty-first_b200.api")
    n = 1 << log2n
    x = rnd(300 + log2n + batch, n * width * batch)
    for inverse in (False, True):
        want = x.copy()
        assert oracle.ntt_batch(want, n, width, batch, inverse) == 0
        got = x.copy()
        api.ntt_batch(got, n, width, inverse)
        assert np.array_equal(got, want)


def test_ntt_linearity_and_impulse_at_2_20(tf, oracle):
    n = 1 << 20
    a, b = rnd(1, n), rnd(2, n)
    s = np.array([oracle.bfe_add(int(x), int(y)) for x, y in zip(a[:64], b[:64])], dtype=np.uint64)
    fa, fb = a.copy(), b.copy()
    tf.ntt(fa)
    tf.ntt(fb)
    # NTT(a)+NTT(b) == NTT(a+b) checked on a sparse vector to keep the CPU side cheap
    sp_a = np.zeros(n, dtype=np.uint64)
    sp_b = np.zeros(n, dtype=np.uint64)
    sp_a[:64] = a[:64]
    sp_b[:64] = b[:64]
    sp_s = np.zeros(n, dtype=np.uint64)
    sp_s[:64] = s
    tf.ntt(sp_a)
    tf.ntt(sp_b)
    tf.ntt(sp_s)
    idx = np.arange(0, n, 4099)
    for i in idx:
        assert int(sp_s[i]) == oracle.bfe_add(int(sp_a[i]), int(sp_b[i]))
    # impulse at position 1 -> powers of omega (raw-one times omega^i)
    imp = np.zeros(n, dtype=np.uint64)
    imp[1] = 0xFFFFFFFF
    tf.ntt(imp)
    omega = oracle.primitive_root_of_unity(n)
    for i in (0, 1, 2, 12345, n - 1):
        assert int(imp[i]) == oracle.bfe_mod_pow(omega, i)


# ---- coset ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("log_coeffs,log_order,width", [(0, 0, 1), (0, 3, 1), (3, 3, 1), (5, 9, 1), (10, 10, 3), (8, 12, 3), (12, 16, 1), (14, 21, 1), (13, 17, 3), (16, 20, 3), (18, 22, 1)])
def test_coset_evaluate_interpolate_match_oracle(tf, oracle, log_coeffs, log_order, width):
    nc, order = 1 << log_coeffs, 1 << log_order
    coeffs = rnd(400 + log_coeffs * 31 + log_order, nc * width)
    offset = int(rnd(77 + log_order, 1)[0])
    rc, want = oracle.coset_evaluate(coeffs, width, offset, order)
    assert rc == 0
    poly = tf.Polynomial(coeffs.reshape(nc, 3) if width == 3 else coeffs)
    got = poly.fast_coset_evaluate(offset, order)
    assert np.array_equal(got.reshape(-1), want)
    rc, want_c = oracle.coset_interpolate(want, width, offset)
    got_c = tf.Polynomial.fast_coset_interpolate(offset, got).coefficients
    assert np.array_equal(got_c.reshape(-1), want_c)
    assert np.array_equal(got_c.reshape(-1)[: nc * width], coeffs)
    assert not got_c.reshape(-1)[nc * width:].any()


def test_coset_evaluate_degree_rules(tf, oracle):
    coeffs = rnd(3, 8)
    g = tf.BFieldElement.generator()
    with pytest.raises(tf.Tf21Error) as ei:
        tf.Polynomial(coeffs).fast_coset_evaluate(g, 4)
    assert ei.value.code == tf.E_ORDER_LE_DEGREE
    coeffs[4:] = 0  # trailing zeros do not count (Polynomial::new strips them)
    got = tf.Polynomial(coeffs).fast_coset_evaluate(g, 4)
    rc, want = oracle.coset_evaluate(coeffs, 1, g, 4)
    assert rc == 0 and np.array_equal(got, want)
    # the zero polynomial evaluates to zeros on any domain, including order 0
    z = tf.Polynomial(np.zeros(4, dtype=np.uint64)).fast_coset_evaluate(g, 8)
    assert not z.any()


@pytest.mark.parametrize("log_in,log_out,width", [(0, 0, 1), (0, 4, 3), (4, 4, 1), (6, 10, 3), (10, 14, 1), (12, 16, 3), (16, 20, 3), (18, 22, 3), (13, 13, 1),
                                                       (13, 16, 1), (11, 16, 3), (17, 21, 1), (14, 23, 1), (12, 18, 3), (13, 19, 1)])
def test_coset_lde_matches_oracle(tf, oracle, log_in, log_out, width):
    n_in, n_out = 1 << log_in, 1 << log_out
    values = rnd(500 + log_in + log_out, n_in * width)
    g = tf.BFieldElement.generator()
    g_in, g_out = g, (g if log_in % 2 == 0 else tf.BFieldElement.new(49))
    rc, coeffs = oracle.coset_interpolate(values, width, g_in)
    assert rc == 0
    rc, want = oracle.coset_evaluate(coeffs, width, g_out, n_out)
    assert rc == 0
    got = tf.Polynomial.coset_lde(values.reshape(n_in, 3) if width == 3 else values, g_in, n_out, g_out)
    assert np.array_equal(got.reshape(-1), want)


# ---- Tip5 --------------------------------------------------------------------------------------------
def test_tip5_reference_kats(tf, oracle, kats):
    t = kats["tip5"]
    pre = np.zeros(10, dtype=np.uint64)
    for i in range(6):
        pre[i:i + 5] = tf.Tip5.hash_10(pre)
    assert tf.Digest.to_hex(tf.Tip5.hash_10(pre)) == t["hash10_snapshot"]["hex"]

    acc = np.zeros(5, dtype=np.uint64)
    for i in range(20):
        d = tf.Tip5.hash_varlen(oracle.to_raw(list(range(i))) if i else np.zeros(0, dtype=np.uint64))
        acc = np.array([oracle.bfe_add(int(a), int(b)) for a, b in zip(acc, d)], dtype=np.uint64)
    assert tf.Digest.to_hex(acc) == t["hash_varlen_sum"]["hex"]

    s = np.array(_h(t["raw_snapshot"]["state_raw"]), dtype=np.uint64)
    tf.Tip5.permutation(s)
    assert [int(v) for v in s[:5]] == _h(t["raw_snapshot"]["expected_first5_raw"])

    s = oracle.to_raw(_h(t["degenerate"]["state_values"]))
    tf.Tip5.permutation(s)
    assert [int(v) for v in oracle.to_values(s)] == _h(t["degenerate"]["expected_values"])


def test_tip5_batches_match_oracle(tf, oracle):
    for count in (1, 2, 31, 128, 129, 5000):
        states = rnd(600 + count, 16 * count)
        want = states.copy()
        oracle.tip5_permute_batch(want)
        got = states.copy().reshape(count, 16)
        tf.Tip5.permutation(got)
        assert np.array_equal(got.reshape(-1), want)
        pairs = rnd(700 + count, 10 * count)
        want_h = oracle.hash_pairs_batch(pairs)
        got_h = tf.Tip5.hash_10(pairs.reshape(count, 10))
        assert np.array_equal(got_h.reshape(-1), want_h)
        got_p = tf.Tip5.hash_pair(np.ascontiguousarray(pairs.reshape(count, 10)[:, :5]),
                                  np.ascontiguousarray(pairs.reshape(count, 10)[:, 5:]))
        assert np.array_equal(got_p.reshape(-1), want_h)
    # adversarial lanes
    for fill in (0, P - 1, 0xFFFFFFFF00000000, 0xFFFFFFFF):
        st = np.full(16 * 3, fill, dtype=np.uint64)
        want = st.copy()
        oracle.tip5_permute_batch(want)
        got = st.copy()
        tf.Tip5.permutation(got)
        assert np.array_equal(got, want)


def test_tip5_hash_varlen_and_rows_match_oracle(tf, oracle):
    for length in (0, 1, 9, 10, 11, 19, 20, 21, 100, 16384):
        x = rnd(800 + length, length)
        assert np.array_equal(tf.Tip5.hash_varlen(x), oracle.hash_varlen(x)), length
    for row_len, n_rows in ((1, 7), (10, 130), (33, 257), (0, 3), (25, 4097), (7, 20000)):  # <= 4096 rows: 16 lanes per row
        rows = rnd(900 + row_len, row_len * n_rows).reshape(n_rows, row_len)
        got = tf.Tip5.hash_rows(rows)
        assert np.array_equal(got, oracle.hash_rows(rows)), (row_len, n_rows)  # every row


# ---- Merkle -------------------------------------------------------------------------------------------
@pytest.mark.parametrize("height", list(range(0, 15)) + [17, 20])
def test_merkle_tree_matches_oracle(tf, oracle, height):
    n = 1 << height
    leafs = rnd(0x5000 + height, 5 * n)
    rc, want = oracle.merkle_par_new(leafs)
    assert rc == 0
    tree = tf.MerkleTree.par_new(leafs.reshape(n, 5))
    assert np.array_equal(tree.nodes.reshape(-1), want)
    assert not tree.nodes[0].any()
    assert tree.num_leafs() == n and tree.height() == height
    assert np.array_equal(tree.leafs().reshape(-1), leafs)
    assert np.array_equal(tf.MerkleTree.par_frugal_root(leafs.reshape(n, 5)), want[5:10])
    assert np.array_equal(tf.MerkleTree.sequential_frugal_root(leafs.reshape(n, 5)), tree.root())


def test_merkle_errors(tf):
    # merkle_tree.rs:1025-1057
    e = np.zeros((0, 5), dtype=np.uint64)
    with pytest.raises(tf.MerkleTreeError) as ei:
        tf.MerkleTree.par_new(e)
    assert ei.value.kind == "TooFewLeafs"
    with pytest.raises(tf.MerkleTreeError) as ei:
        tf.MerkleTree.sequential_frugal_root(e)
    assert ei.value.kind == "TooFewLeafs"
    with pytest.raises(tf.MerkleTreeError) as ei:
        tf.MerkleTree.par_frugal_root(e)
    assert ei.value.kind == "IncorrectNumberOfLeafs"
    for n in (3, 5, 6, 7, 12, 1023):
        with pytest.raises(tf.MerkleTreeError) as ei:
            tf.MerkleTree.par_new(np.zeros((n, 5), dtype=np.uint64))
        assert ei.value.kind == "IncorrectNumberOfLeafs"
        with pytest.raises(tf.MerkleTreeError):
            tf.MerkleTree.par_frugal_root(np.zeros((n, 5), dtype=np.uint64))


def test_merkle_sharded_assembly_equals_single_tree(tf, oracle):
    """(e) of SURVEY.md: subtrees built independently + cap == one tree; scatter reproduces nodes."""
    import torch

    dev = importlib.import_module("twenty-first_b200.device")
    height, shards = 12, 4
    n = 1 << height
    leafs = rnd(0x7777, 5 * n)
    rc, want = oracle.merkle_par_new(leafs)
    cuda = torch.device("cuda:0")
    nl = n // shards
    global_nodes = torch.zeros(10 * n, dtype=torch.int64, device=cuda)
    roots = torch.zeros(5 * shards, dtype=torch.int64, device=cuda)
    for g in range(shards):
        shard = torch.from_numpy(leafs[5 * g * nl: 5 * (g + 1) * nl].view(np.int64)).to(cuda)
        local = torch.zeros(10 * nl, dtype=torch.int64, device=cuda)
        dev.merkle_build(shard, local)
        roots[5 * g: 5 * g + 5] = local[5:10]
        dev.merkle_scatter_subtree(local, g, shards, global_nodes)
    cap = torch.zeros(10 * shards, dtype=torch.int64, device=cuda)
    dev.merkle_build(roots, cap)
    global_nodes[: 5 * shards] = cap[: 5 * shards]
    torch.cuda.synchronize()
    got = global_nodes.cpu().numpy().view(np.uint64)
    assert np.array_equal(got, want)


def test_coset_lde_full_size_config3(tf, oracle):
    """BASELINE configs[3]: 2^22 -> 2^26 XFieldElement, offsets 7 -> 7, every output word against the oracle."""
    n_in, n_out = 1 << 22, 1 << 26
    values = rnd(0x210003, 3 * n_in)
    g = tf.BFieldElement.generator()
    rc, coeffs = oracle.coset_interpolate(values, 3, g)
    assert rc == 0
    rc, want = oracle.coset_evaluate(coeffs, 3, g, n_out)
    assert rc == 0
    got = tf.Polynomial.coset_lde(values.reshape(n_in, 3), g, n_out, g)
    assert np.array_equal(got.reshape(-1), want)


def test_tip5_hash_columns_matches_row_hashing(tf, oracle):
    """column-major codewords -> leaf digests: equals hash_varlen of the transposed rows"""
    import torch

    dev = importlib.import_module("twenty-first_b200.device")
    cuda = torch.device("cuda:0")
    for n_rows, n_cols in ((1, 1), (257, 7), (1024, 10), (4096, 33), (8192, 21), (1 << 16, 4)):
        cols = rnd(0xC0 + n_cols, n_rows * n_cols).reshape(n_cols, n_rows)
        d_cols = torch.from_numpy(cols.view(np.int64)).to(cuda)
        out = torch.zeros(5 * n_rows, dtype=torch.int64, device=cuda)
        dev.tip5_hash_columns(d_cols, n_rows, n_cols, out)
        torch.cuda.synchronize()
        got = out.cpu().numpy().view(np.uint64).reshape(n_rows, 5)
        rows = np.ascontiguousarray(cols.T)
        assert np.array_equal(got, oracle.hash_rows(rows)), (n_rows, n_cols)  # every row
        assert np.array_equal(got, tf.Tip5.hash_rows(rows))


def test_concurrent_callers_are_thread_safe(tf, oracle):
    """The Rust items are callable from any rayon worker (ntt.rs:71 OnceLock tables): hammer the C ABI
    from several host threads at once (ctypes releases the GIL) with different sizes."""
    import threading

    jobs = [(10, 1), (13, 1), (16, 1), (12, 3), (20, 1), (14, 3), (17, 1), (11, 1)]
    inputs = [rnd(0xAB00 + i, (1 << l) * w) for i, (l, w) in enumerate(jobs)]
    wants = []
    for (l, w), x in zip(jobs, inputs):
        y = x.copy()
        oracle.ntt(y, w)
        wants.append(y)
    errors = []

    def work(i):
        try:
            l, w = jobs[i]
            for _ in range(3):
                y = inputs[i].copy()
                arr = y.reshape(-1, 3) if w == 3 else y
                tf.ntt(arr)
                if not np.array_equal(y, wants[i]):
                    errors.append(f"job {i} mismatch")
                leafs = inputs[i][: 5 * 64].reshape(64, 5).copy()
                tf.MerkleTree.par_new(leafs)
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))

    threads = [threading.Thread(target=work, args=(i,)) for i in range(len(jobs))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors


@pytest.mark.parametrize("log2n", [27, 28])
def test_ntt_four_pass_plan_2_27(tf, oracle, log2n):
    """2^27 = 128 x 1024 x 1024 and 2^28 = 256 x 1024 x 1024 (one leading pass through shared memory, ntt_mid_col_kernel
    with split twiddle tables, + column pass + row pass): impulse response and round trip (size-independent
    properties; the oracle would need minutes at these sizes).  The host slice is one pageable array larger than a
    staging chunk."""
    n = 1 << log2n
    imp = np.zeros(n, dtype=np.uint64)
    imp[3] = 0xFFFFFFFF  # raw one at position 3 -> omega^(3 i)
    tf.ntt(imp)
    omega = oracle.primitive_root_of_unity(n)
    w3 = oracle.bfe_mod_pow(omega, 3)
    for i in (0, 1, 2, 1023, 1024, 123456789 % n, n // 2, n - 1):
        assert int(imp[i]) == oracle.bfe_mod_pow(w3, i), i
    assert (imp < np.uint64(P)).all()
    tf.intt(imp)
    assert int(imp[3]) == 0xFFFFFFFF and np.count_nonzero(imp) == 1
    x = rnd(0x2727, n)
    y = x.copy()
    tf.ntt(y)
    assert not np.array_equal(x, y)
    tf.intt(y)
    assert np.array_equal(x, y)


def test_config1_full_batch_256_x_2_20(tf, oracle):
    """BASELINE configs[1] at full size: 256 columns x 2^20 BFieldElement through the host-slice C ABI.
    Eleven columns (random + adversarial) are compared word for word with the oracle, then the inverse
    must restore all 2^28 input words."""
    api = importlib.import_module("twenty-first_b200.api")
    n, cols = 1 << 20, 256
    x = rnd(0x210001, n * cols)
    adv = adversarial(n)
    x[5 * n: 6 * n] = adv[1]      # all p-1
    x[6 * n: 7 * n] = adv[2]      # lanes 0xffffffff00000000
    x[7 * n: 8 * n] = adv[4]      # impulse
    orig = x.copy()
    api.ntt_batch(x, n, 1, False)
    assert (x < np.uint64(P)).all()
    for c in (0, 1, 5, 6, 7, 17, 100, 128, 200, 254, 255):
        want = orig[c * n: (c + 1) * n].copy()
        assert oracle.ntt(want, 1) == 0
        assert np.array_equal(x[c * n: (c + 1) * n], want), c
    api.ntt_batch(x, n, 1, True)
    assert np.array_equal(x, orig)



@pytest.mark.parametrize("log2n,width,batch", [(20, 1, 25), (21, 3, 3), (24, 1, 2), (12, 1, 5000)])
def test_pageable_host_slices_go_through_the_pinned_ring(tf, oracle, log2n, width, batch):
    """A plain numpy array is pageable memory, like the reference caller's Vec<BFieldElement> (ntt.rs:67,109): tf21_ntt /
    tf21_intt move it through the pinned staging ring of csrc/host_stage.cuh (whole chunks, a ragged last chunk, arrays
    larger than a chunk, many small arrays per chunk).  Bit-exact against the oracle, and against the driver-staged copy
    path (TF21_NO_STAGE_RING=1) the ring replaces."""
    api = importlib.import_module("twenty-first_b200.api")
    n = 1 << log2n
    x = rnd(0x7700 + log2n + batch, n * width * batch)
    want = x.copy()
    assert oracle.ntt_batch(want, n, width, batch, False) == 0
    got = x.copy()
    api.ntt_batch(got, n, width, False)
    assert np.array_equal(got, want)
    os.environ["TF21_NO_STAGE_RING"] = "1"
    try:
        direct = x.copy()
        api.ntt_batch(direct, n, width, False)
    finally:
        del os.environ["TF21_NO_STAGE_RING"]
    assert np.array_equal(direct, want)
    api.ntt_batch(got, n, width, True)
    assert np.array_equal(got, x)


def test_host_merkle_build_writes_the_leaf_half_on_the_host(tf, oracle):
    """tf21_merkle_build on pageable slices: leaves through the ring, inner nodes back through the ring, the leaf half of
    the node array copied on the host (merkle_tree.rs:426) -- every node against the oracle, with the driver-staged path
    next to it."""
    n = 1 << 21
    leafs = rnd(0x7800, 5 * n)
    rc, want = oracle.merkle_par_new(leafs)
    assert rc == 0
    tree = tf.MerkleTree.par_new(leafs.reshape(n, 5))
    assert np.array_equal(tree.nodes.reshape(-1), want)
    os.environ["TF21_NO_STAGE_RING"] = "1"
    try:
        tree2 = tf.MerkleTree.par_new(leafs.reshape(n, 5))
    finally:
        del os.environ["TF21_NO_STAGE_RING"]
    assert np.array_equal(tree2.nodes.reshape(-1), want)


def test_config2_merkle_2_24_full_node_array(tf, oracle):
    """BASELINE configs[2] at full size: every one of the 2^25 node digests against the oracle."""
    n = 1 << 24
    leafs = rnd(0x210002, 5 * n)
    rc, want = oracle.merkle_par_new(leafs)
    assert rc == 0
    tree = tf.MerkleTree.par_new(leafs.reshape(n, 5))
    assert np.array_equal(tree.nodes.reshape(-1), want)
    assert np.array_equal(tf.MerkleTree.par_frugal_root(leafs.reshape(n, 5)), want[5:10])


@pytest.mark.parametrize("na,nb,width", [(1, 1, 1), (1, 1, 3), (2, 2, 1), (5, 3, 3), (33, 31, 1), (64, 65, 3),
                                         (1000, 25, 1), (1 << 12, 1 << 12, 1), (3000, 5000, 3),
                                         (1 << 17, (1 << 17) + 1, 1), (1 << 19, 1 << 19, 3)])
def test_poly_fast_multiply_matches_oracle(tf, oracle, na, nb, width):
    """next wave: Polynomial::fast_multiply on the device, every coefficient vs the oracle"""
    a, b = rnd(0x900 + na, na * width), rnd(0x901 + nb, nb * width)
    if na * nb <= 1 << 16:
        want = oracle.poly_naive_multiply(a, b, width)
    else:
        rc, want = oracle.poly_fast_multiply(a, b, width)
        assert rc == 0
    pa = tf.Polynomial(a.reshape(na, 3) if width == 3 else a)
    pb = tf.Polynomial(b.reshape(nb, 3) if width == 3 else b)
    got = pa.fast_multiply(pb).coefficients
    assert np.array_equal(got.reshape(-1), want)
    assert (got < np.uint64(P)).all()


def test_poly_multiply_evaluation_property(tf, oracle):
    """(a * b)(x) == a(x) * b(x) at random points (size-independent check at 2^21-coefficient operands)"""
    na = nb = 1 << 21
    a, b = rnd(0xA1, na), rnd(0xB1, nb)
    prod = tf.Polynomial(a).fast_multiply(tf.Polynomial(b)).coefficients
    assert prod.shape[0] == na + nb - 1
    for seed in (1, 2, 3):
        x = int(rnd(0xE0 + seed, 1)[0])
        assert oracle.poly_evaluate(prod, x) == oracle.bfe_mul(oracle.poly_evaluate(a, x), oracle.poly_evaluate(b, x))
    # zero polynomial
    z = tf.Polynomial(np.zeros(0, dtype=np.uint64)).fast_multiply(tf.Polynomial(a[:4].copy()))
    assert z.coefficients.size == 0


# ---- next wave: authentication structures ------------------------------------------
@pytest.mark.parametrize("height,k", [(0, 1), (1, 1), (3, 2), (6, 5), (10, 1), (10, 40), (14, 160), (18, 80)])
def test_authentication_structure_matches_oracle(tf, oracle, height, k):
    """merkle_tree.rs:514-542, 614-622: device-built tree + device gather == oracle tree indexed on the host"""
    import torch

    n = 1 << height
    leafs = rnd(0x8000 + height, 5 * n)
    rng = np.random.default_rng(height * 100 + k)
    idx = rng.integers(0, n, size=k).astype(np.uint64)
    rc, nodes = oracle.merkle_par_new(leafs)
    assert rc == 0
    rc, want_idx = oracle.auth_structure_node_indices(n, idx)
    assert rc == 0
    want = nodes.reshape(-1, 5)[want_idx.astype(np.int64)]
    # host leaves in, host digests out
    got = tf.MerkleTree.par_authentication_structure_from_leafs(leafs.reshape(n, 5), idx)
    assert np.array_equal(got, want)
    # host-resident tree (plain gather through the mirror)
    tree = tf.MerkleTree.par_new(leafs.reshape(n, 5))
    assert np.array_equal(tree.authentication_structure(idx), want)
    # device-resident tree
    d_leafs = torch.from_numpy(leafs.view(np.int64)).cuda()
    d_nodes = torch.zeros(10 * n, dtype=torch.int64, device="cuda")
    tf.device.merkle_build(d_leafs, d_nodes)
    d_out = torch.zeros(5 * max(1, want.shape[0]), dtype=torch.int64, device="cuda")
    cnt = tf.device.merkle_authentication_structure(d_nodes, idx, d_out)
    assert cnt == want.shape[0]
    assert np.array_equal(d_out.cpu().numpy().view(np.uint64)[: 5 * cnt].reshape(-1, 5), want)


def test_authentication_structure_errors(tf):
    leafs = np.zeros((8, 5), dtype=np.uint64)
    with pytest.raises(tf.MerkleTreeError) as ei:
        tf.MerkleTree.par_authentication_structure_from_leafs(leafs, [8])
    assert ei.value.kind == "LeafIndexInvalid"
    with pytest.raises(tf.MerkleTreeError) as ei:
        tf.MerkleTree.par_authentication_structure_from_leafs(leafs[:6], [1])
    assert ei.value.kind == "IncorrectNumberOfLeafs"
    assert tf.MerkleTree.par_authentication_structure_from_leafs(leafs, []).shape == (0, 5)


# ---- next wave: MMR bulk operations ----------------------------------------------------
@pytest.mark.parametrize("n", [0, 1, 2, 3, 4, 5, 7, 8, 11, 100, 255, 256, 257, 1000, 4097, (1 << 16) + 3, (1 << 18) - 1])
def test_mmr_peaks_and_bag_peaks_match_oracle(tf, oracle, n):
    """mmr_accumulator.rs:29-34, 96-115, 379-391 for arbitrary (non power of two) leaf counts"""
    leafs = rnd(0x9000 + n, 5 * n)
    want_peaks = oracle.mmr_peaks_from_leafs(leafs)
    mmr = tf.MmrAccumulator.new_from_leafs(leafs.reshape(n, 5))
    assert mmr.num_leafs() == n and mmr.is_consistent()
    assert np.array_equal(mmr.peaks(), want_peaks)
    assert np.array_equal(mmr.bag_peaks(), oracle.mmr_bag_peaks(want_peaks, n))


def test_mmr_bag_peaks_reference_snapshot(tf, oracle):
    """mmr_accumulator.rs:1038-1047 (empty MMR) and `init` with ten peaks and a 10-bit leaf count (:1061-1065)"""
    empty = tf.MmrAccumulator.new_from_leafs(np.zeros((0, 5), dtype=np.uint64))
    assert tf.Digest.to_hex(empty.bag_peaks()) == (
        "cd65052100640f0d27e5654f97c47e49899add2f265967ccbefee7264e9bc08f588542d9dc3d5ac5")
    peaks = rnd(0x9999, 50).reshape(10, 5)
    for count in (0b11_1111_1111, (1 << 40) + 5, (1 << 64) - 1):
        got = tf.MmrAccumulator(peaks, count).bag_peaks()
        assert np.array_equal(got, oracle.mmr_bag_peaks(peaks, count))


@pytest.mark.parametrize("na,width", [(1, 1), (1, 3), (2, 1), (3, 3), (33, 1), (64, 3), (1000, 1), (1025, 3), (5000, 1)])
def test_poly_fast_square_matches_oracle(tf, oracle, na, width):
    """Polynomial::fast_square (polynomial.rs:780-802) == fast_multiply(self, self) of the oracle"""
    a = rnd(0xA000 + na, na * width)
    rc, want = oracle.poly_fast_multiply(a, a.copy(), width)
    assert rc == 0
    got = tf.Polynomial(a.reshape(na) if width == 1 else a.reshape(na, 3)).fast_square()
    assert np.array_equal(got.coefficients.reshape(-1), want)


# ---- next wave: coset extrapolation to out-of-domain points -------------------------
@pytest.mark.parametrize("log_n,n_cw,n_pts,width", [(0, 1, 1, 1), (1, 2, 3, 3), (5, 2, 2, 1), (8, 3, 5, 3),
                                                    (10, 7, 4, 1), (12, 2, 9, 3), (16, 3, 2, 1), (18, 2, 3, 3)])
def test_batch_coset_extrapolate_matches_oracle(tf, oracle, log_n, n_cw, n_pts, width):
    """polynomial.rs:2188-2331; the doc example (:2202-2213) is the first hard-coded case below"""
    n = 1 << log_n
    cws = rnd(0xB000 + log_n, n * width * n_cw)
    pts = rnd(0xB100 + log_n, n_pts * width)
    offset = tf.BFieldElement.generator()
    rc, want = oracle.batch_coset_extrapolate(offset, n, cws, width, pts)
    assert rc == 0
    got = tf.Polynomial.par_batch_coset_extrapolate(offset, n, cws if width == 1 else cws.reshape(-1, 3),
                                                    pts if width == 1 else pts.reshape(-1, 3))
    assert np.array_equal(got.reshape(-1), want)


def test_batch_coset_extrapolate_reference_doc_example(tf, oracle):
    n = 1 << 5
    cws = np.concatenate([oracle.to_raw([3] * n), oracle.to_raw([2] * n)])
    pts = oracle.to_raw([0, 1])
    got = tf.Polynomial.batch_coset_extrapolate(tf.BFieldElement.new(7), n, cws, pts)
    assert oracle.to_values(got).tolist() == [3, 3, 2, 2]
    with pytest.raises(tf.Tf21Error):
        tf.Polynomial.batch_coset_extrapolate(tf.BFieldElement.new(7), 12, cws[:24], pts)


def test_extrapolation_in_the_domain_reproduces_the_codeword(tf, oracle):
    """size-independent property: extrapolating to points of the coset itself returns the codeword values"""
    log_n, width = 14, 3
    n = 1 << log_n
    cw = rnd(0xB200, n * width)
    offset = tf.BFieldElement.generator()
    omega = oracle.primitive_root_of_unity(n)
    idx = [0, 1, 2, 77, n // 2, n - 1]
    pts = np.zeros((len(idx), 3), dtype=np.uint64)
    for k, i in enumerate(idx):
        pts[k, 0] = oracle.bfe_mul(offset, oracle.bfe_mod_pow(omega, i))
    got = tf.Polynomial.coset_extrapolate(offset, cw.reshape(n, 3), pts)
    assert np.array_equal(got, cw.reshape(n, 3)[idx])


def test_tip5_sample_indices_matches_oracle(tf, oracle):
    """tip5/mod.rs:636-656, incl. a state that squeezes BFieldElement::MAX (dropped by the rejection sampling)"""
    for seed, ub, count in [(1, 1 << 10, 1), (2, 1 << 20, 45), (3, 1, 12), (4, 1 << 31, 333), (5, 1 << 16, 10)]:
        state = rnd(0xC000 + seed, 16)
        if seed == 5:
            state[3] = np.uint64(oracle.bfe_new(0xFFFFFFFF00000000))  # MAX in the first squeeze
        want_state = state.copy()
        want = oracle.tip5_sample_indices(want_state, ub, count)
        got_state = state.copy()
        got = tf.Tip5.sample_indices(got_state, ub, count)
        assert np.array_equal(got, want)
        assert np.array_equal(got_state, want_state)
    with pytest.raises(tf.Tf21Error):
        tf.Tip5.sample_indices(rnd(1, 16), 12, 3)


def test_ntt_batch_beyond_grid_y_limit(tf, oracle):
    """more than 65535 arrays in one call: the row pass leaves its 2-D grid (arrays on blockIdx.y) for the
    flat indexing; arrays are independent, so a sample of them is compared with the oracle"""
    import torch

    dev = importlib.import_module("twenty-first_b200.device")
    log2n, batch = 12, 70000
    n = 1 << log2n
    x = torch.randint(0, 2**62, (batch * n,), dtype=torch.int64, device="cuda:0")
    picks = [0, 1, 65534, 65535, 65536, batch - 1]
    before = {b: x[b * n:(b + 1) * n].cpu().numpy().view(np.uint64).copy() for b in picks}
    dev.ntt_(x, n, 1, False)
    torch.cuda.synchronize()
    for b in picks:
        want = before[b].copy()
        assert oracle.ntt(want, 1) == 0
        assert np.array_equal(x[b * n:(b + 1) * n].cpu().numpy().view(np.uint64), want), b
    dev.ntt_(x, n, 1, True)
    torch.cuda.synchronize()
    for b in picks:
        assert np.array_equal(x[b * n:(b + 1) * n].cpu().numpy().view(np.uint64), before[b]), b


@pytest.mark.parametrize("nq,nb", [(1, 1), (1, 5), (7, 1), (33, 31), (600, 700), (1, 2000), (3000, 1100), (5000, 4000)])
def test_poly_clean_divide_matches_long_division(tf, oracle, nq, nb):
    """polynomial.rs:2358-2413: (q * b) / b == q, checked against the oracle's naive_divide (remainder zero)"""
    q = rnd(0xD000 + nq, nq)
    b = rnd(0xD100 + nb, nb)
    q[-1] |= np.uint64(1)
    b[-1] |= np.uint64(1)
    a = oracle.poly_naive_multiply(q, b, 1)
    want_q, rem = oracle.poly_naive_divide(a, b)
    assert not rem.any() and np.array_equal(want_q, q)
    got = tf.Polynomial(a).clean_divide(tf.Polynomial(b))
    assert np.array_equal(got.coefficients, q)


def test_poly_clean_divide_edge_cases(tf, oracle):
    b = rnd(0xD200, 9)
    # a root of the divisor on the first coset (offset 7): b = (x - 7) * c  ->  the offset is changed, result unchanged
    c = rnd(0xD201, 20)
    lin = np.array([oracle.bfe_new(P - 7), oracle.bfe_new(1)], dtype=np.uint64)
    bb = oracle.poly_naive_multiply(lin, c, 1)
    q = rnd(0xD202, 50)
    a = oracle.poly_naive_multiply(q, bb, 1)
    assert np.array_equal(tf.Polynomial(a).clean_divide(tf.Polynomial(bb)).coefficients, q)
    # trailing zero coefficients are ignored; divisor with zero constant term (polynomial.rs:2372-2379)
    xb = np.concatenate([np.zeros(1, dtype=np.uint64), b, np.zeros(3, dtype=np.uint64)])
    a2 = oracle.poly_naive_multiply(q, xb[:10], 1)
    assert np.array_equal(tf.Polynomial(np.concatenate([a2, np.zeros(5, dtype=np.uint64)])).clean_divide(
        tf.Polynomial(xb)).coefficients, q)
    # zero dividend, zero divisor
    assert tf.Polynomial(np.zeros(4, dtype=np.uint64)).clean_divide(tf.Polynomial(b)).coefficients.size == 0
    with pytest.raises(tf.Tf21Error):
        tf.Polynomial(a).clean_divide(tf.Polynomial(np.zeros(3, dtype=np.uint64)))


@pytest.mark.parametrize("n,log_dl,tail,width", [(10, 4, 5, 1), (16, 4, 5, 1), (17, 4, 5, 1), (100, 4, 3, 3), (1000, 6, 20, 1),
                                                 (5000, 8, 100, 3), (70000, 12, 1500, 1), (9, 3, 0, 1)])
def test_poly_reduce_by_ntt_friendly_modulus_matches_oracle(tf, oracle, n, log_dl, tail, width):
    """polynomial.rs:1087-1148, and its defining property: the result is f mod (X^dl + shift)"""
    dl = 1 << log_dl
    f = rnd(0xE000 + n, n * width)
    shift = np.zeros(dl * width, dtype=np.uint64)
    shift[: tail * width] = rnd(0xE100 + n, tail * width)   # deg shift < tail_length
    shift_ntt = shift.copy()
    assert oracle.ntt(shift_ntt, width) == 0
    k, want = oracle.poly_reduce_by_ntt_friendly_modulus(f, width, shift_ntt, tail)
    assert k == min(n, dl)
    shp = (lambda a: a if width == 1 else a.reshape(-1, 3))
    got = tf.Polynomial(shp(f)).reduce_by_ntt_friendly_modulus(shp(shift_ntt), tail)
    assert np.array_equal(got.coefficients.reshape(-1), want)
    if width == 1 and n >= dl:
        # property: f - result is a multiple of the modulus M = X^dl + shift  ->  long division leaves remainder 0
        modulus = np.concatenate([shift[:dl], np.array([oracle.bfe_new(1)], dtype=np.uint64)])
        diff = f.copy()
        for i in range(dl):
            diff[i] = oracle.bfe_sub(int(diff[i]), int(want[i]))
        _, rem = oracle.poly_naive_divide(diff, modulus)
        assert not rem.any()


@pytest.mark.parametrize("n_shards", [1, 2, 4, 8])
def test_single_process_sharded_entry_points(tf, oracle, n_shards):
    """columns / subtrees split over host threads (shard s on device s % visible devices, so the
    index algebra is exercised on one GPU as well); results equal the unsharded ones"""
    log2n, batch = 12, 11
    x = rnd(0xF000 + n_shards, batch << log2n)
    want = x.copy()
    assert oracle.ntt_batch(want, 1 << log2n, 1, batch, False) == 0
    got = x.copy()
    tf.check(tf.lib.tf21_ntt_sharded(got.ctypes.data, 1 << log2n, 1, batch, 0, n_shards))
    assert np.array_equal(got, want)
    tf.check(tf.lib.tf21_ntt_sharded(got.ctypes.data, 1 << log2n, 1, batch, 1, n_shards))
    assert np.array_equal(got, x)
    for height in (0, 2, 3, 11):
        n = 1 << height
        leafs = rnd(0xF100 + height, 5 * n)
        rc, want_nodes = oracle.merkle_par_new(leafs)
        assert rc == 0
        nodes = np.zeros(10 * n, dtype=np.uint64)
        tf.check(tf.lib.tf21_merkle_build_sharded(leafs.ctypes.data, n, nodes.ctypes.data, n_shards))
        assert np.array_equal(nodes, want_nodes), (height, n_shards)
    assert tf.lib.tf21_merkle_build_sharded(leafs.ctypes.data, n, nodes.ctypes.data, 3) == tf.E_BAD_ARG


def test_device_entry_points_on_a_non_default_stream(tf, oracle):
    """the *_dev entry points are stream ordered: run them on a side stream (with work queued on the default
    stream meanwhile) and compare with the oracle"""
    import torch

    dev = importlib.import_module("twenty-first_b200.device")
    side = torch.cuda.Stream()
    n, batch = 1 << 20, 3
    x = rnd(0x5EED, n * batch)
    want = x.copy()
    assert oracle.ntt_batch(want, n, 1, batch, False) == 0
    leafs = rnd(0x5EEE, 5 << 12)
    rc, want_nodes = oracle.merkle_par_new(leafs)
    assert rc == 0
    d_x = torch.from_numpy(x.view(np.int64)).cuda()
    d_leafs = torch.from_numpy(leafs.view(np.int64)).cuda()
    d_nodes = torch.zeros(10 << 12, dtype=torch.int64, device="cuda")
    noise = torch.randint(0, 2**62, (64 << 20,), dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    with torch.cuda.stream(side):
        dev.ntt_(d_x, n, 1, False)
        dev.merkle_build(d_leafs, d_nodes)
    dev.ntt_(noise, n, 1, False)  # concurrent work on the default stream
    side.synchronize()
    torch.cuda.synchronize()
    assert np.array_equal(d_x.cpu().numpy().view(np.uint64), want)
    assert np.array_equal(d_nodes.cpu().numpy().view(np.uint64), want_nodes)


def test_shutdown_then_reuse(tf, oracle):
    """tf21_shutdown frees every cached table; the next call must rebuild them (a stale pointer to a freed
    twiddle table would give silently wrong transforms at n >= 2^10)"""
    for log2n in (10, 16, 20):
        x = rnd(0xAD00 + log2n, 1 << log2n)
        y = x.copy()
        tf.ntt(y)  # populate the caches
    tf.check(tf.lib.tf21_shutdown())
    for log2n in (10, 16, 20, 22):
        x = rnd(0xAD10 + log2n, 1 << log2n)
        want = x.copy()
        assert oracle.ntt(want, 1) == 0
        got = x.copy()
        tf.ntt(got)
        assert np.array_equal(got, want), log2n
        tf.intt(got)
        assert np.array_equal(got, x), log2n
    leafs = rnd(0xAD20, 5 << 10)
    rc, want_nodes = oracle.merkle_par_new(leafs)
    assert np.array_equal(tf.MerkleTree.par_new(leafs.reshape(-1, 5)).nodes.reshape(-1), want_nodes)
    tf.check(tf.lib.tf21_shutdown())
    tf.check(tf.lib.tf21_init(0))


@pytest.mark.parametrize("log2n", [10, 11, 13, 16, 20, 21])
def test_ntt_on_an_8_byte_aligned_view(tf, oracle, log2n):
    """the C ABI promises only u64 alignment: a device view at an odd word offset (8-byte aligned) must not take
    a 16-byte store path"""
    import torch

    dev = importlib.import_module("twenty-first_b200.device")
    n, batch = 1 << log2n, 3
    x = rnd(0xA110 + log2n, n * batch)
    want = x.copy()
    assert oracle.ntt_batch(want, n, 1, batch, False) == 0
    buf = torch.zeros(n * batch + 1, dtype=torch.int64, device="cuda")
    view = buf[1:]
    assert view.data_ptr() % 16 == 8
    view.copy_(torch.from_numpy(x.view(np.int64)))
    dev.ntt_(view, n, 1, False)
    torch.cuda.synchronize()
    assert np.array_equal(view.cpu().numpy().view(np.uint64), want)
    dev.ntt_(view, n, 1, True)
    torch.cuda.synchronize()
    assert np.array_equal(view.cpu().numpy().view(np.uint64), x)


def test_tip5_dev_entry_points_refuse_8_byte_aligned_inputs(tf):
    import torch

    buf = torch.zeros(1 + 16 * 8, dtype=torch.int64, device="cuda")
    out = torch.zeros(5 * 8, dtype=torch.int64, device="cuda")
    odd = buf[1:].data_ptr()
    assert tf.lib.tf21_tip5_permute_dev(odd, 8, None) == tf.E_BAD_ARG
    assert tf.lib.tf21_tip5_hash_10_dev(odd, 8, out.data_ptr(), None) == tf.E_BAD_ARG
    assert tf.lib.tf21_merkle_build_dev(odd, 8, buf.data_ptr(), None) == tf.E_BAD_ARG
    assert tf.lib.tf21_merkle_root_dev(odd, 8, out.data_ptr(), None) == tf.E_BAD_ARG


def test_init_semantics_and_device_count(tf):
    """tf21_init(n_devices): 0 = all visible; more than visible is refused"""
    import torch

    visible = torch.cuda.device_count()
    tf.check(tf.lib.tf21_init(0))
    assert tf.lib.tf21_device_count() == visible
    tf.check(tf.lib.tf21_init(1))
    assert tf.lib.tf21_device_count() == 1
    assert tf.lib.tf21_sharded_uses_nccl(2) == 0  # one device in use: two shards share it, host-memory cap
    assert tf.lib.tf21_init(visible + 1) == tf.E_BAD_ARG
    assert tf.lib.tf21_init(-1) == tf.E_BAD_ARG
    tf.check(tf.lib.tf21_init(0))
    tf.check(tf.lib.tf21_set_device(0))


def test_sharded_merkle_cap_over_nccl(tf, oracle):
    """one shard per device: the tree cap crosses devices with one ncclAllGather inside the library;
    the full node array equals the oracle's.  Needs >= 2 visible GPUs (gpurun --gpus 2)."""
    import torch

    n_dev = torch.cuda.device_count()
    if n_dev < 2:
        pytest.skip("needs at least two GPUs")
    tf.check(tf.lib.tf21_init(0))
    for n_shards in [s for s in (2, 4, 8) if s <= n_dev]:
        assert tf.lib.tf21_sharded_uses_nccl(n_shards) == 1
        for height in (3, 10, 16):
            n = 1 << height
            leafs = rnd(0xCC00 + height + n_shards, 5 * n)
            rc, want_nodes = oracle.merkle_par_new(leafs)
            assert rc == 0
            nodes = np.zeros(10 * n, dtype=np.uint64)
            tf.check(tf.lib.tf21_merkle_build_sharded(leafs.ctypes.data, n, nodes.ctypes.data, n_shards))
            assert np.array_equal(nodes, want_nodes), (height, n_shards)
        # columns over the devices, no exchange
        log2n, batch = 14, 4 * n_shards + 1
        x = rnd(0xCC50 + n_shards, batch << log2n)
        want = x.copy()
        assert oracle.ntt_batch(want, 1 << log2n, 1, batch, False) == 0
        tf.check(tf.lib.tf21_ntt_sharded(x.ctypes.data, 1 << log2n, 1, batch, 0, n_shards))
        assert np.array_equal(x, want)


def test_tma_tile_layout_matches_the_swizzle_formula(tf):
    """the column pass reads and writes its [1024 rows][4 words] tile in the layout TMA produces with
    CU_TENSOR_MAP_SWIZZLE_32B: word (r, c) at 4 r + 2 ((c >> 1) ^ ((r >> 2) & 1)) + (c & 1) (csrc/tma.cuh)"""
    import torch

    inner, n_tiles = 16, 4
    m = torch.arange(1024 * inner, dtype=torch.int64, device="cuda")
    out = torch.zeros(n_tiles * 4096, dtype=torch.int64, device="cuda")
    tf.check(tf.lib.tf21_selftest_tma_tile_dev(m.data_ptr(), inner, n_tiles, out.data_ptr(), None))
    torch.cuda.synchronize()
    got = out.cpu().numpy().reshape(n_tiles, 4096)
    r, c = np.meshgrid(np.arange(1024), np.arange(4), indexing="ij")
    idx = 4 * r + 2 * ((c >> 1) ^ ((r >> 2) & 1)) + (c & 1)
    for t in range(n_tiles):
        want = np.zeros(4096, dtype=np.int64)
        want[idx.reshape(-1)] = (r * inner + 4 * t + c).reshape(-1)
        assert np.array_equal(got[t], want), t


@pytest.mark.parametrize("width", [1, 3])
@pytest.mark.parametrize("log2n", list(range(1, 10)))
def test_small_n_register_kernel(tf, oracle, log2n, width):
    """n < 2^10 (reference benches at 2^7, benches/ntt.rs:19): a warp takes 1024 consecutive elements = whole
    columns; batches that do not fill the last 1024-element chunk, forward and inverse, on the device path"""
    import torch

    dev = importlib.import_module("twenty-first_b200.device")
    n = 1 << log2n
    for batch in (1, 3, (1024 >> log2n), (1024 >> log2n) + 1, 777):
        x = rnd(0x5A00 + 16 * log2n + width + batch, n * width * batch)
        x[: min(x.size, 4)] = [P - 1, 0, 0xFFFFFFFF00000000, 1][: min(x.size, 4)]
        want = x.copy()
        assert oracle.ntt_batch(want, n, width, batch, False) == 0
        d = torch.from_numpy(x.view(np.int64)).cuda()
        dev.ntt_(d, n, width, False)
        torch.cuda.synchronize()
        assert np.array_equal(d.cpu().numpy().view(np.uint64), want), (log2n, width, batch)
        dev.ntt_(d, n, width, True)
        torch.cuda.synchronize()
        assert np.array_equal(d.cpu().numpy().view(np.uint64), x), (log2n, width, batch)


def test_register_leading_pass_for_every_size_in_a_subprocess(tf):
    """ntt_col_n_kernel is only selected where it measured faster; TF21_COL_N_ALL=1 forces it for every leading pass
    of 2 .. 512 points so that all nine instantiations stay covered (the switch is read once per process)"""
    import subprocess
    import sys

    code = r'''
import importlib, sys, numpy as np
sys.path.insert(0, %r)
import oracle
tf = importlib.import_module("twenty-first_b200")
o = oracle.get()
for log2n, w in [(k, 1) for k in range(11, 20)] + [(13, 3), (16, 3), (19, 3), (21, 1), (23, 1), (25, 1)]:
    x = oracle.splitmix64_words(0xC01 + log2n, (1 << log2n) * w) %% np.uint64(oracle.P)
    want = x.copy(); assert o.ntt(want, w) == 0
    got = x.copy(); tf.ntt(got.reshape(-1, w) if w == 3 else got)
    assert np.array_equal(got, want), (log2n, w)
    tf.intt(got.reshape(-1, w) if w == 3 else got)
    assert np.array_equal(got, x), (log2n, w)
print("ok")
''' % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, TF21_COL_N_ALL="1")
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "ok" in r.stdout, r.stdout + r.stderr
