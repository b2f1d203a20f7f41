"""Pins the CPU oracle against every known-answer test the reference holds for the hot path
(SURVEY.md 8c).  CPU only."""
import numpy as np
import pytest

P = 0xFFFFFFFF00000001


def _h(xs):
    return [int(x, 16) if isinstance(x, str) else int(x) for x in xs]


# ---- field: b_field_element.rs:1478-1514, 1366-1386 -------------------------------------------
def test_field_fixed_vectors(oracle, kats):
    f = kats["field"]
    a = oracle.bfe_new(f["fixed_inverse"]["a"])
    assert oracle.bfe_value(oracle.bfe_inverse_or_zero(a)) == f["fixed_inverse"]["inv"]
    mp = f["fixed_modpow"]
    assert oracle.bfe_value(oracle.bfe_mod_pow(oracle.bfe_new(mp["base"]), mp["exp"])) == mp["expected"]
    for m in f["fixed_mul"]:
        c = oracle.bfe_mul(oracle.bfe_new(m["a"]), oracle.bfe_new(m["b"]))
        assert oracle.bfe_value(c) == m["c"]
    for base, e in f["roots_pow_is_one"]["cases"]:
        assert oracle.bfe_value(oracle.bfe_mod_pow(oracle.bfe_new(base), e)) == 1


def test_montgomery_word_is_value_times_2_64(oracle):
    # F1 of SURVEY.md: raw = v * 2^64 mod p, canonical
    rng = np.random.default_rng(1)
    for v in [0, 1, 2, P - 1, P - 2, 0xFFFFFFFF, 1 << 32, 1 << 63] + [int(x) for x in rng.integers(0, P, 50, dtype=np.uint64)]:
        raw = oracle.bfe_new(v)
        assert raw == (v << 64) % P
        assert oracle.bfe_value(raw) == v % P


def test_field_ops_match_python_bigint(oracle):
    rng = np.random.default_rng(2)
    xs = [0, 1, P - 1, P - 2, 0xFFFFFFFF, 0xFFFFFFFF00000000] + [int(x) for x in rng.integers(0, P, 100, dtype=np.uint64)]
    for a in xs[:20]:
        for b in xs:
            ra, rb = oracle.bfe_new(a), oracle.bfe_new(b)
            assert oracle.bfe_value(oracle.bfe_add(ra, rb)) == (a + b) % P
            assert oracle.bfe_value(oracle.bfe_sub(ra, rb)) == (a - b) % P
            assert oracle.bfe_value(oracle.bfe_mul(ra, rb)) == (a * b) % P


def test_primitive_roots(oracle):
    # get_primitive_root_of_unity_test, b_field_element.rs:1373-1386
    for i in range(1, 33):
        n = 1 << i
        root = oracle.primitive_root_of_unity(n)
        assert oracle.bfe_value(oracle.bfe_mod_pow(root, n)) == 1
        assert oracle.bfe_value(oracle.bfe_mod_pow(root, n // 2)) != 1
    # omega_n = 7^((p-1)/n) and omega_64 = 2^39
    for i in range(0, 33):
        n = 1 << i
        assert oracle.bfe_value(oracle.primitive_root_of_unity(n)) == pow(7, (P - 1) // n, P)
    assert oracle.bfe_value(oracle.primitive_root_of_unity(64)) == 1 << 39


# ---- ntt: ntt.rs:397-469, 511-560 -------------------------------------------------------------
def test_ntt_kats(oracle, kats):
    for name in ("bfield_basic", "bfield_max", "bfield_len32"):
        k = kats["ntt"][name]
        x = oracle.to_raw(k["input_values"])
        orig = x.copy()
        assert oracle.ntt(x, 1) == 0
        assert list(map(int, oracle.to_values(x))) == k["expected_values"], name
        assert oracle.intt(x, 1) == 0
        assert np.array_equal(x, orig), name
    k = kats["ntt"]["xfield_basic"]
    x = oracle.to_raw(np.array(k["input_values"], dtype=np.uint64).reshape(-1))
    orig = x.copy()
    assert oracle.ntt(x, 3) == 0
    assert oracle.to_values(x).reshape(-1, 3).tolist() == k["expected_values"]
    assert oracle.intt(x, 3) == 0
    assert np.array_equal(x, orig)


def test_ntt_edge_lengths_and_errors(oracle):
    # ntt_on_empty_input / length one / 0-1-0 ordering, ntt.rs:471-498 ; panics :135-137
    import oracle as o

    e = np.zeros(0, dtype=np.uint64)
    assert oracle.ntt(e, 1) == 0 and oracle.intt(e, 1) == 0
    one = oracle.to_raw([12345])
    c = one.copy()
    assert oracle.ntt(one, 1) == 0 and np.array_equal(one, c)
    assert oracle.ntt(e, 1) == 0
    bad = np.zeros(3, dtype=np.uint64)
    assert oracle.ntt(bad, 1) == o.E_LEN_NOT_POW2
    assert oracle.intt(np.zeros(12, dtype=np.uint64), 1) == o.E_LEN_NOT_POW2


def test_ntt_roundtrip_and_is_evaluation(oracle):
    # chu_ntt_b_field_prop_test ntt.rs:345-365 ; test_compare_ntt_to_eval :562-579
    import oracle as o

    for log2n in range(1, 10):
        n = 1 << log2n
        x = o.splitmix64_words(100 + log2n, n)
        x[0] = oracle.bfe_new(P - 1)  # BFieldElement::MAX
        orig = x.copy()
        oracle.ntt(x, 1)
        assert not np.array_equal(x, orig)
        omega = oracle.primitive_root_of_unity(n)
        for i in (0, 1, n // 2, n - 1):
            pt = oracle.bfe_mod_pow(omega, i)
            assert int(x[i]) == oracle.poly_evaluate(orig, pt)
        oracle.intt(x, 1)
        assert np.array_equal(x, orig)


def test_xfe_ntt_is_three_interleaved_bfe_ntts(oracle):
    # SURVEY F3; x_field_element.rs:540-548, 620-625
    import oracle as o

    for log2n in (1, 4, 9, 11):
        n = 1 << log2n
        x = o.splitmix64_words(7 + log2n, 3 * n)
        y = x.copy()
        oracle.ntt(y, 3)
        for c in range(3):
            col = np.ascontiguousarray(x[c::3])
            oracle.ntt(col, 1)
            assert np.array_equal(col, y[c::3])
        oracle.intt(y, 3)
        assert np.array_equal(x, y)


def test_plain_dft_on_raw_words_equals_reference_ntt(oracle):
    # the GPU path relies on NTT(R*a) = R*NTT(a): a plain mod-p DFT of the raw words
    # with canonical twiddles gives the raw words of the reference result.
    import oracle as o

    n = 64
    x = o.splitmix64_words(99, n)
    y = x.copy()
    oracle.ntt(y, 1)
    w = oracle.bfe_value(oracle.primitive_root_of_unity(n))
    xi = [int(v) for v in x]
    for i in range(n):
        acc = sum(xi[j] * pow(w, i * j, P) for j in range(n)) % P
        assert acc == int(y[i])
    z = y.copy()
    oracle.intt(z, 1)
    winv = pow(w, P - 2, P)
    ninv = pow(n, P - 2, P)
    yi = [int(v) for v in y]
    for i in range(n):
        acc = sum(yi[j] * pow(winv, i * j, P) for j in range(n)) * ninv % P
        assert acc == int(z[i]) == xi[i]


# ---- tip5: tip5/mod.rs:1034-1053, 1145-1206, 1294-1362, 1525-1531 ---------------------------------
def test_tip5_hash10_snapshot(oracle, kats):
    pre = np.zeros(10, dtype=np.uint64)
    for i in range(6):
        d = oracle.hash_10(pre)
        pre[i:i + 5] = d
    assert oracle.digest_to_hex(oracle.hash_10(pre)) == kats["tip5"]["hash10_snapshot"]["hex"]


def test_tip5_hash_varlen_snapshot(oracle, kats):
    acc = np.zeros(5, dtype=np.uint64)
    for i in range(20):
        d = oracle.hash_varlen(oracle.to_raw(list(range(i))) if i else np.zeros(0, dtype=np.uint64))
        acc = np.array([oracle.bfe_add(int(a), int(b)) for a, b in zip(acc, d)], dtype=np.uint64)
    assert oracle.digest_to_hex(acc) == kats["tip5"]["hash_varlen_sum"]["hex"]


def test_tip5_raw_state_snapshot(oracle, kats):
    k = kats["tip5"]["raw_snapshot"]
    s = np.array(_h(k["state_raw"]), dtype=np.uint64)
    oracle.tip5_permutation(s)
    assert [int(v) for v in s[:5]] == _h(k["expected_first5_raw"])


def test_tip5_degenerate_representation_kat(oracle, kats):
    k = kats["tip5"]["degenerate"]
    s = oracle.to_raw(_h(k["state_values"]))
    oracle.tip5_permutation(s)
    assert [int(v) for v in oracle.to_values(s)] == _h(k["expected_values"])


def test_tip5_hasher_snapshot(oracle, kats):
    k = kats["tip5"]["hasher_hello_world"]
    assert oracle.hasher_bytes(k["bytes"].encode()) == k["finish"]


def test_hash_pair_is_hash_10_of_concatenation(oracle):
    import oracle as o

    w = o.splitmix64_words(5, 10)
    assert np.array_equal(oracle.hash_pair(w[:5].copy(), w[5:].copy()), oracle.hash_10(w))


def test_tip5_plain_modp_model(oracle, kats):
    """the permutation equals plain mod-p arithmetic on raw words
    (raw^7, circulant MDS, RC*2^64).  This is the model the CUDA kernel implements."""
    import oracle as o

    lut = [((b + 1) ** 3 + 256) % 257 % 256 for b in range(256)]
    assert lut[:16] == kats["tip5"]["lookup_table_first16"]["values"]
    mds = [61402, 1108, 28750, 33823, 7454, 43244, 53865, 12034, 56951, 27521, 41351, 40901, 12021, 59689, 26798, 17845]
    # round constants: take them from the oracle's Montgomery conversion of the canonical table
    import re, os

    src = open(os.path.join(os.path.dirname(o.__file__), "oracle.c")).read()
    block = src.split("ROUND_CONSTANTS[NUM_ROUNDS * STATE_SIZE] = {")[1].split("};")[0]
    rc = [int(x) for x in re.findall(r"(\d+)ULL", block)]
    assert len(rc) == 80
    rc_raw = [(c << 64) % P for c in rc]
    assert rc_raw[:2] == _h(kats["tip5"]["rc_raw_first2"]["raw"])

    def perm(s):
        s = list(s)
        for r in range(5):
            for i in range(4):
                s[i] = int.from_bytes(bytes(lut[b] for b in s[i].to_bytes(8, "little")), "little")
            for i in range(4, 16):
                s[i] = pow(s[i], 7, P)
            t = [sum(mds[(i - j) % 16] * s[j] for j in range(16)) % P for i in range(16)]
            s = [(t[i] + rc_raw[16 * r + i]) % P for i in range(16)]
        return s

    for seed in range(5):
        st = o.splitmix64_words(1000 + seed, 16)
        exp = st.copy()
        oracle.tip5_permutation(exp)
        assert perm([int(v) for v in st]) == [int(v) for v in exp]
    k = kats["tip5"]["raw_snapshot"]
    assert perm(_h(k["state_raw"]))[:5] == _h(k["expected_first5_raw"])


# ---- merkle: merkle_tree.rs:1025-1128 ----------------------------------------------------------
def test_merkle_variants_agree_and_errors(oracle):
    import oracle as o

    assert oracle.merkle_sequential_new(np.zeros(0, dtype=np.uint64))[0] == o.E_TOO_FEW_LEAFS
    assert oracle.merkle_par_new(np.zeros(15, dtype=np.uint64))[0] == o.E_INCORRECT_NUMBER_OF_LEAFS
    assert oracle.merkle_sequential_frugal_root(np.zeros(0, dtype=np.uint64))[0] == o.E_TOO_FEW_LEAFS
    assert oracle.merkle_sequential_frugal_root(np.zeros(15, dtype=np.uint64))[0] == o.E_INCORRECT_NUMBER_OF_LEAFS
    for h in range(0, 12):
        n = 1 << h
        leafs = o.splitmix64_words(0x5000 + h, 5 * n)
        rc, seq = oracle.merkle_sequential_new(leafs)
        assert rc == 0
        assert not seq[:5].any()  # nodes[0] == ALL_ZERO
        assert np.array_equal(seq[5 * n:], leafs)
        for cutoff in (2, 4, 16, 512):
            for threads in (1, 3, 8):
                rc, par = oracle.merkle_par_new(leafs, threads, cutoff)
                assert rc == 0 and np.array_equal(par, seq)
                rc, r = oracle.merkle_par_frugal_root(leafs, threads, cutoff)
                assert rc == 0 and np.array_equal(r, seq[5:10])
        rc, r = oracle.merkle_sequential_frugal_root(leafs)
        assert rc == 0 and np.array_equal(r, seq[5:10])
        if n >= 2:
            assert np.array_equal(seq[5:10], oracle.hash_pair(seq[10:15].copy(), seq[15:20].copy()))


# ---- coset: polynomial.rs:3645-3679 ------------------------------------------------------------
def test_coset_evaluate_is_horner_on_coset(oracle):
    import oracle as o

    for log2 in range(0, 8):
        order = 1 << log2
        for ncoef in {1, max(1, order // 2), order}:
            coeffs = o.splitmix64_words(31 * log2 + ncoef, ncoef)
            offset = int(o.splitmix64_words(77 + log2, 1)[0])
            rc, vals = oracle.coset_evaluate(coeffs, 1, offset, order)
            assert rc == 0
            omega = oracle.primitive_root_of_unity(order)
            for i in range(order):
                pt = oracle.bfe_mul(offset, oracle.bfe_mod_pow(omega, i))
                assert int(vals[i]) == oracle.poly_evaluate(coeffs, pt)
            if ncoef == order:
                rc, back = oracle.coset_interpolate(vals, 1, offset)
                assert rc == 0 and np.array_equal(back, coeffs)
    # order <= degree panics in the reference
    coeffs = o.splitmix64_words(3, 8)
    assert oracle.coset_evaluate(coeffs, 1, oracle.bfe_new(7), 4)[0] == o.E_ORDER_LE_DEGREE
    # trailing zero coefficients do not count towards the degree
    coeffs[4:] = 0
    assert oracle.coset_evaluate(coeffs, 1, oracle.bfe_new(7), 4)[0] == 0


def test_coset_xfe_lanes_independent(oracle):
    import oracle as o

    n, order = 16, 64
    coeffs = o.splitmix64_words(11, 3 * n)
    g = oracle.bfe_new(7)
    rc, vals = oracle.coset_evaluate(coeffs, 3, g, order)
    assert rc == 0
    for c in range(3):
        rc, v1 = oracle.coset_evaluate(np.ascontiguousarray(coeffs[c::3]), 1, g, order)
        assert np.array_equal(v1, vals[c::3])
    rc, vals_n = oracle.coset_evaluate(coeffs, 3, g, n)
    rc, back = oracle.coset_interpolate(vals_n, 3, g)
    assert np.array_equal(back, coeffs)


def test_poly_multiply_naive_equals_fast_and_doc_example(oracle):
    """polynomial.rs doc example (2 + 3x)^2 = 4 + 12x + 9x^2 (:806-808) and the reference's property
    `fast_multiply == naive_multiply` for both field types."""
    import oracle as o

    a = oracle.to_raw([2, 3])
    sq = oracle.poly_naive_multiply(a, a, 1)
    assert [int(v) for v in oracle.to_values(sq)] == [4, 12, 9]
    rc, sqf = oracle.poly_fast_multiply(a, a, 1)
    assert rc == 0 and np.array_equal(sq, sqf)
    for w in (1, 3):
        for na, nb in ((1, 1), (1, 7), (5, 3), (33, 31), (64, 65), (200, 57)):
            x, y = o.splitmix64_words(na * 31 + nb, na * w), o.splitmix64_words(nb * 17 + na + 5, nb * w)
            rc, f = oracle.poly_fast_multiply(x, y, w)
            assert rc == 0
            assert np.array_equal(f, oracle.poly_naive_multiply(x, y, w)), (w, na, nb)
    # XFE product restates x_field_element.rs:512-535: x * x^2 = x^3 = x - 1 (mod x^3 - x + 1)
    one, zero = oracle.to_raw([1])[0], np.uint64(0)
    xx = np.array([zero, one, zero], dtype=np.uint64)
    x2 = np.array([zero, zero, one], dtype=np.uint64)
    got = oracle.poly_naive_multiply(xx, x2, 3)
    assert [int(v) for v in oracle.to_values(got)] == [(1 << 64) - (1 << 32) + 1 - 1, 1, 0]


# ---- next wave: authentication structures and MMR bulk operations ---------------
def test_auth_structure_node_indices_reference_vectors(oracle):
    """merkle_tree.rs:1594-1609 and the doc example :590-604"""
    assert oracle.auth_structure_node_indices(8, [0, 1])[1].tolist() == [5, 3]
    assert oracle.auth_structure_node_indices(8, [0, 2])[1].tolist() == [11, 9, 3]
    assert oracle.auth_structure_node_indices(8, [4, 5, 6, 7])[1].tolist() == [2]
    assert oracle.auth_structure_node_indices(8, [2])[1].tolist() == [11, 4, 3]
    assert oracle.auth_structure_node_indices(8, [8])[0] == -9   # LeafIndexInvalid
    assert oracle.auth_structure_node_indices(6, [0])[0] == -4   # IncorrectNumberOfLeafs
    assert oracle.auth_structure_node_indices(1, [0])[1].tolist() == []


def test_abi_auth_structure_node_indices_match_oracle_without_gpu(oracle):
    """the index logic of the product is host code: same vectors + random cases, no device needed"""
    import importlib

    tf = importlib.import_module("twenty-first_b200")
    assert tf.MerkleTree.authentication_structure_node_indices(8, [0, 2]).tolist() == [11, 9, 3]
    rng = np.random.default_rng(7)
    for height in (0, 1, 2, 5, 9, 16, 30):
        n = 1 << height
        for k in (0, 1, 2, 7, 40):
            idx = rng.integers(0, n, size=k).astype(np.uint64)
            rc, want = oracle.auth_structure_node_indices(n, idx)
            assert rc == 0
            assert np.array_equal(tf.MerkleTree.authentication_structure_node_indices(n, idx), want)
    import pytest

    with pytest.raises(tf.MerkleTreeError) as ei:
        tf.MerkleTree.authentication_structure_node_indices(8, [8])
    assert ei.value.kind == "LeafIndexInvalid"
    with pytest.raises(tf.MerkleTreeError) as ei:
        tf.MerkleTree.authentication_structure_node_indices(12, [1])
    assert ei.value.kind == "IncorrectNumberOfLeafs"


def test_mmr_bag_peaks_reference_snapshot_and_peak_structure(oracle):
    """mmr_accumulator.rs:1038-1047: the snapshot of the empty MMR is the one vector that does not depend on
    Rust's StdRng; peaks_from_leafs is checked structurally against the Merkle roots of the binary runs."""
    empty = np.zeros(0, dtype=np.uint64)
    assert oracle.digest_to_hex(oracle.mmr_bag_peaks(empty, 0)) == (
        "cd65052100640f0d27e5654f97c47e49899add2f265967ccbefee7264e9bc08f588542d9dc3d5ac5")
    for n in (0, 1, 2, 3, 5, 8, 11, 100, 255, 257):
        leafs = __import__("oracle").splitmix64_words(0x7000 + n, 5 * n)
        peaks = oracle.mmr_peaks_from_leafs(leafs)
        assert peaks.shape[0] == bin(n).count("1")
        off, k = 0, 0
        for b in range(n.bit_length() - 1, -1, -1):
            if (n >> b) & 1:
                rc, root = oracle.merkle_sequential_frugal_root(leafs[5 * off: 5 * (off + (1 << b))])
                assert rc == 0 and np.array_equal(peaks[k], root)
                off += 1 << b
                k += 1
        # bag_peaks of a single-peak MMR = hash_pair(peak, hash_10(count, 0..))
        if n and n & (n - 1) == 0:
            cnt = np.zeros(10, dtype=np.uint64)
            cnt[0] = oracle.bfe_new(n & 0xffffffff)
            cnt[1] = oracle.bfe_new(n >> 32)
            want = oracle.hash_pair(peaks[0].copy(), oracle.hash_10(cnt))
            assert np.array_equal(oracle.mmr_bag_peaks(peaks, n), want)


def test_tip5_mds_generated_equals_naive_round(oracle):
    """the reference's own differential test (tip5/naive.rs:94-106): for every state and round the scalar round
    (sbox_layer + mds_generated + round constants, tip5/mod.rs:175-253) equals NaiveTip5::round -- here over random
    states, the all-(p-1) state, small values and the degenerate-representation trigger of tip5/mod.rs:222-242"""
    from oracle import P, splitmix64_words

    states = [splitmix64_words(0x7155 + i, 16) % np.uint64(P) for i in range(200)]
    states.append(np.full(16, P - 1, dtype=np.uint64))
    states.append(np.arange(16, dtype=np.uint64))
    states.append(np.full(16, 0xFFFFFFFF00000000, dtype=np.uint64))
    for k, st in enumerate(states):
        for r in range(5):
            a, b = st.copy(), st.copy()
            oracle.tip5_round(a, r)
            oracle.tip5_round(b, r, naive=True)
            assert np.array_equal(a, b), (k, r)
            assert (a < np.uint64(P)).all()  # the round-constant addition leaves canonical words (:1122-1142)


def test_tip5_avx512_round_equals_scalar_round_and_kats(oracle, kats):
    """the reference keeps its Tip5 snapshots to pin the AVX-512 build against the scalar one (tip5/mod.rs:1285-1295):
    the restated AVX-512 round (oracle/tip5_avx512.c <- tip5/avx512.rs) must agree round by round and on the KATs"""
    if not oracle.tip5_avx512_available():
        pytest.skip("host without avx512f/bw/ifma/vbmi (or gcc without the intrinsics)")
    from oracle import P, splitmix64_words

    states = [splitmix64_words(0xA512 + i, 16) % np.uint64(P) for i in range(300)]
    states += [np.full(16, P - 1, dtype=np.uint64), np.zeros(16, dtype=np.uint64), np.arange(16, dtype=np.uint64),
               np.full(16, 0xFFFFFFFF00000000, dtype=np.uint64)]
    for k, st in enumerate(states):
        for r in range(5):
            a, b = st.copy(), st.copy()
            oracle.tip5_round(a, r)
            assert oracle.tip5_round_avx512(b, r)
            assert np.array_equal(a, b), (k, r)
    try:
        assert oracle.tip5_set_impl("avx512") == "avx512"
        test_tip5_hash10_snapshot(oracle, kats)
        test_tip5_hash_varlen_snapshot(oracle, kats)
        test_tip5_raw_state_snapshot(oracle, kats)
        leafs = splitmix64_words(0xA513, 5 << 10) % np.uint64(P)
        rc, nodes_avx = oracle.merkle_par_new(leafs)
        oracle.tip5_set_impl("scalar")
        rc2, nodes_scalar = oracle.merkle_par_new(leafs)
        assert rc == 0 and rc2 == 0 and np.array_equal(nodes_avx, nodes_scalar)
    finally:
        oracle.tip5_set_impl("scalar")
