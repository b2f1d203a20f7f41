// test_host.cpp -- parity tests of the C++ host mirror (twenty_first_b200.hpp) against the CPU oracle.
// They read like the reference's own tests: ntt.rs:397-469, 562-579 (KATs, round trip, NTT = evaluation),
// polynomial.rs:3645-3679 (coset evaluate / interpolate), tip5/mod.rs:1294-1362 (hash KATs),
// merkle_tree.rs:1025-1116, 1594-1609 (errors, accessors, authentication structure indices),
// mmr_accumulator.rs:1038-1047 (bag_peaks snapshot).
// Build: g++ -std=c++17 -O2 test_host.cpp -o test_host -L<dir of libtf21.so> -ltf21 -L<oracle/_build> -loracle
// The oracle is TEST INFRASTRUCTURE: it is linked here as the checker only.
#include <cstdio>
#include <cstdlib>
#include <functional>

#include "../../oracle/oracle.h"
#include "twenty_first_b200.hpp"

using namespace twenty_first;
using math::ntt::intt;
using math::ntt::ntt;

static int g_failed = 0, g_run = 0;
#define CHECK(cond)                                                        \
    do {                                                                   \
        if (!(cond)) {                                                     \
            std::printf("  FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); \
            g_failed++;                                                    \
            return;                                                        \
        }                                                                  \
    } while (0)

static uint64_t g_seed = 0x210000;
static uint64_t splitmix() {
    for (;;) {
        uint64_t z = (g_seed += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        z ^= z >> 31;
        if (z < BFieldElement::P) return z;  // uniform canonical raw word
    }
}
template <typename T>
static std::vector<T> random_elements(size_t n) {
    std::vector<T> v(n);
    uint64_t *w = words(v.data());
    for (size_t i = 0; i < n * sizeof(T) / 8; i++) w[i] = splitmix();
    return v;
}
static void run(const char *name, const std::function<void()> &f) {
    int before = g_failed;
    g_run++;
    f();
    std::printf("%s %s\n", g_failed == before ? "ok    " : "FAILED", name);
}

// ntt.rs:423-445: chu_ntt_b_field_prop_test-style KAT on n = 4
static void ntt_on_four_elements() {
    std::vector<BFieldElement> x = {BFieldElement::new_(1), BFieldElement::new_(4), BFieldElement::new_(0),
                                    BFieldElement::new_(0)};
    std::vector<BFieldElement> want = {BFieldElement::new_(5), BFieldElement::new_(1125899906842625ull),
                                       BFieldElement::new_(18446744069414584318ull),
                                       BFieldElement::new_(18445618169507741698ull)};
    auto orig = x;
    ntt(x);
    CHECK(x == want);
    intt(x);
    CHECK(x == orig);
}

template <typename FF>
static void ntt_matches_oracle_and_round_trips() {
    for (unsigned log2n : {0u, 1u, 5u, 10u, 11u, 12u, 13u, 16u, 20u}) {
        auto x = random_elements<FF>(size_t(1) << log2n);
        auto want = x;
        CHECK(oracle_ntt(words(want.data()), want.size(), width_of<FF>()) == 0);
        auto got = x;
        ntt(got);
        CHECK(got == want);
        intt(got);
        CHECK(got == x);
    }
}

// ntt.rs:135-137: panics for lengths that are not a power of two
static void ntt_panics_on_bad_length() {
    std::vector<BFieldElement> x(12);
    bool panicked = false;
    try {
        ntt(x);
    } catch (const Panic &p) {
        panicked = p.code == TF21_E_LEN_NOT_POW2;
    }
    CHECK(panicked);
    std::vector<BFieldElement> empty;
    ntt(empty);  // len 0 and 1 are no-ops (ntt.rs:178-181)
}

// polynomial.rs:3645-3679: coset evaluation == evaluation on the explicit coset; interpolation inverts it
static void fast_coset_evaluation_and_interpolation() {
    const BFieldElement offset = BFieldElement::generator();
    Polynomial<BFieldElement> poly(random_elements<BFieldElement>(37));
    const size_t order = 128;
    auto values = poly.fast_coset_evaluate(offset, order);
    const uint64_t omega = oracle_bfe_primitive_root_of_unity(order);
    uint64_t x = offset.raw;
    for (size_t i = 0; i < order; i++) {
        CHECK(values[i].raw == oracle_poly_evaluate(words(poly.coefficients.data()), poly.coefficients.size(), x));
        x = oracle_bfe_mul(x, omega);
    }
    auto back = Polynomial<BFieldElement>::fast_coset_interpolate(offset, values);
    for (size_t i = 0; i < order; i++)
        CHECK(back.coefficients[i].raw == (i < poly.coefficients.size() ? poly.coefficients[i].raw : 0));
    bool panicked = false;  // polynomial.rs:1388-1392: order must exceed the degree
    try {
        poly.fast_coset_evaluate(offset, 32);
    } catch (const Panic &p) {
        panicked = p.code == TF21_E_ORDER_LE_DEGREE;
    }
    CHECK(panicked);
    // XFieldElement coefficients against the oracle
    Polynomial<XFieldElement> xp(random_elements<XFieldElement>(100));
    auto xv = xp.fast_coset_evaluate(offset, 1024);
    std::vector<XFieldElement> want(1024);
    CHECK(oracle_coset_evaluate(words(xp.coefficients.data()), 100, 3, offset.raw, 1024, words(want.data())) == 0);
    CHECK(xv == want);
}

static void fast_multiply_and_square() {
    Polynomial<BFieldElement> a(random_elements<BFieldElement>(100)), b(random_elements<BFieldElement>(77));
    std::vector<BFieldElement> want(176);
    oracle_poly_naive_multiply(words(a.coefficients.data()), 100, words(b.coefficients.data()), 77, 1, words(want.data()));
    CHECK(a.fast_multiply(b).coefficients == want);
    CHECK(a.fast_multiply(b).coefficients == b.fast_multiply(a).coefficients);  // polynomial.rs:3412
    std::vector<BFieldElement> sq(199);
    oracle_poly_naive_multiply(words(a.coefficients.data()), 100, words(a.coefficients.data()), 100, 1, words(sq.data()));
    CHECK(a.fast_square().coefficients == sq);
    // clean_divide: (a * b) / b == a  (polynomial.rs:2358-2413)
    auto prod = a.fast_multiply(b);
    CHECK(prod.clean_divide(b).coefficients == a.coefficients);
}

// tip5/mod.rs:1294-1325: hash_10 of zeros / hash_varlen snapshots, through the oracle that is pinned to them
static void tip5_matches_oracle() {
    BFieldElement in[10];
    for (auto &e : in) e.raw = splitmix();
    Digest want;
    oracle_tip5_hash_10(words(in), words(want.values));
    CHECK(Tip5::hash_10(in) == want);
    Digest l, r;
    for (auto &e : l.values) e.raw = splitmix();
    for (auto &e : r.values) e.raw = splitmix();
    oracle_tip5_hash_pair(words(l.values), words(r.values), words(want.values));
    CHECK(Tip5::hash_pair(l, r) == want);
    for (size_t len : {size_t(0), size_t(1), size_t(10), size_t(23), size_t(1000)}) {
        auto x = random_elements<BFieldElement>(len);
        uint64_t dummy = 0;
        oracle_tip5_hash_varlen(len ? words(x.data()) : &dummy, len, words(want.values));
        CHECK(Tip5::hash_varlen(x) == want);
    }
    Tip5 sponge;
    for (auto &e : sponge.state) e.raw = splitmix();
    uint64_t st[16];
    std::memcpy(st, sponge.state, sizeof(st));
    oracle_tip5_permutation(st);
    sponge.permutation();
    CHECK(std::memcmp(st, sponge.state, sizeof(st)) == 0);
    // tip5/mod.rs:1294-1306: the reference's snapshot of hash_10 over [0; 10]... first word, via to_hex of hash_pair(0,0)
    uint32_t want_idx[33];
    uint64_t st2[16];
    std::memcpy(st2, sponge.state, sizeof(st2));
    oracle_tip5_sample_indices(st2, 1u << 20, 33, want_idx);
    auto got_idx = sponge.sample_indices(1u << 20, 33);
    for (int i = 0; i < 33; i++) CHECK(got_idx[i] == want_idx[i]);
}

// merkle_tree.rs:1025-1116: errors, root == frugal root, accessors
static void merkle_tree_behaviour() {
    for (unsigned h : {0u, 1u, 4u, 9u, 13u}) {
        auto leafs = random_elements<Digest>(size_t(1) << h);
        auto tree = MerkleTree::par_new(leafs);
        std::vector<Digest> want(2 * leafs.size());
        CHECK(oracle_merkle_sequential_new(words(leafs.data()), leafs.size(), words(want.data())) == 0);
        for (size_t i = 1; i < want.size(); i++) CHECK(*tree.node(i) == want[i]);
        CHECK(tree.node(0) == nullptr);
        CHECK(tree.num_leafs() == leafs.size() && tree.height() == h);
        CHECK(tree.leafs() == leafs);
        CHECK(MerkleTree::par_frugal_root(leafs) == tree.root());
        CHECK(MerkleTree::sequential_frugal_root(leafs) == tree.root());
    }
    auto kind_of = [](const std::function<void()> &f) {
        try {
            f();
        } catch (const MerkleTreeError &e) {
            return (int)e.kind;
        }
        return -1;
    };
    std::vector<Digest> none, three(3);
    CHECK(kind_of([&] { MerkleTree::par_new(none); }) == MerkleTreeError::TooFewLeafs);
    CHECK(kind_of([&] { MerkleTree::par_new(three); }) == MerkleTreeError::IncorrectNumberOfLeafs);
    CHECK(kind_of([&] { MerkleTree::par_frugal_root(none); }) == MerkleTreeError::IncorrectNumberOfLeafs);
    CHECK(kind_of([&] { MerkleTree::sequential_frugal_root(none); }) == MerkleTreeError::TooFewLeafs);
}

// merkle_tree.rs:1594-1609 + :590-604
static void authentication_structures() {
    using V = std::vector<uint64_t>;
    CHECK((MerkleTree::authentication_structure_node_indices(8, {0, 1}) == V{5, 3}));
    CHECK((MerkleTree::authentication_structure_node_indices(8, {0, 2}) == V{11, 9, 3}));
    CHECK((MerkleTree::authentication_structure_node_indices(8, {4, 5, 6, 7}) == V{2}));
    auto leafs = random_elements<Digest>(1 << 10);
    auto tree = MerkleTree::par_new(leafs);
    V idx = {3, 77, 78, 1000};
    auto a = tree.authentication_structure(idx);
    auto b = MerkleTree::par_authentication_structure_from_leafs(leafs, idx);
    CHECK(a == b && !a.empty());
    bool err = false;
    try {
        MerkleTree::authentication_structure_node_indices(8, {8});
    } catch (const MerkleTreeError &e) {
        err = e.kind == MerkleTreeError::LeafIndexInvalid;
    }
    CHECK(err);
}

// mmr_accumulator.rs:1038-1047
static void mmr_bag_peaks_snapshot() {
    auto empty = MmrAccumulator::new_from_leafs({});
    CHECK(empty.bag_peaks().to_hex() == "cd65052100640f0d27e5654f97c47e49899add2f265967ccbefee7264e9bc08f588542d9dc3d5ac5");
    auto leafs = random_elements<Digest>(11);
    auto mmr = MmrAccumulator::new_from_leafs(leafs);
    uint64_t want_peaks[64 * 5];
    uint64_t k = oracle_mmr_peaks_from_leafs(words(leafs.data()), leafs.size(), want_peaks);
    CHECK(k == mmr.peaks().size() && k == 3);
    CHECK(std::memcmp(want_peaks, mmr.peaks().data(), k * 40) == 0);
    Digest want;
    oracle_mmr_bag_peaks(want_peaks, k, 11, words(want.values));
    CHECK(mmr.bag_peaks() == want);
}

int main(int argc, char **argv) {
    if (argc > 1 && std::string(argv[1]) == "--link-only") {
        // CPU-side check: the program links against libtf21.so and the error path needs no device
        std::vector<BFieldElement> x(12);
        try {
            ntt(x);
        } catch (const Panic &p) {
            std::printf("link ok: %s\n", p.what());
            return p.code == TF21_E_LEN_NOT_POW2 ? 0 : 1;
        }
        return 1;
    }
    check(tf21_init(0));
    run("ntt_on_four_elements", ntt_on_four_elements);
    run("bfe_ntt_matches_oracle_and_round_trips", ntt_matches_oracle_and_round_trips<BFieldElement>);
    run("xfe_ntt_matches_oracle_and_round_trips", ntt_matches_oracle_and_round_trips<XFieldElement>);
    run("ntt_panics_on_bad_length", ntt_panics_on_bad_length);
    run("fast_coset_evaluation_and_interpolation", fast_coset_evaluation_and_interpolation);
    run("fast_multiply_and_square", fast_multiply_and_square);
    run("tip5_matches_oracle", tip5_matches_oracle);
    run("merkle_tree_behaviour", merkle_tree_behaviour);
    run("authentication_structures", authentication_structures);
    run("mmr_bag_peaks_snapshot", mmr_bag_peaks_snapshot);
    std::printf("%d tests, %d failed\n", g_run, g_failed);
    return g_failed ? 1 : 0;
}
