// twenty_first_b200.hpp -- C++17 host mirror of the reference crate's interface for the hot path,
// header only, over the C ABI of include/tf21.h (links libtf21.so).
//
// The reference is compiled code (Rust); this image has no Rust toolchain, so next to the uncompiled Rust
// shim (host/rust/) this header is the compiled host side above the boundary: same item names, argument
// meaning and error behaviour as the crate (reference: twenty-first v2.0.2, paths relative to
// twenty-first/src/):
//
//   math::ntt::{ntt, intt}                                  math/ntt.rs:67-82, 109-125
//   Polynomial::{fast_coset_evaluate, fast_coset_interpolate, fast_multiply, fast_square,
//                par_batch_coset_extrapolate}               math/polynomial.rs:780-932, 1374-1399, 1907-1918, 2255-2331
//   Tip5::{permutation, hash_10, hash_pair, hash_varlen, sample_indices}   tip5/mod.rs:529-656
//   MerkleTree::{par_new, sequential_new, par_frugal_root, sequential_frugal_root, root, node, leafs,
//                num_leafs, height, authentication_structure,
//                par_authentication_structure_from_leafs}   util_types/merkle_tree.rs:149-364, 449-653
//   MmrAccumulator::{new_from_leafs, init, peaks, bag_peaks} util_types/mmr/mmr_accumulator.rs:25-34, 96-134, 379-391
//
// Where the reference panics this throws `twenty_first::Panic` (a std::logic_error with the reference's
// message); where it returns `Err(MerkleTreeError::..)` this throws `twenty_first::MerkleTreeError`.
// There is no CPU fallback: without a usable CUDA device every call throws Panic("CUDA error ..").
#pragma once
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

#include "../../include/tf21.h"

namespace twenty_first {

struct Panic : std::logic_error {
    int code;
    Panic(int c, const std::string &msg) : std::logic_error(msg), code(c) {}
};

[[noreturn]] inline void fail(int code) {
    std::string msg = tf21_strerror(code);
    if (code == TF21_E_CUDA) msg += std::string(": ") + tf21_last_cuda_error();
    throw Panic(code, msg);
}
inline void check(int code) {
    if (code != 0) fail(code);
}

// ---- field elements: exactly the memory layout of the Rust types ----------------------------------------
// BFieldElement is #[repr(transparent)] u64 holding value * 2^64 mod p (b_field_element.rs:84-86, 235-237).
struct BFieldElement {
    static constexpr uint64_t P = 0xFFFFFFFF00000001ull;  // b_field_element.rs:225
    static constexpr uint64_t MAX = P - 1;                // :226
    uint64_t raw = 0;

    static uint64_t mul_mod(uint64_t a, uint64_t b) { return (uint64_t)(((unsigned __int128)a * b) % P); }
    static BFieldElement new_(uint64_t value) {  // `new`, :235-237: Montgomery form
        BFieldElement e;
        e.raw = mul_mod(value % P, 0xFFFFFFFFull);  // 2^64 mod p = 2^32 - 1
        return e;
    }
    static BFieldElement from_raw_u64(uint64_t r) {  // :414
        BFieldElement e;
        e.raw = r;
        return e;
    }
    uint64_t raw_u64() const { return raw; }  // :419
    uint64_t value() const {                  // :248, montyred(raw) = raw * 2^-64 mod p
        return mul_mod(raw, 0xFFFFFFFE00000001ull);  // (2^64)^-1 mod p
    }
    static BFieldElement generator() { return new_(7); }  // :312
    bool operator==(const BFieldElement &o) const { return raw == o.raw; }
    bool operator!=(const BFieldElement &o) const { return raw != o.raw; }
};
static_assert(sizeof(BFieldElement) == 8, "BFieldElement must be one u64 word");

// XFieldElement is #[repr(transparent)] [BFieldElement; 3] (x_field_element.rs:56-59)
struct XFieldElement {
    BFieldElement coefficients[3];
    bool operator==(const XFieldElement &o) const {
        return coefficients[0] == o.coefficients[0] && coefficients[1] == o.coefficients[1] &&
               coefficients[2] == o.coefficients[2];
    }
};
static_assert(sizeof(XFieldElement) == 24, "XFieldElement must be three u64 words");

// Digest: five BFieldElements (tip5/digest.rs:28-29, 49)
struct Digest {
    static constexpr size_t LEN = 5;
    BFieldElement values[5];
    bool operator==(const Digest &o) const { return std::memcmp(values, o.values, sizeof(values)) == 0; }
    bool operator!=(const Digest &o) const { return !(*this == o); }
    // digest.rs:144-152: canonical values, little-endian bytes, lower-case hex
    std::string to_hex() const {
        static const char *hex = "0123456789abcdef";
        std::string s;
        for (const auto &v : values) {
            uint64_t x = v.value();
            for (int b = 0; b < 8; b++) {
                unsigned byte = (unsigned)(x >> (8 * b)) & 0xff;
                s.push_back(hex[byte >> 4]);
                s.push_back(hex[byte & 15]);
            }
        }
        return s;
    }
};
static_assert(sizeof(Digest) == 40, "Digest must be five u64 words");

template <typename FF>
constexpr uint32_t width_of() {
    static_assert(std::is_same<FF, BFieldElement>::value || std::is_same<FF, XFieldElement>::value,
                  "FF is BFieldElement or XFieldElement");
    return (uint32_t)(sizeof(FF) / 8);
}
template <typename T>
inline uint64_t *words(T *p) {
    return reinterpret_cast<uint64_t *>(p);
}
template <typename T>
inline const uint64_t *words(const T *p) {
    return reinterpret_cast<const uint64_t *>(p);
}

// ---- math::ntt ----------------------------------------------------------------------------------------
namespace math {
namespace ntt {
// ntt.rs:67-82: in place, natural order in and out; panics unless the length is 0 or a power of two <= u32::MAX
template <typename FF>
inline void ntt(std::vector<FF> &x) {
    check(tf21_ntt(words(x.data()), x.size(), width_of<FF>(), 1));
}
// ntt.rs:109-125 (the unscale of :220-228 is fused on the device)
template <typename FF>
inline void intt(std::vector<FF> &x) {
    check(tf21_intt(words(x.data()), x.size(), width_of<FF>(), 1));
}
// the caller-side `columns.par_iter_mut().for_each(|c| ntt(c))` as one call over contiguous columns
template <typename FF>
inline void ntt_batch(std::vector<FF> &columns, size_t n, bool inverse) {
    const uint64_t batch = n ? columns.size() / n : 0;
    check(inverse ? tf21_intt(words(columns.data()), n, width_of<FF>(), batch)
                  : tf21_ntt(words(columns.data()), n, width_of<FF>(), batch));
}
}  // namespace ntt
}  // namespace math

// ---- math::polynomial::Polynomial (the NTT-backed fast paths) -------------------------------------------
template <typename FF>
struct Polynomial {
    std::vector<FF> coefficients;
    Polynomial() = default;
    explicit Polynomial(std::vector<FF> c) : coefficients(std::move(c)) {}

    // polynomial.rs:1374-1399; panics if order <= degree
    std::vector<FF> fast_coset_evaluate(BFieldElement offset, size_t order) const {
        std::vector<FF> out(order);
        check(tf21_coset_evaluate(words(coefficients.data()), coefficients.size(), width_of<FF>(), offset.raw, order,
                                  words(out.data())));
        return out;
    }
    // polynomial.rs:1907-1918
    static Polynomial fast_coset_interpolate(BFieldElement offset, const std::vector<FF> &values) {
        Polynomial p;
        p.coefficients.resize(values.size());
        check(tf21_coset_interpolate(words(values.data()), values.size(), width_of<FF>(), offset.raw,
                                     words(p.coefficients.data())));
        return p;
    }
    // polynomial.rs:900-932 (operands of the same field)
    Polynomial fast_multiply(const Polynomial &other) const {
        const size_t na = coefficients.size(), nb = other.coefficients.size();
        Polynomial p;
        p.coefficients.resize(na && nb ? na + nb - 1 : 0);
        check(tf21_poly_mul(words(coefficients.data()), na, words(other.coefficients.data()), nb, width_of<FF>(),
                            words(p.coefficients.data())));
        return p;
    }
    // polynomial.rs:780-802
    Polynomial fast_square() const {
        const size_t na = coefficients.size();
        Polynomial p;
        p.coefficients.resize(na ? 2 * na - 1 : 0);
        check(tf21_poly_square(words(coefficients.data()), na, width_of<FF>(), words(p.coefficients.data())));
        return p;
    }
    // polynomial.rs:2358-2413 (Polynomial<BFieldElement> only): exact quotient of a division without remainder
    Polynomial clean_divide(const Polynomial &divisor) const {
        static_assert(std::is_same<FF, BFieldElement>::value, "clean_divide is defined for Polynomial<BFieldElement>");
        Polynomial q;
        q.coefficients.resize(coefficients.size() ? coefficients.size() : 1);
        uint64_t n_q = 0;
        check(tf21_poly_clean_divide(words(coefficients.data()), coefficients.size(), words(divisor.coefficients.data()),
                                     divisor.coefficients.size(), words(q.coefficients.data()), &n_q));
        q.coefficients.resize(n_q);
        return q;
    }
    // polynomial.rs:2188-2331: result[(codeword, point)] flattened like the reference's flat_map
    static std::vector<FF> par_batch_coset_extrapolate(BFieldElement domain_offset, size_t codeword_length,
                                                       const std::vector<FF> &codewords,
                                                       const std::vector<FF> &points) {
        const size_t n_cw = codeword_length ? codewords.size() / codeword_length : 0;
        std::vector<FF> out(n_cw * points.size());
        check(tf21_batch_coset_extrapolate(domain_offset.raw, codeword_length, words(codewords.data()), n_cw,
                                           width_of<FF>(), words(points.data()), points.size(), words(out.data())));
        return out;
    }
};

// ---- tip5::Tip5 -------------------------------------------------------------------------------------------
struct Tip5 {
    static constexpr size_t STATE_SIZE = 16, RATE = 10;  // tip5/mod.rs:27-32
    BFieldElement state[16];

    // tip5/mod.rs:529-533
    void permutation() { check(tf21_tip5_permute(words(state), 1)); }
    // batch forms: a leading batch axis replaces the caller's par_iter().map(..) (benches/tip5.rs:43-49)
    static void permutation_batch(std::vector<BFieldElement> &states) {
        check(tf21_tip5_permute(words(states.data()), states.size() / 16));
    }
    // tip5/mod.rs:559-569
    static Digest hash_10(const BFieldElement (&input)[10]) {
        Digest d;
        check(tf21_tip5_hash_10(words(input), 1, words(d.values)));
        return d;
    }
    // tip5/mod.rs:577-586
    static Digest hash_pair(const Digest &left, const Digest &right) {
        uint64_t in[10];
        std::memcpy(in, left.values, 40);
        std::memcpy(in + 5, right.values, 40);
        Digest d;
        check(tf21_tip5_hash_pairs(in, 1, words(d.values)));
        return d;
    }
    static std::vector<Digest> hash_pairs(const std::vector<Digest> &left_right_interleaved) {
        std::vector<Digest> out(left_right_interleaved.size() / 2);
        check(tf21_tip5_hash_pairs(words(left_right_interleaved.data()), out.size(), words(out.data())));
        return out;
    }
    // tip5/mod.rs:617-623
    static Digest hash_varlen(const std::vector<BFieldElement> &input) {
        Digest d;
        check(tf21_tip5_hash_varlen(words(input.data()), input.size(), words(d.values)));
        return d;
    }
    // hash_varlen of every row of a row-major table: the leaves of the table's Merkle tree
    static std::vector<Digest> hash_rows(const std::vector<BFieldElement> &rows, size_t row_len) {
        std::vector<Digest> out(row_len ? rows.size() / row_len : 0);
        check(tf21_tip5_hash_rows(words(rows.data()), row_len, out.size(), words(out.data())));
        return out;
    }
    // tip5/mod.rs:636-656; panics unless upper_bound is a power of two
    std::vector<uint32_t> sample_indices(uint32_t upper_bound, size_t num_indices) {
        std::vector<uint32_t> out(num_indices);
        check(tf21_tip5_sample_indices(words(state), upper_bound, num_indices, out.data()));
        return out;
    }
};

// ---- util_types::merkle_tree ------------------------------------------------------------------------------
struct MerkleTreeError : std::runtime_error {
    enum Kind { TooFewLeafs, IncorrectNumberOfLeafs, TreeTooHigh, LeafIndexInvalid } kind;
    MerkleTreeError(Kind k, const char *what) : std::runtime_error(what), kind(k) {}
};
inline void merkle_check(int code) {
    switch (code) {
        case 0: return;
        case TF21_E_TOO_FEW_LEAFS: throw MerkleTreeError(MerkleTreeError::TooFewLeafs, "TooFewLeafs");  // :394-396
        case TF21_E_INCORRECT_NUMBER_OF_LEAFS:
            throw MerkleTreeError(MerkleTreeError::IncorrectNumberOfLeafs, "IncorrectNumberOfLeafs");  // :398-401
        case TF21_E_ALLOC: throw MerkleTreeError(MerkleTreeError::TreeTooHigh, "TreeTooHigh");  // :403-410
        case TF21_E_LEAF_INDEX_INVALID:
            throw MerkleTreeError(MerkleTreeError::LeafIndexInvalid, "LeafIndexInvalid");  // :487-489
        default: fail(code);
    }
}

class MerkleTree {
    std::vector<Digest> nodes_;  // heap indexed like the reference's Vec<Digest> (merkle_tree.rs:85-88)

   public:
    // merkle_tree.rs:165-212 (and sequential_new :149-153: same result by definition)
    static MerkleTree par_new(const std::vector<Digest> &leafs) {
        MerkleTree t;
        t.nodes_.resize(2 * leafs.size());
        merkle_check(tf21_merkle_build(words(leafs.data()), leafs.size(), words(t.nodes_.data())));
        return t;
    }
    static MerkleTree sequential_new(const std::vector<Digest> &leafs) { return par_new(leafs); }
    // merkle_tree.rs:332-364: checks the power of two first
    static Digest par_frugal_root(const std::vector<Digest> &leafs) {
        const size_t n = leafs.size();
        if (n == 0 || (n & (n - 1))) merkle_check(TF21_E_INCORRECT_NUMBER_OF_LEAFS);
        return sequential_frugal_root(leafs);
    }
    // merkle_tree.rs:299-309
    static Digest sequential_frugal_root(const std::vector<Digest> &leafs) {
        Digest root;
        merkle_check(tf21_merkle_root(words(leafs.data()), leafs.size(), words(root.values)));
        return root;
    }
    Digest root() const { return nodes_[1]; }              // :624
    size_t num_leafs() const { return nodes_.size() / 2; }  // :628
    uint32_t height() const {                               // :634
        uint32_t h = 0;
        while ((size_t(1) << h) < num_leafs()) h++;
        return h;
    }
    const Digest *node(size_t index) const {  // :644, None for index 0 or out of range
        return (index == 0 || index >= nodes_.size()) ? nullptr : &nodes_[index];
    }
    std::vector<Digest> leafs() const { return std::vector<Digest>(nodes_.begin() + num_leafs(), nodes_.end()); }  // :653

    // merkle_tree.rs:449-504
    static std::vector<uint64_t> authentication_structure_node_indices(uint64_t num_leafs,
                                                                       const std::vector<uint64_t> &leaf_indices) {
        uint64_t count = 0;
        int rc = tf21_merkle_auth_structure_node_indices(num_leafs, leaf_indices.data(), leaf_indices.size(), nullptr, 0,
                                                         &count);
        if (rc != TF21_E_CAPACITY) merkle_check(rc);
        std::vector<uint64_t> out(count);
        merkle_check(tf21_merkle_auth_structure_node_indices(num_leafs, leaf_indices.data(), leaf_indices.size(),
                                                             out.data(), out.size(), &count));
        return out;
    }
    // merkle_tree.rs:614-622
    std::vector<Digest> authentication_structure(const std::vector<uint64_t> &leaf_indices) const {
        std::vector<Digest> out;
        for (uint64_t idx : authentication_structure_node_indices(num_leafs(), leaf_indices)) out.push_back(nodes_[idx]);
        return out;
    }
    // merkle_tree.rs:514-542
    static std::vector<Digest> par_authentication_structure_from_leafs(const std::vector<Digest> &leafs,
                                                                       const std::vector<uint64_t> &leaf_indices) {
        const size_t need = authentication_structure_node_indices(leafs.size(), leaf_indices).size();
        std::vector<Digest> out(need);
        uint64_t count = 0;
        merkle_check(tf21_merkle_authentication_structure_from_leafs(words(leafs.data()), leafs.size(),
                                                                     leaf_indices.data(), leaf_indices.size(),
                                                                     words(out.data()), need, &count));
        return out;
    }
};

// ---- util_types::mmr::MmrAccumulator (bulk construction and commitment) -----------------------------------
class MmrAccumulator {
    std::vector<Digest> peaks_;
    uint64_t leaf_count_ = 0;

   public:
    static MmrAccumulator init(std::vector<Digest> peaks, uint64_t leaf_count) {  // mmr_accumulator.rs:25-27
        MmrAccumulator m;
        m.peaks_ = std::move(peaks);
        m.leaf_count_ = leaf_count;
        return m;
    }
    static MmrAccumulator new_from_leafs(const std::vector<Digest> &leafs) {  // :29-34, 96-115
        MmrAccumulator m;
        m.peaks_.resize(64);
        uint64_t n_peaks = 0;
        check(tf21_mmr_peaks_from_leafs(words(leafs.data()), leafs.size(), words(m.peaks_.data()), &n_peaks));
        m.peaks_.resize(n_peaks);
        m.leaf_count_ = leafs.size();
        return m;
    }
    const std::vector<Digest> &peaks() const { return peaks_; }  // :133-135
    uint64_t num_leafs() const { return leaf_count_; }           // :143-145
    Digest bag_peaks() const {                                   // :127-129, 379-391
        Digest d;
        check(tf21_mmr_bag_peaks(words(peaks_.data()), peaks_.size(), leaf_count_, words(d.values)));
        return d;
    }
};

}  // namespace twenty_first
