//! B200 bodies behind the reference crate's own names for the STARK hot path
//! (reference: twenty-first v2.0.2; paths below are relative to twenty-first/src/).
//!
//! | reference item                                                            | here                                   |
//! |---------------------------------------------------------------------------|----------------------------------------|
//! | `math::ntt::{ntt, intt}` (math/ntt.rs:67, 109)                            | [`ntt`], [`intt`], [`ntt_batch`]       |
//! | `Polynomial::{fast_coset_evaluate, fast_coset_interpolate}` (:1374, :1907) | [`PolynomialGpu`]                      |
//! | `Polynomial::{fast_multiply, fast_square, clean_divide, reduce_by_ntt_friendly_modulus, coset_extrapolate, batch_coset_extrapolate, par_batch_coset_extrapolate}` (:780-2413) | [`PolynomialGpu`] |
//! | `Tip5::{permutation, hash_10, hash_pair, hash_varlen, sample_indices}` (tip5/mod.rs:529-656) | [`Tip5Gpu`]         |
//! | `MerkleTree::{par_new, sequential_new, par_frugal_root, sequential_frugal_root, root, node, leaf, leafs, num_leafs, height, authentication_structure*, }` (util_types/merkle_tree.rs:149-665) | [`MerkleTree`] |
//! | `MmrAccumulator::{new_from_leafs, bag_peaks}` (util_types/mmr/mmr_accumulator.rs:29, 127)     | [`mmr`]             |
//!
//! Two ways to use it: (a) as this separate crate -- free functions, an extension trait on
//! `Polynomial` and a `MerkleTree` newtype with the reference's method names, so a caller changes `use` lines only;
//! (b) as an in-crate patch -- a maintainer pastes the bodies below into the reference's items.
//!
//! NOTE: UNCOMPILED in the build container (no cargo / rustc there).  The C ABI underneath is exercised by the
//! Python ctypes binding and the compiled C++17 mirror (host/cpp), which declare the same signatures.
use core::mem::size_of;
use tf21_sys as sys;
use twenty_first::prelude::*;
use twenty_first::util_types::merkle_tree::{MerkleTreeError, MerkleTreeHeight, MerkleTreeLeafIndex, MerkleTreeNodeIndex};

// Layout facts the FFI relies on: both field types are repr(transparent) over u64 words in
// Montgomery form; a Digest is five of them (tip5/digest.rs:28-29 has no explicit repr: checked here).
const _: () = assert!(size_of::<BFieldElement>() == 8);
const _: () = assert!(size_of::<XFieldElement>() == 24);
const _: () = assert!(size_of::<Digest>() == 40);

/// `tf21_init(n_devices)` (0 = all visible devices); optional -- every entry point prepares its device lazily.
pub fn init(n_devices: i32) -> Result<(), i32> {
    match unsafe { sys::tf21_init(n_devices) } {
        0 => Ok(()),
        rc => Err(rc),
    }
}

#[inline]
fn width_of<FF>() -> u32 {
    (size_of::<FF>() / 8) as u32 // 1 = BFieldElement, 3 = XFieldElement
}

#[cold]
fn fail(code: i32) -> ! {
    // length violations etc. are panics in the reference (ntt.rs:135-137, polynomial.rs:1388-1392, :556-559);
    // tf21_strerror carries the reference's message for each code
    let msg = unsafe { std::ffi::CStr::from_ptr(sys::tf21_strerror(code)) };
    panic!("{}", msg.to_string_lossy());
}

#[inline]
fn ok(rc: i32) {
    if rc != 0 {
        fail(rc)
    }
}

// ---- math::ntt ---------------------------------------------------------------------------------------------
/// `math::ntt::ntt` (ntt.rs:67-82): in place, natural order in and out, panics on a bad length.
pub fn ntt<FF: FiniteField>(x: &mut [FF]) {
    ok(unsafe { sys::tf21_ntt(x.as_mut_ptr() as *mut u64, x.len() as u64, width_of::<FF>(), 1) })
}

/// `math::ntt::intt` (ntt.rs:109-125), including the `unscale` by n^-1 (:220-228).
pub fn intt<FF: FiniteField>(x: &mut [FF]) {
    ok(unsafe { sys::tf21_intt(x.as_mut_ptr() as *mut u64, x.len() as u64, width_of::<FF>(), 1) })
}

/// The caller-side `columns.par_iter_mut().for_each(|c| ntt(c))` pattern (ntt.rs:250-269) as ONE call:
/// `columns.len() / n` contiguous columns of `n` elements each.
pub fn ntt_batch<FF: FiniteField>(columns: &mut [FF], n: usize, inverse: bool) {
    let batch = if n == 0 { 0 } else { columns.len() / n } as u64;
    let p = columns.as_mut_ptr() as *mut u64;
    ok(unsafe {
        if inverse { sys::tf21_intt(p, n as u64, width_of::<FF>(), batch) } else { sys::tf21_ntt(p, n as u64, width_of::<FF>(), batch) }
    })
}

/// Same, split over the devices of the box (columns shard independently, no exchange; include/tf21.h).
pub fn ntt_batch_sharded<FF: FiniteField>(columns: &mut [FF], n: usize, inverse: bool, n_shards: u32) {
    let batch = if n == 0 { 0 } else { columns.len() / n } as u64;
    ok(unsafe {
        sys::tf21_ntt_sharded(columns.as_mut_ptr() as *mut u64, n as u64, width_of::<FF>(), batch, inverse as i32, n_shards)
    })
}

// ---- math::polynomial --------------------------------------------------------------------------------------
/// The NTT-backed `Polynomial` methods under the reference's names.  `use twenty_first_b200::PolynomialGpu;`
/// brings them into scope next to the inherent ones; the in-crate patch replaces the inherent bodies instead.
/// Offsets are `BFieldElement`s (the reference is generic over `S`; every caller in the crate and in triton-vm
/// passes a base-field offset).
pub trait PolynomialGpu<FF: FiniteField> {
    /// `Polynomial::fast_coset_evaluate` (polynomial.rs:1374-1399): panics unless `order > degree`.
    fn fast_coset_evaluate(&self, offset: BFieldElement, order: usize) -> Vec<FF>;
    /// `Polynomial::fast_coset_interpolate` (polynomial.rs:1907-1918).
    fn fast_coset_interpolate(offset: BFieldElement, values: &[FF]) -> Polynomial<'static, FF>;
    /// interpolate on `(offset_in, values.len())`, evaluate on `(offset_out, n_out)`: the low-degree extension of
    /// BASELINE configs[3] in one call (the two `scale`s and the n^-1 fold into one factor per coefficient).
    fn coset_lde(values: &[FF], offset_in: BFieldElement, n_out: usize, offset_out: BFieldElement) -> Vec<FF>;
    /// `Polynomial::fast_multiply` (polynomial.rs:900-932), both operands of the same field.
    fn fast_multiply(&self, other: &Polynomial<FF>) -> Polynomial<'static, FF>;
    /// `Polynomial::fast_square` (polynomial.rs:780-802).
    fn fast_square(&self) -> Polynomial<'static, FF>;
    /// `Polynomial::reduce_by_ntt_friendly_modulus` (polynomial.rs:1087-1148).
    fn reduce_by_ntt_friendly_modulus(&self, shift_ntt: &[FF], tail_length: usize) -> Polynomial<'static, FF>;
    /// `Polynomial::coset_extrapolate` (polynomial.rs:2117-2129).
    fn coset_extrapolate(domain_offset: BFieldElement, codeword: &[FF], points: &[FF]) -> Vec<FF>;
    /// `Polynomial::{batch_,par_batch_}coset_extrapolate` (polynomial.rs:2188-2331).
    fn par_batch_coset_extrapolate(domain_offset: BFieldElement, codeword_length: usize, codewords: &[FF], points: &[FF]) -> Vec<FF>;
}

impl<FF: FiniteField> PolynomialGpu<FF> for Polynomial<'_, FF> {
    fn fast_coset_evaluate(&self, offset: BFieldElement, order: usize) -> Vec<FF> {
        let coeffs = self.coefficients();
        let mut out = vec![FF::ZERO; order];
        ok(unsafe {
            sys::tf21_coset_evaluate(coeffs.as_ptr() as *const u64, coeffs.len() as u64, width_of::<FF>(),
                                     offset.raw_u64(), order as u64, out.as_mut_ptr() as *mut u64)
        }); // TF21_E_ORDER_LE_DEGREE carries the reference's assert message
        out
    }

    fn fast_coset_interpolate(offset: BFieldElement, values: &[FF]) -> Polynomial<'static, FF> {
        let mut out = vec![FF::ZERO; values.len()];
        ok(unsafe {
            sys::tf21_coset_interpolate(values.as_ptr() as *const u64, values.len() as u64, width_of::<FF>(),
                                        offset.raw_u64(), out.as_mut_ptr() as *mut u64)
        });
        Polynomial::new(out)
    }

    fn coset_lde(values: &[FF], offset_in: BFieldElement, n_out: usize, offset_out: BFieldElement) -> Vec<FF> {
        let mut out = vec![FF::ZERO; n_out];
        ok(unsafe {
            sys::tf21_coset_lde(values.as_ptr() as *const u64, values.len() as u64, offset_in.raw_u64(), n_out as u64,
                                offset_out.raw_u64(), width_of::<FF>(), out.as_mut_ptr() as *mut u64)
        });
        out
    }

    fn fast_multiply(&self, other: &Polynomial<FF>) -> Polynomial<'static, FF> {
        let (a, b) = (self.coefficients(), other.coefficients());
        if a.is_empty() || b.is_empty() {
            return Polynomial::zero(); // polynomial.rs:909-911
        }
        let mut out = vec![FF::ZERO; a.len() + b.len() - 1];
        ok(unsafe {
            sys::tf21_poly_mul(a.as_ptr() as *const u64, a.len() as u64, b.as_ptr() as *const u64, b.len() as u64,
                               width_of::<FF>(), out.as_mut_ptr() as *mut u64)
        });
        Polynomial::new(out)
    }

    fn fast_square(&self) -> Polynomial<'static, FF> {
        let a = self.coefficients();
        if a.is_empty() {
            return Polynomial::zero();
        }
        let mut out = vec![FF::ZERO; 2 * a.len() - 1];
        ok(unsafe { sys::tf21_poly_square(a.as_ptr() as *const u64, a.len() as u64, width_of::<FF>(), out.as_mut_ptr() as *mut u64) });
        Polynomial::new(out)
    }

    fn reduce_by_ntt_friendly_modulus(&self, shift_ntt: &[FF], tail_length: usize) -> Polynomial<'static, FF> {
        let c = self.coefficients();
        let mut out = vec![FF::ZERO; c.len().min(shift_ntt.len()).max(1)];
        let mut n_out = 0u64;
        ok(unsafe {
            sys::tf21_poly_reduce_by_ntt_friendly_modulus(c.as_ptr() as *const u64, c.len() as u64, width_of::<FF>(),
                                                          shift_ntt.as_ptr() as *const u64, shift_ntt.len() as u64,
                                                          tail_length as u64, out.as_mut_ptr() as *mut u64, &mut n_out)
        });
        out.truncate(n_out as usize);
        Polynomial::new(out)
    }

    fn coset_extrapolate(domain_offset: BFieldElement, codeword: &[FF], points: &[FF]) -> Vec<FF> {
        Self::par_batch_coset_extrapolate(domain_offset, codeword.len(), codeword, points)
    }

    fn par_batch_coset_extrapolate(domain_offset: BFieldElement, codeword_length: usize, codewords: &[FF], points: &[FF]) -> Vec<FF> {
        let n_codewords = if codeword_length == 0 { 0 } else { codewords.len() / codeword_length };
        let mut out = vec![FF::ZERO; n_codewords * points.len()];
        ok(unsafe {
            sys::tf21_batch_coset_extrapolate(domain_offset.raw_u64(), codeword_length as u64, codewords.as_ptr() as *const u64,
                                              n_codewords as u64, width_of::<FF>(), points.as_ptr() as *const u64,
                                              points.len() as u64, out.as_mut_ptr() as *mut u64)
        });
        out
    }
}

/// `Polynomial<BFieldElement>::clean_divide` (polynomial.rs:2358-2413): `a / b` when the division leaves no remainder.
pub fn clean_divide(dividend: &Polynomial<BFieldElement>, divisor: &Polynomial<BFieldElement>) -> Polynomial<'static, BFieldElement> {
    let (a, b) = (dividend.coefficients(), divisor.coefficients());
    let mut q = vec![BFieldElement::ZERO; a.len().max(1)];
    let mut n_q = 0u64;
    ok(unsafe {
        sys::tf21_poly_clean_divide(a.as_ptr() as *const u64, a.len() as u64, b.as_ptr() as *const u64, b.len() as u64,
                                    q.as_mut_ptr() as *mut u64, &mut n_q)
    }); // TF21_E_DIVISION_BY_ZERO -> "divisor should be non-zero" (polynomial.rs:556-559)
    q.truncate(n_q as usize);
    Polynomial::new(q)
}

// ---- tip5 ---------------------------------------------------------------------------------------------------
/// `Tip5`'s hashing entry points (tip5/mod.rs:529-656) on the device.  Single hashes are latency bound (one
/// dependent permutation after the other): the batch forms are what a prover should call.
pub struct Tip5Gpu;

impl Tip5Gpu {
    /// `Tip5::permutation` (tip5/mod.rs:529-533) over a batch of sponges, in place.
    pub fn permutation(states: &mut [Tip5]) {
        const _: () = assert!(size_of::<Tip5>() == 128); // #[repr(align(64))] [BFieldElement; 16]
        ok(unsafe { sys::tf21_tip5_permute(states.as_mut_ptr() as *mut u64, states.len() as u64) })
    }

    /// `Tip5::hash_10` (tip5/mod.rs:559-569)
    pub fn hash_10(input: &[BFieldElement; 10]) -> [BFieldElement; Digest::LEN] {
        let mut out = [BFieldElement::ZERO; Digest::LEN];
        ok(unsafe { sys::tf21_tip5_hash_10(input.as_ptr() as *const u64, 1, out.as_mut_ptr() as *mut u64) });
        out
    }

    /// the caller-side `par_iter().map(Tip5::hash_10)` pattern (benches/tip5.rs:43-49) as one call
    pub fn hash_10_batch(inputs: &[[BFieldElement; 10]]) -> Vec<[BFieldElement; Digest::LEN]> {
        let mut out = vec![[BFieldElement::ZERO; Digest::LEN]; inputs.len()];
        ok(unsafe { sys::tf21_tip5_hash_10(inputs.as_ptr() as *const u64, inputs.len() as u64, out.as_mut_ptr() as *mut u64) });
        out
    }

    /// `Tip5::hash_pair` (tip5/mod.rs:577-586)
    pub fn hash_pair(left: Digest, right: Digest) -> Digest {
        Self::hash_pairs(&[[left, right]])[0]
    }

    /// out[i] = hash_pair(pairs[i][0], pairs[i][1])
    pub fn hash_pairs(pairs: &[[Digest; 2]]) -> Vec<Digest> {
        let mut out = vec![Digest::default(); pairs.len()];
        ok(unsafe { sys::tf21_tip5_hash_pairs(pairs.as_ptr() as *const u64, pairs.len() as u64, out.as_mut_ptr() as *mut u64) });
        out
    }

    /// `Tip5::hash_varlen` (tip5/mod.rs:617-623; padding of sponge.rs:41-56)
    pub fn hash_varlen(input: &[BFieldElement]) -> Digest {
        let mut out = Digest::default();
        ok(unsafe { sys::tf21_tip5_hash_varlen(input.as_ptr() as *const u64, input.len() as u64, &mut out as *mut Digest as *mut u64) });
        out
    }

    /// `hash_varlen` of every row of a row-major table: the step between NTT codewords and `MerkleTree::par_new`.
    pub fn hash_rows(rows: &[BFieldElement], row_len: usize) -> Vec<Digest> {
        let n_rows = if row_len == 0 { 0 } else { rows.len() / row_len };
        let mut out = vec![Digest::default(); n_rows];
        ok(unsafe { sys::tf21_tip5_hash_rows(rows.as_ptr() as *const u64, row_len as u64, n_rows as u64, out.as_mut_ptr() as *mut u64) });
        out
    }

    /// `Tip5::sample_indices` (tip5/mod.rs:636-656): `sponge` is updated like `&mut self`.
    pub fn sample_indices(sponge: &mut Tip5, upper_bound: u32, num_indices: usize) -> Vec<u32> {
        let mut out = vec![0u32; num_indices];
        ok(unsafe {
            sys::tf21_tip5_sample_indices(sponge.state.as_mut_ptr() as *mut u64, upper_bound, num_indices as u64, out.as_mut_ptr())
        }); // TF21_E_LEN_NOT_POW2 <-> assert!(upper_bound.is_power_of_two())
        out
    }
}

// ---- util_types::merkle_tree ----------------------------------------------------------------------------------
type Result<T> = core::result::Result<T, MerkleTreeError>;

fn merkle_err(code: i32) -> MerkleTreeError {
    match code {
        sys::TF21_E_TOO_FEW_LEAFS => MerkleTreeError::TooFewLeafs,
        sys::TF21_E_INCORRECT_NUMBER_OF_LEAFS => MerkleTreeError::IncorrectNumberOfLeafs,
        sys::TF21_E_ALLOC => MerkleTreeError::TreeTooHigh,
        sys::TF21_E_LEAF_INDEX_INVALID => MerkleTreeError::LeafIndexInvalid,
        other => fail(other),
    }
}

/// Same shape as the reference's `MerkleTree { nodes: Vec<Digest> }` (merkle_tree.rs:85-88; the field is private
/// there, hence the newtype outside the crate): heap indexed, nodes[0] = 0, nodes[1] = root, nodes[n..2n) = leafs.
#[derive(Debug, Clone, PartialEq, Eq)]
pub struct MerkleTree {
    nodes: Vec<Digest>,
}

impl MerkleTree {
    /// `MerkleTree::par_new` (merkle_tree.rs:165-212).  The thread count / parallelisation cutoff of the reference
    /// (config.rs:73) does not exist here and the result does not depend on it (merkle_tree.rs:1076-1087).
    pub fn par_new(leafs: &[Digest]) -> Result<Self> {
        let mut nodes = vec![Digest::default(); 2 * leafs.len()];
        match unsafe { sys::tf21_merkle_build(leafs.as_ptr() as *const u64, leafs.len() as u64, nodes.as_mut_ptr() as *mut u64) } {
            0 => Ok(Self { nodes }),
            rc => Err(merkle_err(rc)),
        }
    }

    /// `MerkleTree::sequential_new` (merkle_tree.rs:149-153): same tree.
    pub fn sequential_new(leafs: &[Digest]) -> Result<Self> {
        Self::par_new(leafs)
    }

    /// subtrees on `n_shards` devices, the tree cap gathered over NCCL (include/tf21.h `tf21_merkle_build_sharded`)
    pub fn par_new_sharded(leafs: &[Digest], n_shards: u32) -> Result<Self> {
        let mut nodes = vec![Digest::default(); 2 * leafs.len()];
        match unsafe {
            sys::tf21_merkle_build_sharded(leafs.as_ptr() as *const u64, leafs.len() as u64, nodes.as_mut_ptr() as *mut u64, n_shards)
        } {
            0 => Ok(Self { nodes }),
            rc => Err(merkle_err(rc)),
        }
    }

    /// `MerkleTree::par_frugal_root` (merkle_tree.rs:332-364)
    pub fn par_frugal_root(leafs: &[Digest]) -> Result<Digest> {
        let mut root = Digest::default();
        match unsafe { sys::tf21_merkle_root(leafs.as_ptr() as *const u64, leafs.len() as u64, &mut root as *mut Digest as *mut u64) } {
            0 => Ok(root),
            rc => Err(merkle_err(rc)),
        }
    }

    /// `MerkleTree::sequential_frugal_root` (merkle_tree.rs:299-309)
    pub fn sequential_frugal_root(leafs: &[Digest]) -> Result<Digest> {
        Self::par_frugal_root(leafs)
    }

    /// `MerkleTree::authentication_structure_node_indices` (merkle_tree.rs:449-504), descending, de-duplicated
    pub fn authentication_structure_node_indices(
        num_leafs: MerkleTreeLeafIndex,
        leaf_indices: &[MerkleTreeLeafIndex],
    ) -> Result<impl ExactSizeIterator<Item = MerkleTreeNodeIndex>> {
        let idx: Vec<u64> = leaf_indices.iter().map(|&i| i as u64).collect();
        let mut count = 0u64;
        let rc = unsafe {
            sys::tf21_merkle_auth_structure_node_indices(num_leafs as u64, idx.as_ptr(), idx.len() as u64, core::ptr::null_mut(), 0, &mut count)
        };
        if rc != 0 && rc != sys::TF21_E_CAPACITY {
            return Err(merkle_err(rc));
        }
        let mut out = vec![0u64; count as usize];
        let rc = unsafe {
            sys::tf21_merkle_auth_structure_node_indices(num_leafs as u64, idx.as_ptr(), idx.len() as u64, out.as_mut_ptr(), count, &mut count)
        };
        if rc != 0 {
            return Err(merkle_err(rc));
        }
        Ok(out.into_iter().map(|i| i as MerkleTreeNodeIndex))
    }

    /// `MerkleTree::{sequential,par}_authentication_structure_from_leafs` (merkle_tree.rs:514-542): one device build
    /// and a gather instead of one frugal root per needed node; same digests in the same order.
    pub fn par_authentication_structure_from_leafs(leafs: &[Digest], leaf_indices: &[MerkleTreeLeafIndex]) -> Result<Vec<Digest>> {
        let idx: Vec<u64> = leaf_indices.iter().map(|&i| i as u64).collect();
        let count = Self::authentication_structure_node_indices(leafs.len(), leaf_indices)?.len() as u64;
        let mut out = vec![Digest::default(); count as usize];
        let mut written = 0u64;
        match unsafe {
            sys::tf21_merkle_authentication_structure_from_leafs(leafs.as_ptr() as *const u64, leafs.len() as u64, idx.as_ptr(),
                                                                 idx.len() as u64, out.as_mut_ptr() as *mut u64, count, &mut written)
        } {
            0 => Ok(out),
            rc => Err(merkle_err(rc)),
        }
    }

    pub fn sequential_authentication_structure_from_leafs(leafs: &[Digest], leaf_indices: &[MerkleTreeLeafIndex]) -> Result<Vec<Digest>> {
        Self::par_authentication_structure_from_leafs(leafs, leaf_indices)
    }

    /// `MerkleTree::authentication_structure` (merkle_tree.rs:614-622): host-resident nodes, plain indexing
    pub fn authentication_structure(&self, leaf_indices: &[MerkleTreeLeafIndex]) -> Result<Vec<Digest>> {
        let indices = Self::authentication_structure_node_indices(self.num_leafs(), leaf_indices)?;
        Ok(indices.map(|idx| self.node(idx).unwrap()).collect())
    }

    pub fn root(&self) -> Digest {
        self.nodes[1] // merkle_tree.rs:624
    }
    pub fn num_leafs(&self) -> MerkleTreeLeafIndex {
        self.nodes.len() / 2 // :628
    }
    pub fn height(&self) -> MerkleTreeHeight {
        self.num_leafs().ilog2() // :634
    }
    pub fn node(&self, index: MerkleTreeNodeIndex) -> Option<Digest> {
        if index == 0 { None } else { self.nodes.get(index).copied() } // :644
    }
    pub fn leafs(&self) -> impl Iterator<Item = &Digest> {
        self.nodes.iter().skip(self.num_leafs()) // :653
    }
    pub fn leaf(&self, index: MerkleTreeLeafIndex) -> Option<Digest> {
        self.node(self.num_leafs() + index) // :658
    }
    pub fn indexed_leafs(&self, indices: &[MerkleTreeLeafIndex]) -> Result<Vec<(MerkleTreeLeafIndex, Digest)>> {
        indices.iter().map(|&i| self.leaf(i).map(|l| (i, l)).ok_or(MerkleTreeError::LeafIndexInvalid)).collect() // :665
    }
    /// the node vector, for `MerkleTree::try_from(nodes)`-style hand-over to the reference's own struct
    pub fn into_nodes(self) -> Vec<Digest> {
        self.nodes
    }
}

// ---- util_types::mmr -------------------------------------------------------------------------------------------
pub mod mmr {
    use super::*;

    /// `MmrAccumulator::new_from_leafs` (mmr/mmr_accumulator.rs:29-34, 96-115) for any leaf count:
    /// `MmrAccumulator::init(peaks, leafs.len() as u64)` turns the result into the reference's struct.
    pub fn peaks_from_leafs(leafs: &[Digest]) -> Vec<Digest> {
        let mut peaks = vec![Digest::default(); 64];
        let mut n_peaks = 0u64;
        ok(unsafe { sys::tf21_mmr_peaks_from_leafs(leafs.as_ptr() as *const u64, leafs.len() as u64, peaks.as_mut_ptr() as *mut u64, &mut n_peaks) });
        peaks.truncate(n_peaks as usize);
        peaks
    }

    /// `bag_peaks` (mmr/mmr_accumulator.rs:379-391)
    pub fn bag_peaks(peaks: &[Digest], leaf_count: u64) -> Digest {
        let mut out = Digest::default();
        ok(unsafe { sys::tf21_mmr_bag_peaks(peaks.as_ptr() as *const u64, peaks.len() as u64, leaf_count, &mut out as *mut Digest as *mut u64) });
        out
    }
}
