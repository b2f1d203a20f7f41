//! Same public signatures as the reference crate for the hot path, bodies call the C ABI.
//! (reference: twenty-first v2.0.2; paths below are relative to twenty-first/src/)
//!
//!   math::ntt::{ntt, intt}                         math/ntt.rs:67, 109
//!   Polynomial::fast_coset_{evaluate,interpolate}  math/polynomial.rs:1374, 1907
//!   Tip5::{permutation, hash_10, hash_pair, hash_varlen}  tip5/mod.rs:529-623
//!   MerkleTree::{par_new, sequential_new, par_frugal_root, sequential_frugal_root}
//!                                                  util_types/merkle_tree.rs:149-364
//!
//! A maintainer who wants the GPU path inside the crate itself replaces the bodies of those
//! items with the bodies below ("in-crate patch").  NOTE: uncompiled here --
//! the build container has no Rust toolchain.
use core::mem::size_of;
use tf21_sys as sys;
use twenty_first::prelude::*;
use twenty_first::util_types::merkle_tree::MerkleTreeError;

// Layout facts the FFI relies on: both field types are repr(transparent) over
// u64 words in Montgomery form; a Digest is five of them.
const _: () = assert!(size_of::<BFieldElement>() == 8);
const _: () = assert!(size_of::<XFieldElement>() == 24);
const _: () = assert!(size_of::<Digest>() == 40);

#[inline]
fn width_of<FF>() -> u32 {
    (size_of::<FF>() / 8) as u32 // 1 = BFieldElement, 3 = XFieldElement
}

#[cold]
fn fail(code: i32) -> ! {
    // NTT length violations are panics in the reference (ntt.rs:135-137)
    let msg = unsafe { std::ffi::CStr::from_ptr(sys::tf21_strerror(code)) };
    panic!("{}", msg.to_string_lossy());
}

/// math::ntt::ntt (ntt.rs:67-82): in place, natural order, panics on bad length.
pub fn ntt<FF: FiniteField>(x: &mut [FF]) {
    let rc = unsafe { sys::tf21_ntt(x.as_mut_ptr() as *mut u64, x.len() as u64, width_of::<FF>(), 1) };
    if rc != 0 {
        fail(rc)
    }
}

/// math::ntt::intt (ntt.rs:109-125)
pub fn intt<FF: FiniteField>(x: &mut [FF]) {
    let rc = unsafe { sys::tf21_intt(x.as_mut_ptr() as *mut u64, x.len() as u64, width_of::<FF>(), 1) };
    if rc != 0 {
        fail(rc)
    }
}

/// The caller-side `columns.par_iter_mut().for_each(|c| ntt(c))` pattern as one call:
/// `batch` contiguous columns of `n` elements.
pub fn ntt_batch<FF: FiniteField>(columns: &mut [FF], n: usize, inverse: bool) {
    let batch = if n == 0 { 0 } else { columns.len() / n } as u64;
    let p = columns.as_mut_ptr() as *mut u64;
    let rc = unsafe {
        if inverse { sys::tf21_intt(p, n as u64, width_of::<FF>(), batch) } else { sys::tf21_ntt(p, n as u64, width_of::<FF>(), batch) }
    };
    if rc != 0 {
        fail(rc)
    }
}

/// Polynomial::fast_coset_evaluate (polynomial.rs:1374-1399) for a BFieldElement offset.
pub fn fast_coset_evaluate<FF: FiniteField>(poly: &Polynomial<FF>, offset: BFieldElement, order: usize) -> Vec<FF> {
    let coeffs = poly.coefficients();
    let mut out = vec![FF::ZERO; order];
    let rc = unsafe {
        sys::tf21_coset_evaluate(coeffs.as_ptr() as *const u64, coeffs.len() as u64, width_of::<FF>(),
                                 offset.raw_u64(), order as u64, out.as_mut_ptr() as *mut u64)
    };
    if rc != 0 {
        fail(rc) // TF21_E_ORDER_LE_DEGREE carries the reference's assert message
    }
    out
}

/// Polynomial::fast_coset_interpolate (polynomial.rs:1907-1918)
pub fn fast_coset_interpolate<FF: FiniteField>(offset: BFieldElement, values: &[FF]) -> Polynomial<'static, FF> {
    let mut out = vec![FF::ZERO; values.len()];
    let rc = unsafe {
        sys::tf21_coset_interpolate(values.as_ptr() as *const u64, values.len() as u64, width_of::<FF>(),
                                    offset.raw_u64(), out.as_mut_ptr() as *mut u64)
    };
    if rc != 0 {
        fail(rc)
    }
    Polynomial::new(out)
}

/// Tip5::hash_pair over a batch (tip5/mod.rs:577-586): out[i] = hash_pair(pairs[i].0, pairs[i].1)
pub fn hash_pairs(pairs: &[[Digest; 2]]) -> Vec<Digest> {
    let mut out = vec![Digest::default(); pairs.len()];
    let rc = unsafe {
        sys::tf21_tip5_hash_pairs(pairs.as_ptr() as *const u64, pairs.len() as u64, out.as_mut_ptr() as *mut u64)
    };
    if rc != 0 {
        fail(rc)
    }
    out
}

/// Tip5::hash_varlen (tip5/mod.rs:617-623)
pub fn hash_varlen(input: &[BFieldElement]) -> Digest {
    let mut out = Digest::default();
    let rc = unsafe {
        sys::tf21_tip5_hash_varlen(input.as_ptr() as *const u64, input.len() as u64, &mut out as *mut Digest as *mut u64)
    };
    if rc != 0 {
        fail(rc)
    }
    out
}

fn merkle_err(code: i32) -> MerkleTreeError {
    match code {
        sys::TF21_E_TOO_FEW_LEAFS => MerkleTreeError::TooFewLeafs,
        sys::TF21_E_INCORRECT_NUMBER_OF_LEAFS => MerkleTreeError::IncorrectNumberOfLeafs,
        sys::TF21_E_ALLOC => MerkleTreeError::TreeTooHigh,
        sys::TF21_E_LEAF_INDEX_INVALID => MerkleTreeError::LeafIndexInvalid,
        other => fail(other),
    }
}

/// The node vector of MerkleTree::par_new / sequential_new (merkle_tree.rs:149-212): heap indexed,
/// nodes[0] = 0, nodes[1] = root, nodes[n..2n) = leafs.  (`MerkleTree { nodes }` has a private
/// field, so the in-crate patch constructs the struct; outside the crate this returns the nodes.)
pub fn merkle_nodes(leafs: &[Digest]) -> Result<Vec<Digest>, MerkleTreeError> {
    let mut nodes = vec![Digest::default(); 2 * leafs.len()];
    let rc = unsafe {
        sys::tf21_merkle_build(leafs.as_ptr() as *const u64, leafs.len() as u64, nodes.as_mut_ptr() as *mut u64)
    };
    if rc != 0 {
        return Err(merkle_err(rc));
    }
    Ok(nodes)
}

/// MerkleTree::par_frugal_root / sequential_frugal_root (merkle_tree.rs:299-364)
pub fn merkle_frugal_root(leafs: &[Digest]) -> Result<Digest, MerkleTreeError> {
    let mut root = Digest::default();
    let rc = unsafe {
        sys::tf21_merkle_root(leafs.as_ptr() as *const u64, leafs.len() as u64, &mut root as *mut Digest as *mut u64)
    };
    if rc != 0 {
        return Err(merkle_err(rc));
    }
    Ok(root)
}

/// MerkleTree::par_authentication_structure_from_leafs / sequential_… (merkle_tree.rs:514-542): the tree is
/// built once on the device and the needed nodes are gathered there (the reference computes one frugal root
/// per needed node); same digests, same (descending node index) order.
pub fn authentication_structure_from_leafs(leafs: &[Digest], leaf_indices: &[usize]) -> Result<Vec<Digest>, MerkleTreeError> {
    let idx: Vec<u64> = leaf_indices.iter().map(|&i| i as u64).collect();
    let mut count = 0u64;
    let rc = unsafe {
        sys::tf21_merkle_auth_structure_node_indices(leafs.len() as u64, idx.as_ptr(), idx.len() as u64,
                                                     core::ptr::null_mut(), 0, &mut count)
    };
    if rc != 0 && rc != sys::TF21_E_CAPACITY {
        return Err(merkle_err(rc));
    }
    let mut out = vec![Digest::default(); count as usize];
    let rc = unsafe {
        sys::tf21_merkle_authentication_structure_from_leafs(leafs.as_ptr() as *const u64, leafs.len() as u64,
                                                             idx.as_ptr(), idx.len() as u64,
                                                             out.as_mut_ptr() as *mut u64, count, &mut count)
    };
    if rc != 0 {
        return Err(merkle_err(rc));
    }
    Ok(out)
}

/// MmrAccumulator::new_from_leafs -> (peaks, leaf_count) (mmr/mmr_accumulator.rs:29-34, 96-115); any leaf count.
/// `MmrAccumulator::init(peaks, leaf_count)` (:25-27) turns the pair into the reference's struct.
pub fn mmr_peaks_from_leafs(leafs: &[Digest]) -> Vec<Digest> {
    let mut peaks = vec![Digest::default(); 64];
    let mut n_peaks = 0u64;
    let rc = unsafe {
        sys::tf21_mmr_peaks_from_leafs(leafs.as_ptr() as *const u64, leafs.len() as u64,
                                       peaks.as_mut_ptr() as *mut u64, &mut n_peaks)
    };
    if rc != 0 {
        fail(rc)
    }
    peaks.truncate(n_peaks as usize);
    peaks
}

/// bag_peaks (mmr/mmr_accumulator.rs:379-391)
pub fn mmr_bag_peaks(peaks: &[Digest], leaf_count: u64) -> Digest {
    let mut out = Digest::default();
    let rc = unsafe {
        sys::tf21_mmr_bag_peaks(peaks.as_ptr() as *const u64, peaks.len() as u64, leaf_count,
                                &mut out as *mut Digest as *mut u64)
    };
    if rc != 0 {
        fail(rc)
    }
    out
}
