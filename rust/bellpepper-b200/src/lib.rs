//! `B200ConstraintSystem<Scalar>`: a drop-in for `bellpepper_core::test_cs::TestConstraintSystem` whose witness and
//! constraint matrices live in GPU memory and whose `which_is_unsatisfied` / `is_satisfied` run as CUDA kernels.
//!
//! Names, namespaces, closures and `LinearCombination` construction stay in Rust exactly as in the reference
//! (crates/bellpepper-core/src/util_cs/test_cs.rs:377-447); only flat data crosses the C ABI:
//!   * `alloc`/`alloc_input`  -> value pushed to a pending buffer, flushed with `bp_cs_alloc`
//!   * `enforce`              -> the three LCs flattened (`iter_inputs` then `iter_aux`, lc.rs:162-170) into
//!                               (tagged u32 column, 32-byte little-endian `to_repr()`) and flushed with `bp_cs_enforce`
//!   * `which_is_unsatisfied` -> `bp_cs_first_unsatisfied`, row mapped back to its path
//!
//! The accessors take `&self` exactly like the reference's (`which_is_unsatisfied`, `is_satisfied`, `get`:
//! test_cs.rs:239, 255, 311), so reference tests that call them on an immutable binding compile unchanged; the staging
//! buffers they may have to flush sit behind a `RefCell`.
//!
//! NOT COMPILED in this repository (no Rust toolchain in the build image); kept in sync with include/bp_r1cs.h.
pub mod ffi;

use std::cell::RefCell;
use std::collections::HashMap;
use std::ffi::CStr;
use std::marker::PhantomData;

use bellpepper_core::{ConstraintSystem, Index, LinearCombination, SynthesisError, Variable};
use ff::PrimeField;

/// Which of the three supported scalar fields `Scalar` is.  `new_on` asserts that `Scalar::MODULUS` is the modulus the
/// library uses for that id, so a wrong id cannot silently evaluate in another field.
pub trait B200Field: PrimeField {
    const FIELD_ID: i32;
}

/// `PrimeField::MODULUS` (hex, `0x` prefix, big-endian) of the field each id stands for (include/bp_r1cs.h).
const MODULI: [&str; 3] = [
    "73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001", // BP_FIELD_BLS12_381_FR
    "40000000000000000000000000000000224698fc0994a8dd8c46eb2100000001", // BP_FIELD_PALLAS_FR (pasta Fq)
    "40000000000000000000000000000000224698fc094cf91b992d30ed00000001", // BP_FIELD_VESTA_FR  (pasta Fp)
];

#[derive(Debug)]
enum NamedObject {
    Constraint(usize),
    Var(Variable),
    Namespace,
}

pub struct B200ConstraintSystem<Scalar: B200Field> {
    h: *mut ffi::bp_cs,
    named_objects: HashMap<String, NamedObject>,
    current_namespace: Vec<String>,
    constraint_paths: Vec<String>,
    input_names: Vec<String>,
    aux_names: Vec<String>,
    count: [u64; 2],
    // Staging buffers.  They sit behind a RefCell because the reference's accessors take `&self`
    // (`which_is_unsatisfied`, `is_satisfied`, `get`: test_cs.rs:239, 255, 311) and must still be able to flush.
    pend: RefCell<Pending>,
    _s: PhantomData<Scalar>,
}

/// Pending (not yet flushed) data.  Witness values are packed: one byte per value that fits a byte (AllocatedBit / Boolean
/// witnesses, i.e. nearly everything a gadget circuit allocates); a value that does not fit leaves a 0 placeholder and goes
/// to `wide` as (position in `bytes`, limbs), patched with bp_cs_set after the batch is appended.
#[derive(Default)]
struct Pending {
    bytes: [Vec<u8>; 2],
    wide: [Vec<(u64, [u64; 4])>; 2],
    lens: Vec<u32>,
    cols: Vec<u32>,
    coeffs: Vec<u64>,
}

// The handle is thread-compatible (no TLS, every call selects its device): `ConstraintSystem: Send` holds.
unsafe impl<Scalar: B200Field> Send for B200ConstraintSystem<Scalar> {}

/// Stage one witness value: a byte when it fits, else a placeholder plus an entry in the wide list.
fn stage_value<S: PrimeField>(s: &S, bytes: &mut Vec<u8>, wide: &mut Vec<(u64, [u64; 4])>) {
    let repr = s.to_repr();
    let r = repr.as_ref();
    if r[1..].iter().all(|&b| b == 0) {
        bytes.push(r[0]);
    } else {
        let mut l = [0u64; 4];
        for (i, chunk) in r.chunks_exact(8).enumerate() {
            l[i] = u64::from_le_bytes(chunk.try_into().unwrap());
        }
        wide.push((bytes.len() as u64, l));
        bytes.push(0);
    }
}

fn repr_to_limbs<S: PrimeField>(s: &S, out: &mut Vec<u64>) {
    let repr = s.to_repr(); // canonical little-endian 32 bytes (test_cs.rs:108-111 reverses it for big-endian)
    for chunk in repr.as_ref().chunks_exact(8) {
        out.push(u64::from_le_bytes(chunk.try_into().unwrap()));
    }
}

fn compute_path(ns: &[String], this: &str) -> String {
    assert!(!this.chars().any(|a| a == '/'), "'/' is not allowed in names");
    if ns.is_empty() {
        return this.to_string();
    }
    format!("{}/{}", ns.join("/"), this)
}

impl<Scalar: B200Field> B200ConstraintSystem<Scalar> {
    pub fn new_on(device: i32) -> Self {
        let id = Scalar::FIELD_ID;
        assert!((0..3).contains(&id), "unknown B200Field::FIELD_ID {id}");
        assert!(
            Scalar::MODULUS.trim_start_matches("0x").eq_ignore_ascii_case(MODULI[id as usize]),
            "B200Field::FIELD_ID {id} does not name the field whose modulus is {}",
            Scalar::MODULUS
        );
        let mut h = std::ptr::null_mut();
        let rc = unsafe { ffi::bp_cs_new(id, device, 0, 0, 0, &mut h) };
        assert_eq!(rc, ffi::BP_OK, "bp_cs_new failed ({rc}): no CUDA device? there is no CPU fallback");
        let mut named_objects = HashMap::new();
        named_objects.insert("ONE".into(), NamedObject::Var(Self::one()));
        Self {
            h,
            named_objects,
            current_namespace: vec![],
            constraint_paths: vec![],
            input_names: vec!["ONE".into()],
            aux_names: vec![],
            count: [1, 0],
            pend: RefCell::new(Pending::default()),
            _s: PhantomData,
        }
    }

    fn check(&self, rc: i32) {
        if rc != ffi::BP_OK {
            let msg = unsafe { CStr::from_ptr(ffi::bp_cs_last_error(self.h)) }.to_string_lossy().into_owned();
            panic!("bp_r1cs error {rc}: {msg}"); // BP_E_RANGE at check time == the reference's slice-index panic
        }
    }

    /// Send everything staged so far to the device (`&self`: see `pend`).
    pub fn flush(&self) {
        let mut p = self.pend.borrow_mut();
        for k in 0..2 {
            if !p.bytes[k].is_empty() {
                let n = p.bytes[k].len() as u64;
                let mut first = 0u64;
                let rc = unsafe { ffi::bp_cs_alloc_u8(self.h, k as i32, p.bytes[k].as_ptr(), n, &mut first) };
                self.check(rc);
                for (pos, limbs) in std::mem::take(&mut p.wide[k]) {
                    let rc = unsafe { ffi::bp_cs_set(self.h, k as i32, first + pos, limbs.as_ptr()) };
                    self.check(rc);
                }
                p.bytes[k].clear();
            }
        }
        if !p.lens.is_empty() {
            let rc = unsafe { ffi::bp_cs_enforce(self.h, (p.lens.len() / 3) as u64, p.lens.as_ptr(), p.cols.as_ptr(), p.coeffs.as_ptr()) };
            self.check(rc);
            p.lens.clear();
            p.cols.clear();
            p.coeffs.clear();
        }
    }

    fn push_lc(p: &mut Pending, lc: &LinearCombination<Scalar>) {
        let mut n = 0u32;
        for (i, c) in lc.iter_inputs() {
            p.cols.push(*i as u32);
            repr_to_limbs(c, &mut p.coeffs);
            n += 1;
        }
        for (i, c) in lc.iter_aux() {
            p.cols.push(*i as u32 | ffi::BP_COL_AUX);
            repr_to_limbs(c, &mut p.coeffs);
            n += 1;
        }
        p.lens.push(n);
    }

    fn set_named_obj(&mut self, path: String, to: NamedObject) {
        assert!(!self.named_objects.contains_key(&path), "tried to create object at existing path: {}", path);
        self.named_objects.insert(path, to);
    }

    /// test_cs.rs:239-253
    pub fn which_is_unsatisfied(&self) -> Option<&str> {
        self.flush();
        let mut row = 0i64;
        let rc = unsafe { ffi::bp_cs_first_unsatisfied(self.h, &mut row) };
        self.check(rc);
        if row < 0 { None } else { Some(&self.constraint_paths[row as usize]) }
    }

    /// test_cs.rs:255-264
    pub fn is_satisfied(&self) -> bool {
        match self.which_is_unsatisfied() {
            Some(b) => {
                println!("fail: {:?}", b);
                false
            }
            None => true,
        }
    }

    /// Same circuit, next witness (the `SizedWitness` flow, witness_cs.rs:7-41): replace every input and aux value and
    /// return the first unsatisfied constraint.  `inputs` includes ONE.  Values are sent one bit each when they are all
    /// 0 or 1, else one byte each; a value that does not fit a byte is patched with `bp_cs_set` before the check.
    pub fn recheck(&mut self, inputs: &[Scalar], aux: &[Scalar]) -> Option<&str> {
        self.flush();
        assert_eq!(inputs.len(), self.input_names.len());
        assert_eq!(aux.len(), self.aux_names.len());
        let mut bytes: [Vec<u8>; 2] = [Vec::with_capacity(inputs.len()), Vec::with_capacity(aux.len())];
        let mut wide: [Vec<(u64, [u64; 4])>; 2] = [vec![], vec![]];
        for v in inputs { stage_value(v, &mut bytes[0], &mut wide[0]); }
        for v in aux { stage_value(v, &mut bytes[1], &mut wide[1]); }
        let mut row = 0i64;
        if wide[0].is_empty() && wide[1].is_empty() {
            let all_bits = bytes.iter().all(|b| b.iter().all(|&x| x <= 1));
            let rc = if all_bits {
                let pack = |b: &Vec<u8>| -> Vec<u8> {
                    let mut out = vec![0u8; (b.len() + 7) / 8];
                    for (i, &x) in b.iter().enumerate() { out[i >> 3] |= x << (i & 7); }
                    out
                };
                let (pi, pa) = (pack(&bytes[0]), pack(&bytes[1]));
                unsafe { ffi::bp_cs_recheck_bits(self.h, pi.as_ptr(), pa.as_ptr(), &mut row) }
            } else {
                unsafe { ffi::bp_cs_recheck_u8(self.h, bytes[0].as_ptr(), bytes[1].as_ptr(), &mut row) }
            };
            self.check(rc);
        } else {
            for k in 0..2 {
                let rc = unsafe { ffi::bp_cs_set_range_u8(self.h, k as i32, 0, bytes[k].len() as u64, bytes[k].as_ptr()) };
                self.check(rc);
                for (pos, limbs) in &wide[k] {
                    let rc = unsafe { ffi::bp_cs_set(self.h, k as i32, *pos, limbs.as_ptr()) };
                    self.check(rc);
                }
            }
            let rc = unsafe { ffi::bp_cs_first_unsatisfied(self.h, &mut row) };
            self.check(rc);
        }
        if row < 0 { None } else { Some(&self.constraint_paths[row as usize]) }
    }

    /// `recheck` without touching the values on the Rust side: the two slices are handed over as they sit in memory.
    /// `blstrs::Scalar` and `pasta_curves::{Fp, Fq}` are `[u64; 4]` Montgomery limbs (x * 2^256 mod p, little-endian); the
    /// library's packing pass matches the limb patterns of 0 and 1 in that form and converts the few other elements itself
    /// (include/bp_r1cs.h: bp_cs_recheck_scalars_mont), so no `to_repr()` pass over the witness is needed.  A scalar type
    /// with another in-memory layout must use `recheck`.
    pub fn recheck_in_memory(&mut self, inputs: &[Scalar], aux: &[Scalar]) -> Option<&str> {
        assert_eq!(std::mem::size_of::<Scalar>(), 32, "Scalar is not four 64-bit Montgomery limbs");
        assert_eq!(std::mem::align_of::<Scalar>() % 8, 0);
        self.flush();
        assert_eq!(inputs.len(), self.input_names.len());
        assert_eq!(aux.len(), self.aux_names.len());
        let mut row = 0i64;
        let rc = unsafe {
            ffi::bp_cs_recheck_scalars_mont(self.h, inputs.as_ptr() as *const u64, aux.as_ptr() as *const u64, &mut row)
        };
        self.check(rc);
        if row < 0 { None } else { Some(&self.constraint_paths[row as usize]) }
    }

    pub fn num_constraints(&self) -> usize { self.constraint_paths.len() }
    pub fn num_inputs(&self) -> usize { self.input_names.len() }

    fn var_at(&self, path: &str) -> Variable {
        match self.named_objects.get(path) {
            Some(NamedObject::Var(v)) => *v,
            Some(e) => panic!("tried to access path `{}`, but `{:?}` exists there (not a variable)", path, e),
            _ => panic!("no variable exists at path: {}", path),
        }
    }

    /// test_cs.rs:270-282
    pub fn set(&mut self, path: &str, to: Scalar) {
        let (is_aux, idx) = match self.var_at(path).get_unchecked() {
            Index::Input(i) => (0, i),
            Index::Aux(i) => (1, i),
        };
        self.flush();
        let mut v = Vec::with_capacity(4);
        repr_to_limbs(&to, &mut v);
        let rc = unsafe { ffi::bp_cs_set(self.h, is_aux, idx as u64, v.as_ptr()) };
        self.check(rc);
    }

    /// test_cs.rs:311-323
    pub fn get(&self, path: &str) -> Scalar {
        let (is_aux, idx) = match self.var_at(path).get_unchecked() {
            Index::Input(i) => (0, i),
            Index::Aux(i) => (1, i),
        };
        self.flush();
        let mut v = [0u64; 4];
        let rc = unsafe { ffi::bp_cs_get(self.h, is_aux, idx as u64, v.as_mut_ptr()) };
        self.check(rc);
        let mut repr = Scalar::Repr::default();
        for (dst, limb) in repr.as_mut().chunks_exact_mut(8).zip(v.iter()) {
            dst.copy_from_slice(&limb.to_le_bytes());
        }
        Option::from(Scalar::from_repr(repr)).expect("device returned a canonical element")
    }
}

impl<Scalar: B200Field> Drop for B200ConstraintSystem<Scalar> {
    fn drop(&mut self) {
        unsafe { ffi::bp_cs_free(self.h) }
    }
}

impl<Scalar: B200Field> ConstraintSystem<Scalar> for B200ConstraintSystem<Scalar> {
    type Root = Self;

    fn new() -> Self {
        Self::new_on(0)
    }

    fn alloc<F, A, AR>(&mut self, annotation: A, f: F) -> Result<Variable, SynthesisError>
    where
        F: FnOnce() -> Result<Scalar, SynthesisError>,
        A: FnOnce() -> AR,
        AR: Into<String>,
    {
        let index = self.count[1] as usize;
        let path = compute_path(&self.current_namespace, &annotation().into());
        let value = f()?; // an Err leaves no variable behind (test_cs.rs:388)
        {
            let p = self.pend.get_mut();
            stage_value(&value, &mut p.bytes[1], &mut p.wide[1]);
        }
        self.count[1] += 1;
        self.aux_names.push(path.clone());
        let var = Variable::new_unchecked(Index::Aux(index));
        self.set_named_obj(path, NamedObject::Var(var));
        Ok(var)
    }

    fn alloc_input<F, A, AR>(&mut self, annotation: A, f: F) -> Result<Variable, SynthesisError>
    where
        F: FnOnce() -> Result<Scalar, SynthesisError>,
        A: FnOnce() -> AR,
        AR: Into<String>,
    {
        let index = self.count[0] as usize;
        let path = compute_path(&self.current_namespace, &annotation().into());
        let value = f()?;
        {
            let p = self.pend.get_mut();
            stage_value(&value, &mut p.bytes[0], &mut p.wide[0]);
        }
        self.count[0] += 1;
        self.input_names.push(path.clone());
        let var = Variable::new_unchecked(Index::Input(index));
        self.set_named_obj(path, NamedObject::Var(var));
        Ok(var)
    }

    fn enforce<A, AR, LA, LB, LC>(&mut self, annotation: A, a: LA, b: LB, c: LC)
    where
        A: FnOnce() -> AR,
        AR: Into<String>,
        LA: FnOnce(LinearCombination<Scalar>) -> LinearCombination<Scalar>,
        LB: FnOnce(LinearCombination<Scalar>) -> LinearCombination<Scalar>,
        LC: FnOnce(LinearCombination<Scalar>) -> LinearCombination<Scalar>,
    {
        let path = compute_path(&self.current_namespace, &annotation().into());
        let index = self.constraint_paths.len();
        self.set_named_obj(path.clone(), NamedObject::Constraint(index));
        let a = a(LinearCombination::zero());
        let b = b(LinearCombination::zero());
        let c = c(LinearCombination::zero());
        let staged = {
            let p = self.pend.get_mut();
            Self::push_lc(p, &a);
            Self::push_lc(p, &b);
            Self::push_lc(p, &c);
            p.cols.len()
        };
        self.constraint_paths.push(path);
        if staged >= (1 << 20) {
            self.flush();
        }
    }

    fn push_namespace<NR, N>(&mut self, name_fn: N)
    where
        NR: Into<String>,
        N: FnOnce() -> NR,
    {
        let name = name_fn().into();
        let path = compute_path(&self.current_namespace, &name);
        self.set_named_obj(path, NamedObject::Namespace);
        self.current_namespace.push(name);
    }

    fn pop_namespace(&mut self) {
        assert!(self.current_namespace.pop().is_some());
    }

    fn get_root(&mut self) -> &mut Self::Root {
        self
    }
}

#[cfg(test)]
mod tests {
    // Mirrors crates/bellpepper-core/src/util_cs/test_cs.rs:472-510 and the sha256 KATs; `impl B200Field for
    // blstrs::Scalar { const FIELD_ID: i32 = 0; }` lives in the test module of a build that has blstrs.
}
