//! b200tfhe-sys: thin `extern "C"` binding of libb200tfhe.so (include/b200tfhe.h) plus the safe
//! wrapper `B200Engine` that `tfhe::shortint::ServerKey` holds when the `b200` feature is on.
//!
//! Source only: this image has no cargo/rustc, so the crate is reviewed by signature against
//! include/b200tfhe.h; tests/test_abi.py checks that every `extern` name below is exported by the
//! built library.  See INTEGRATION.md for the patch to the reference that uses it.
#![allow(non_camel_case_types)]

use std::ffi::c_void;
use std::os::raw::{c_char, c_int};
use std::ptr;

#[repr(C)]
pub struct b200tfhe_ctx {
    _private: [u8; 0],
}
#[repr(C)]
pub struct b200tfhe_program {
    _private: [u8; 0],
}
#[repr(C)]
pub struct b200tfhe_boolean_ctx {
    _private: [u8; 0],
}

/// `ClassicPBSParameters` (tfhe/src/shortint/parameters/mod.rs:62-76) flattened.
#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct b200tfhe_params {
    pub lwe_dimension: u32,
    pub glwe_dimension: u32,
    pub polynomial_size: u32,
    pub pbs_base_log: u32,
    pub pbs_level: u32,
    pub ks_base_log: u32,
    pub ks_level: u32,
    pub message_modulus: u32,
    pub carry_modulus: u32,
}

/// Where the key material sits inside a serialised server key (b200tfhe_parse_server_key).
#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct b200tfhe_key_view {
    pub ksk_offset: u64,
    pub ksk_len: u64,
    pub bsk_offset: u64,
    pub bsk_len: u64,
    pub bsk_poly_stride_bytes: u64,
    pub bsk_is_fourier: u32,
    pub pbs_order: u32,
    pub max_degree: u64,
    pub max_noise_level: u64,
}

/// A caller-built level schedule (b200tfhe_program_create_from_circuit).
#[repr(C)]
pub struct b200tfhe_circuit_desc {
    pub n_inputs: usize,
    pub n_nodes: usize,
    pub n_luts: usize,
    pub n_outputs: usize,
    pub node_term_begin: *const u32,
    pub term_block: *const i32,
    pub term_coeff: *const i64,
    pub node_plaintext: *const u64,
    pub node_lut: *const i32,
    pub luts: *const u64,
    pub outputs: *const i32,
}

#[link(name = "b200tfhe")]
extern "C" {
    pub fn b200tfhe_ctx_create(params: *const b200tfhe_params, device: c_int, out: *mut *mut b200tfhe_ctx) -> c_int;
    pub fn b200tfhe_ctx_create_multi(params: *const b200tfhe_params, devices: *const c_int, n_devices: c_int, out: *mut *mut b200tfhe_ctx) -> c_int;
    pub fn b200tfhe_ctx_device_count(ctx: *const b200tfhe_ctx, n_devices: *mut c_int) -> c_int;
    pub fn b200tfhe_ctx_destroy(ctx: *mut b200tfhe_ctx) -> c_int;
    pub fn b200tfhe_last_error(ctx: *const b200tfhe_ctx, buf: *mut c_char, buf_len: usize) -> c_int;
    pub fn b200tfhe_last_global_error(buf: *mut c_char, buf_len: usize) -> c_int;
    pub fn b200tfhe_load_ksk(ctx: *mut b200tfhe_ctx, ksk: *const u64, n_u64: usize) -> c_int;
    pub fn b200tfhe_load_bsk_standard(ctx: *mut b200tfhe_ctx, bsk: *const u64, n_u64: usize) -> c_int;
    pub fn b200tfhe_key_arena(ctx: *mut b200tfhe_ctx, device_ptr: *mut *mut c_void, bytes: *mut usize) -> c_int;
    pub fn b200tfhe_keys_adopt(ctx: *mut b200tfhe_ctx) -> c_int;
    pub fn b200tfhe_register_lut(ctx: *mut b200tfhe_ctx, glwe_acc: *const u64, id: *mut u32) -> c_int;
    pub fn b200tfhe_register_lut_from_table(ctx: *mut b200tfhe_ctx, table: *const u64, table_len: usize, id: *mut u32) -> c_int;
    pub fn b200tfhe_keyswitch_batch(ctx: *mut b200tfhe_ctx, input: *const u64, out: *mut u64, batch: usize) -> c_int;
    pub fn b200tfhe_pbs_batch(ctx: *mut b200tfhe_ctx, input: *const u64, lut_id: *const u32, out: *mut u64, batch: usize) -> c_int;
    pub fn b200tfhe_ks_pbs_batch(ctx: *mut b200tfhe_ctx, input: *const u64, lut_id: *const u32, out: *mut u64, batch: usize) -> c_int;
    pub fn b200tfhe_pbs_ks_batch(ctx: *mut b200tfhe_ctx, input: *const u64, lut_id: *const u32, out: *mut u64, batch: usize) -> c_int;
    pub fn b200tfhe_pbs_ks_batch_device(ctx: *mut b200tfhe_ctx, d_in: *const u64, d_lut_id: *const u32, d_out: *mut u64, batch: usize) -> c_int;
    pub fn b200tfhe_keyswitch_batch_device(ctx: *mut b200tfhe_ctx, d_in: *const u64, d_out: *mut u64, batch: usize) -> c_int;
    pub fn b200tfhe_pbs_batch_device(ctx: *mut b200tfhe_ctx, d_in: *const u64, d_lut_id: *const u32, d_out: *mut u64, batch: usize) -> c_int;
    pub fn b200tfhe_ks_pbs_batch_device(ctx: *mut b200tfhe_ctx, d_in: *const u64, d_lut_id: *const u32, d_out: *mut u64, batch: usize) -> c_int;
    pub fn b200tfhe_lwe_linear_batch_device(
        ctx: *mut b200tfhe_ctx, d_x: *const u64, d_y: *const u64, d_ia: *const i32, d_ib: *const i32,
        d_ca: *const i64, d_cb: *const i64, d_pt: *const u64, d_out: *mut u64, batch: usize, lwe_size: usize,
    ) -> c_int;
    pub fn b200tfhe_sync(ctx: *mut b200tfhe_ctx) -> c_int;
    pub fn b200tfhe_stream(ctx: *mut b200tfhe_ctx, stream: *mut *mut c_void) -> c_int;
    pub fn b200tfhe_set_profiling(ctx: *mut b200tfhe_ctx, enabled: c_int) -> c_int;
    pub fn b200tfhe_get_kernel_times(
        ctx: *mut b200tfhe_ctx, ks_ms: *mut f64, ks_launches: *mut u64, pbs_ms: *mut f64, pbs_launches: *mut u64, reset: c_int,
    ) -> c_int;
    pub fn b200tfhe_kernel_launch_count(ctx: *mut b200tfhe_ctx, count: *mut u64) -> c_int;
    pub fn b200tfhe_ks_pbs_batch_device_multi(
        ctx: *mut b200tfhe_ctx, d_in: *const *const u64, d_lut_id: *const *const u32, d_out: *const *mut u64, batch: *const usize,
    ) -> c_int;
    pub fn b200tfhe_parse_server_key(bytes: *const u8, n_bytes: usize, params: *mut b200tfhe_params, view: *mut b200tfhe_key_view) -> c_int;
    pub fn b200tfhe_load_server_key_bytes(ctx: *mut b200tfhe_ctx, bytes: *const u8, n_bytes: usize) -> c_int;
    pub fn b200tfhe_program_create_from_circuit(
        ctx: *mut b200tfhe_ctx, desc: *const b200tfhe_circuit_desc, out: *mut *mut b200tfhe_program,
    ) -> c_int;
    pub fn b200tfhe_boolean_ctx_create(params: *const b200tfhe_params, keyswitch_first: c_int, device: c_int, out: *mut *mut b200tfhe_boolean_ctx) -> c_int;
    pub fn b200tfhe_boolean_ctx_destroy(ctx: *mut b200tfhe_boolean_ctx) -> c_int;
    pub fn b200tfhe_boolean_last_error(ctx: *const b200tfhe_boolean_ctx, buf: *mut c_char, buf_len: usize) -> c_int;
    pub fn b200tfhe_boolean_load_ksk(ctx: *mut b200tfhe_boolean_ctx, ksk: *const u32, n_u32: usize) -> c_int;
    pub fn b200tfhe_boolean_load_bsk_standard(ctx: *mut b200tfhe_boolean_ctx, bsk: *const u32, n_u32: usize) -> c_int;
    pub fn b200tfhe_boolean_gate_batch(
        ctx: *mut b200tfhe_boolean_ctx, gate: c_int, a: *const u32, b: *const u32, out: *mut u32, batch: usize,
    ) -> c_int;
    pub fn b200tfhe_debug_pbs_steps(
        ctx: *mut b200tfhe_ctx, in_small: *const u64, lut_id: *const u32, out: *mut u64, batch: usize, steps: u32,
    ) -> c_int;
    pub fn b200tfhe_program_create(
        ctx: *mut b200tfhe_ctx, op: *const c_char, shape: *const u64, n_shape: usize, out: *mut *mut b200tfhe_program,
    ) -> c_int;
    pub fn b200tfhe_program_info(prog: *const b200tfhe_program, info: *mut u64) -> c_int;
    pub fn b200tfhe_program_run(prog: *mut b200tfhe_program, input: *const u64, out: *mut u64) -> c_int;
    pub fn b200tfhe_program_run_device(prog: *mut b200tfhe_program, d_in: *const u64, d_out: *mut u64) -> c_int;
    pub fn b200tfhe_program_destroy(prog: *mut b200tfhe_program) -> c_int;
    pub fn b200tfhe_debug_negacyclic_mul(ctx: *mut b200tfhe_ctx, a_int: *const u64, b_torus: *const u64, out: *mut u64, count: usize) -> c_int;
    pub fn b200tfhe_debug_from_torus(ctx: *mut b200tfhe_ctx, x: *const f64, out_fp: *mut u64, out_cvt: *mut u64, n: usize) -> c_int;
}

#[derive(Debug)]
pub struct B200Error(pub String);

/// Owns one device context; `Send + Sync` because the C library serialises calls per context.
pub struct B200Engine {
    ctx: *mut b200tfhe_ctx,
    big_size: usize,
}
unsafe impl Send for B200Engine {}
unsafe impl Sync for B200Engine {}

impl B200Engine {
    fn err(ctx: *const b200tfhe_ctx) -> B200Error {
        let mut buf = vec![0u8; 512];
        unsafe { b200tfhe_last_error(ctx, buf.as_mut_ptr() as *mut c_char, buf.len()) };
        let n = buf.iter().position(|&b| b == 0).unwrap_or(buf.len());
        B200Error(String::from_utf8_lossy(&buf[..n]).into_owned())
    }

    /// `ksk` = `key_switching_key.as_ref()`, `bsk_standard` = `bootstrap_key.as_ref()` taken where
    /// the reference still holds the standard-domain key (shortint/engine/server_side.rs:63-86).
    /// `devices`: the GPUs of this box the engine may use (one context over all of them: every batch
    /// call is cut into one contiguous shard per GPU inside the library).
    pub fn new(params: b200tfhe_params, devices: &[i32], ksk: &[u64], bsk_standard: &[u64]) -> Result<Self, B200Error> {
        let mut ctx = ptr::null_mut();
        if unsafe { b200tfhe_ctx_create_multi(&params, devices.as_ptr(), devices.len() as c_int, &mut ctx) } != 0 {
            return Err(Self::err(ptr::null()));
        }
        let e = B200Engine { ctx, big_size: (params.glwe_dimension * params.polynomial_size + 1) as usize };
        if unsafe { b200tfhe_load_ksk(ctx, ksk.as_ptr(), ksk.len()) } != 0
            || unsafe { b200tfhe_load_bsk_standard(ctx, bsk_standard.as_ptr(), bsk_standard.len()) } != 0
        {
            return Err(Self::err(ctx));
        }
        Ok(e)
    }

    /// `LookupTableOwned::acc.as_ref()` -> device-resident table id (content addressed).
    pub fn register_lut(&self, glwe_acc: &[u64]) -> Result<u32, B200Error> {
        let mut id = 0u32;
        match unsafe { b200tfhe_register_lut(self.ctx, glwe_acc.as_ptr(), &mut id) } {
            0 => Ok(id),
            _ => Err(Self::err(self.ctx)),
        }
    }

    /// Batched `keyswitch_programmable_bootstrap_assign` on flat LWE buffers (in place).  `cts` may be an
    /// ordinary (pageable) `Vec<u64>`: the library stages it wave by wave through its own pinned slabs, so
    /// the copies still run under the kernels (measured within a few % of a pinned buffer).
    pub fn ks_pbs_batch(&self, cts: &mut [u64], lut_ids: &[u32]) -> Result<(), B200Error> {
        let batch = lut_ids.len();
        assert_eq!(cts.len(), batch * self.big_size);
        match unsafe { b200tfhe_ks_pbs_batch(self.ctx, cts.as_ptr(), lut_ids.as_ptr(), cts.as_mut_ptr(), batch) } {
            0 => Ok(()),
            _ => Err(Self::err(self.ctx)),
        }
    }
}

impl Drop for B200Engine {
    fn drop(&mut self) {
        unsafe { b200tfhe_ctx_destroy(self.ctx) };
    }
}
