//! Drop-in replacement for `src/transformer.rs` of samuel-vitorino/lm.rs that forwards the pub surface the bins
//! use to liblmrs_b200.so (C ABI in include/lmrs_b200.h).  NOT COMPILED in the authoring environment (no rustc);
//! kept deliberately mechanical.  Same names and signatures as the reference:
//!   Transformer::new(&Mmap) -> (Transformer, usize)          reference src/transformer.rs:134
//!   forward(&mut self, token, pos) -> &mut [f32]             :316
//!   get_embeddings(&self, &[u32]) -> Vec<f32>                :659
//!   fill_kv_cache(&mut self, &mut [f32], u32) -> u32         :672
//!   pub args: TransformerArgs {vocab_size, model_type, multimodal}   :57-74
//! The reference panics on malformed input; so does this shim (non-zero status -> panic! with the C message).
use memmap2::Mmap;
use std::ffi::CStr;
use std::marker::PhantomData;
use std::os::raw::{c_char, c_int};

use crate::functional::{u8_to_f32_slice, u8_to_i8_slice};
use crate::quantization::{QuantType, QuantizedTensor};

// ---- tensor views over the mapped file: `src/vision.rs:2` and `src/processor.rs:2` import these two helpers from this
// module (reference src/transformer.rs:16-48), so the multimodal build needs them here too.  They stay host-side: the
// vision tower still runs on the reference's CPU operators (or on the operator-level C ABI).

/// The next `n * size_each` f32 values of the file; advances `offset` past them.
pub fn init_param<'a>(data: &'a [u8], offset: &mut usize, n: u32, size_each: u32) -> &'a [f32] {
    let start = *offset;
    let end = start + n as usize * size_each as usize * std::mem::size_of::<f32>();
    *offset = end;
    u8_to_f32_slice(&data[start..end])
}

/// `n` consecutive quantized tensors of `size_each` elements: codes (one byte per element, two elements per byte for
/// Q4_0) immediately followed by one f32 scale per `gs` elements.  The returned slice lives for the rest of the process
/// like the reference's (it leaks the small vector of views, not the weights).
pub fn init_param_quant<'a>(data: &'a [u8], offset: &mut usize, n: u32, size_each: u32, gs: u32, q_type: QuantType) -> &'a [QuantizedTensor<'a>] {
    let code_bytes = if q_type == QuantType::Q4_0 { size_each as usize / 2 } else { size_each as usize };
    let scale_bytes = (size_each / gs) as usize * std::mem::size_of::<f32>();
    let views: Vec<QuantizedTensor<'a>> = (0..n)
        .map(|_| {
            let q = u8_to_i8_slice(&data[*offset..*offset + code_bytes]);
            *offset += code_bytes;
            let s = u8_to_f32_slice(&data[*offset..*offset + scale_bytes]);
            *offset += scale_bytes;
            QuantizedTensor { q, s }
        })
        .collect();
    Box::leak(views.into_boxed_slice())
}

#[derive(Debug, Copy, Clone, PartialEq)]
#[repr(u8)]
pub enum ModelType { GEMMA = 0, LLAMA = 1, PHI = 2 }

#[repr(C, packed)]
#[derive(Debug, Copy, Clone)]
pub struct TransformerArgs {
    dim: u32, hidden_dim: u32, n_layers: u32, n_heads: u32, head_size: u32, n_kv_heads: u32,
    pub vocab_size: u32, seq_len: u32, rms_norm_eps: f32, rope_theta: f32,
    q_type: QuantType, pub model_type: ModelType, group_size: u32, pub multimodal: bool,
}

#[repr(C)] struct Handle { _private: [u8; 0] }

#[link(name = "lmrs_b200")]
extern "C" {
    fn lmrs_b200_create(file: *const u8, len: usize, device: c_int, out: *mut *mut Handle, end_offset: *mut usize) -> c_int;
    fn lmrs_b200_create_multi(file: *const u8, len: usize, n_gpus: c_int, out: *mut *mut Handle, end_offset: *mut usize) -> c_int;
    fn lmrs_b200_destroy(m: *mut Handle);
    fn lmrs_b200_args(m: *const Handle, out: *mut TransformerArgs) -> c_int;
    fn lmrs_b200_forward(m: *mut Handle, token: u32, pos: u32, logits_host: *mut *mut f32) -> c_int;
    fn lmrs_b200_get_embeddings(m: *const Handle, tokens: *const u32, n: usize, out: *mut f32) -> c_int;
    fn lmrs_b200_fill_kv_cache(m: *mut Handle, emb: *mut f32, n_floats: usize, pos: u32, new_pos: *mut u32) -> c_int;
    fn lmrs_b200_last_error() -> *const c_char;
}

fn check(rc: c_int) {
    if rc != 0 {
        let msg = unsafe { CStr::from_ptr(lmrs_b200_last_error()) }.to_string_lossy().into_owned();
        panic!("{}", msg);
    }
}

pub struct Transformer<'a> {
    pub args: TransformerArgs,
    handle: *mut Handle,
    _file: PhantomData<&'a Mmap>,
}

impl<'a> Transformer<'a> {
    pub fn new(data: &'a Mmap) -> (Transformer<'a>, usize) {
        let mut h: *mut Handle = std::ptr::null_mut();
        let mut end: usize = 0;
        // LMRS_B200_GPUS=N: all N GPUs of this process behind the one handle (row-sharded, exchange inside the kernels)
        let n_gpus: c_int = std::env::var("LMRS_B200_GPUS").ok().and_then(|v| v.parse().ok()).unwrap_or(1);
        if n_gpus > 1 {
            check(unsafe { lmrs_b200_create_multi(data.as_ptr(), data.len(), n_gpus, &mut h, &mut end) });
        } else {
            check(unsafe { lmrs_b200_create(data.as_ptr(), data.len(), -1, &mut h, &mut end) });
        }
        let mut args = std::mem::MaybeUninit::<TransformerArgs>::uninit();
        check(unsafe { lmrs_b200_args(h, args.as_mut_ptr()) });
        let args = unsafe { args.assume_init() };
        println!("LMRS version: 4");
        println!("Model type: {:?}\n", { args.model_type });
        (Transformer { args, handle: h, _file: PhantomData }, end)
    }

    pub fn forward(&mut self, token: u32, pos: u32) -> &mut [f32] {
        let mut p: *mut f32 = std::ptr::null_mut();
        check(unsafe { lmrs_b200_forward(self.handle, token, pos, &mut p) });
        unsafe { std::slice::from_raw_parts_mut(p, self.args.vocab_size as usize) }
    }

    pub fn get_embeddings(&self, tokens: &[u32]) -> Vec<f32> {
        let dim = self.args.dim as usize;
        let mut out = vec![0.0f32; dim * tokens.len()];
        check(unsafe { lmrs_b200_get_embeddings(self.handle, tokens.as_ptr(), tokens.len(), out.as_mut_ptr()) });
        out
    }

    pub fn fill_kv_cache(&mut self, embeddings: &mut [f32], curr_pos: u32) -> u32 {
        let mut new_pos: u32 = 0;
        check(unsafe { lmrs_b200_fill_kv_cache(self.handle, embeddings.as_mut_ptr(), embeddings.len(), curr_pos, &mut new_pos) });
        new_pos
    }
}

impl<'a> Drop for Transformer<'a> {
    fn drop(&mut self) { unsafe { lmrs_b200_destroy(self.handle) } }
}
// The handle is used from one thread of control at a time (the reference takes &mut self); it may move threads.
unsafe impl<'a> Send for Transformer<'a> {}
