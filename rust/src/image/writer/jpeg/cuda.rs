//! FFI binding of the B200 encode path (C ABI: include/dmmt_cuda.h of dmmt-jpeg-encoder_b200).
//!
//! New module `crate::image::writer::jpeg::cuda` (declared in jpeg.rs by patches/jpeg.rs.patch, compiled with the
//! `cuda` feature).  It replaces the body of `JpegImageWriter::write_image` (src/image/writer/jpeg.rs:64-75):
//! `Transformer::transform` + `Encoder::encode` become ONE call, `dmmt_encode`, and the returned bytes are
//! `write_all`'d into the writer.  There is no CPU fallback behind these symbols: without a CUDA device
//! `dmmt_ctx_create` fails with DMMT_E_NODEVICE.
//!
//! Every `#[repr(C)]` type and every `extern "C"` signature below is checked against the header by
//! rust/check_layout.py (run by tests/test_host_logic.py), because this crate cannot be compiled where the
//! library is developed.
#![allow(dead_code)]

use std::ffi::CStr;
use std::os::raw::{c_char, c_int, c_void};

pub const DMMT_OK: c_int = 0;
pub const DMMT_E_INVALID: c_int = -1;
pub const DMMT_E_NODEVICE: c_int = -2;
pub const DMMT_E_CUDA: c_int = -3;
pub const DMMT_E_NCCL: c_int = -4;
pub const DMMT_E_NOMEM: c_int = -5;
pub const DMMT_E_OVERFLOW: c_int = -6;
pub const DMMT_E_SYMBOL: c_int = -7;
pub const DMMT_E_RANGE: c_int = -8;
pub const DMMT_E_WRITE: c_int = -9;
pub const DMMT_E_SIZE: c_int = -10;

/// dmmt_fmt
pub const DMMT_RGB_F32_NORM: c_int = 0;
pub const DMMT_RGB_U8: c_int = 1;
pub const DMMT_RGB_U16: c_int = 2;

/// status codes of dmmt_ppm_parse (> 0)
pub const DMMT_PPM_MISSING_TOKEN: c_int = 1;
pub const DMMT_PPM_BAD_TOKEN: c_int = 2;
pub const DMMT_PPM_INCOMPLETE_PIXEL: c_int = 3;
pub const DMMT_PPM_SIZE_MISMATCH: c_int = 4;
pub const DMMT_PPM_SAMPLE_ABOVE_MAX: c_int = 5;

#[repr(C)]
pub struct dmmt_ctx {
    _opaque: [u8; 0],
}

#[repr(C)]
pub struct dmmt_image {
    pub width: u16,
    pub height: u16,
    pub max_value: u16,
    pub fmt: c_int,
    pub pixels: *const c_void,
    pub pixels_on_device: c_int,
}

#[repr(C)]
pub struct dmmt_options {
    pub subsampling: u8,
    pub bits_per_channel: u8,
    pub qtable_preset: u8,
}

extern "C" {
    pub fn dmmt_device_count() -> c_int;
    pub fn dmmt_ctx_create(device: c_int, out: *mut *mut dmmt_ctx) -> c_int;
    pub fn dmmt_ctx_destroy(ctx: *mut dmmt_ctx);
    pub fn dmmt_encode(
        ctx: *mut dmmt_ctx,
        image: *const dmmt_image,
        options: *const dmmt_options,
        jpeg: *mut *mut u8,
        len: *mut usize,
    ) -> c_int;
    pub fn dmmt_encode_batch(
        ctxs: *const *mut dmmt_ctx,
        nctx: c_int,
        imgs: *const dmmt_image,
        n: c_int,
        options: *const dmmt_options,
        jpegs: *mut *mut u8,
        lens: *mut usize,
    ) -> c_int;
    pub fn dmmt_encode_sharded(
        ctxs: *const *mut dmmt_ctx,
        nctx: c_int,
        image: *const dmmt_image,
        options: *const dmmt_options,
        jpeg: *mut *mut u8,
        len: *mut usize,
    ) -> c_int;
    pub fn dmmt_free(p: *mut c_void);
    pub fn dmmt_strerror(code: c_int) -> *const c_char;
    pub fn dmmt_last_cuda_error() -> *const c_char;
    pub fn dmmt_ppm_parse(
        text: *const c_char,
        len: usize,
        threads: c_int,
        width: *mut u16,
        height: *mut u16,
        max_value: *mut u16,
        samples: *mut *mut u16,
        n_samples: *mut usize,
        detail: *mut c_int,
    ) -> c_int;
}

/// Text of a DMMT_E_* code (plus the CUDA runtime's message for DMMT_E_CUDA).
pub fn error_text(code: c_int) -> String {
    // SAFETY: both functions return pointers to NUL-terminated static / thread-local strings
    let base = unsafe { CStr::from_ptr(dmmt_strerror(code)) }.to_string_lossy().into_owned();
    if code == DMMT_E_CUDA {
        let detail = unsafe { CStr::from_ptr(dmmt_last_cuda_error()) }.to_string_lossy();
        return format!("{base}: {detail}");
    }
    base
}

/// One device + one stream; not thread-safe (one host thread per context), like the C object.
pub struct CudaContext {
    raw: *mut dmmt_ctx,
}

impl CudaContext {
    pub fn new(device: i32) -> Result<Self, c_int> {
        let mut raw = std::ptr::null_mut();
        // SAFETY: `raw` is a valid out-pointer
        let rc = unsafe { dmmt_ctx_create(device, &mut raw) };
        if rc != DMMT_OK {
            return Err(rc);
        }
        Ok(Self { raw })
    }

    /// `dmmt_encode` on host pixels.  `pixels` must hold width * height interleaved RGB triples of `fmt`.
    pub fn encode(
        &mut self,
        width: u16,
        height: u16,
        fmt: c_int,
        max_value: u16,
        pixels: *const c_void,
        options: &dmmt_options,
    ) -> Result<Vec<u8>, c_int> {
        let image = dmmt_image { width, height, max_value, fmt, pixels, pixels_on_device: 0 };
        let (mut ptr, mut len) = (std::ptr::null_mut::<u8>(), 0usize);
        // SAFETY: all pointers are valid for the call; the library allocates *ptr with malloc
        let rc = unsafe { dmmt_encode(self.raw, &image, options, &mut ptr, &mut len) };
        if rc != DMMT_OK {
            return Err(rc);
        }
        // SAFETY: on success ptr[0..len) is initialised and owned by us until dmmt_free
        let bytes = unsafe { std::slice::from_raw_parts(ptr, len) }.to_vec();
        unsafe { dmmt_free(ptr.cast()) };
        Ok(bytes)
    }
}

impl Drop for CudaContext {
    fn drop(&mut self) {
        // SAFETY: created by dmmt_ctx_create, destroyed once
        unsafe { dmmt_ctx_destroy(self.raw) }
    }
}
