// protocol_decoder/src/gpu_ffi.rs — the C ABI of libppd_b200.so (include/ppd_b200.h) as Rust sees it.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)] pub struct ppd_ctx { _private: [u8; 0] }

/// `done` of ppd_blocks_decode_stream: called once per block from one of the library's host threads, in completion
/// order; `out` (null for a failed block) is the callee's to release with ppd_free.
pub type ppd_block_done_fn = extern "C" fn(user: *mut c_void, index: usize, status: c_int, out: *mut u8, out_len: usize);

extern "C" {
    pub fn ppd_ctx_create(device: c_int, out: *mut *mut ppd_ctx) -> c_int;
    pub fn ppd_ctx_destroy(ctx: *mut ppd_ctx);
    pub fn ppd_last_error(ctx: *const ppd_ctx) -> *const c_char;
    pub fn ppd_free(p: *mut c_void);
    pub fn ppd_alloc_pinned(n: usize) -> *mut c_void;   // page-locked; release with ppd_free
    pub fn ppd_block_decode(ctx: *mut ppd_ctx, flat: *const u8, len: usize,
                            out: *mut *mut u8, out_len: *mut usize) -> c_int;
    pub fn ppd_blocks_decode_batch(ctx: *mut ppd_ctx, flats: *const *const u8, lens: *const usize, n: usize,
                                   outs: *mut *mut u8, out_lens: *mut usize, statuses: *mut c_int) -> c_int;
    pub fn ppd_blocks_decode_stream(ctx: *mut ppd_ctx, flats: *const *const u8, lens: *const usize, n: usize,
                                    done: ppd_block_done_fn, user: *mut c_void) -> c_int;
    pub fn ppd_compact_decode(ctx: *mut ppd_ctx, witness: *const u8, len: usize,
                              out: *mut *mut u8, out_len: *mut usize) -> c_int;
    pub fn ppd_keccak256_batch(ctx: *mut ppd_ctx, data: *const u8, offsets: *const u64, n: usize, out32n: *mut u8) -> c_int;
    // a DirectPreImage payload (FlatBlock kind 2) as the TrieCompact witness of the same tries; host only, no context
    pub fn ppd_direct_to_compact(direct: *const u8, len: usize, out: *mut *mut u8, out_len: *mut usize) -> c_int;
}

/// One context per (thread, device); the library's contexts are not thread-safe.
pub struct GpuDecoder { ctx: *mut ppd_ctx }
unsafe impl Send for GpuDecoder {}

impl GpuDecoder {
    pub fn new(device: i32) -> Result<Self, i32> {
        let mut ctx = std::ptr::null_mut();
        match unsafe { ppd_ctx_create(device, &mut ctx) } { 0 => Ok(Self { ctx }), rc => Err(rc) } // 100 = no usable CUDA device; there is no CPU fallback
    }
    fn last_error(&self) -> String {
        unsafe { std::ffi::CStr::from_ptr(ppd_last_error(self.ctx)) }.to_string_lossy().into_owned()
    }
    /// FlatBlock in, IrDump out (include/ppd_flat.h).
    pub fn block_decode(&mut self, flat: &[u8]) -> Result<Vec<u8>, (i32, String)> {
        let (mut out, mut n) = (std::ptr::null_mut(), 0usize);
        let rc = unsafe { ppd_block_decode(self.ctx, flat.as_ptr(), flat.len(), &mut out, &mut n) };
        if rc != 0 { return Err((rc, self.last_error())); }
        let v = unsafe { std::slice::from_raw_parts(out, n) }.to_vec();
        unsafe { ppd_free(out as *mut c_void) };
        Ok(v)
    }
    /// Many independent blocks: `sink(i, Ok(IrDump) | Err(status))` is called as block i finishes (completion order),
    /// possibly from several of the library's threads at once; the pipeline stays full across the whole slice.
    pub fn blocks_decode_stream<S: Fn(usize, Result<&[u8], i32>) + Sync>(&mut self, flats: &[&[u8]], sink: S) -> Result<(), (i32, String)> {
        extern "C" fn tramp<S: Fn(usize, Result<&[u8], i32>) + Sync>(user: *mut c_void, i: usize, status: c_int, out: *mut u8, n: usize) {
            let sink = unsafe { &*(user as *const S) };
            if status == 0 && !out.is_null() {
                sink(i, Ok(unsafe { std::slice::from_raw_parts(out, n) }));
            } else {
                sink(i, Err(status));
            }
            if !out.is_null() { unsafe { ppd_free(out as *mut c_void) }; }
        }
        let ptrs: Vec<*const u8> = flats.iter().map(|f| f.as_ptr()).collect();
        let lens: Vec<usize> = flats.iter().map(|f| f.len()).collect();
        let rc = unsafe { ppd_blocks_decode_stream(self.ctx, ptrs.as_ptr(), lens.as_ptr(), flats.len(), tramp::<S>, &sink as *const S as *mut c_void) };
        if rc != 0 { return Err((rc, self.last_error())); }
        Ok(())
    }
}
impl Drop for GpuDecoder { fn drop(&mut self) { unsafe { ppd_ctx_destroy(self.ctx) } } }
