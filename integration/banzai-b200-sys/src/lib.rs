//! Bindings to `libbanzai_b200.so` (C ABI: `include/banzai_b200.h`) and `encode`, a drop-in for
//! `banzai::encode(reader, BufWriter, level) -> io::Result<usize>` (banzai `lib/lib.rs:84-132`).
//!
//! Not compiled in the repository's environment (no Rust toolchain there); see INTEGRATION.md.
use std::ffi::CStr;
use std::io;
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)]
pub struct BnzCtx {
    _private: [u8; 0],
}
#[repr(C)]
pub struct BnzStream {
    _private: [u8; 0],
}

pub const BNZ_OK: c_int = 0;
pub const BNZ_EINVAL: c_int = 1;
pub const BNZ_ECUDA: c_int = 2;
pub const BNZ_ENOMEM: c_int = 3;
pub const BNZ_EINTERNAL: c_int = 4;
pub const BNZ_EIO: c_int = 5;
/// "verify" is on and a block failed the self-check (RLE1 decode, inverse BWT, cut chain, CRC)
pub const BNZ_EVERIFY: c_int = 6;

pub type BnzSinkFn = extern "C" fn(user: *mut c_void, data: *const u8, len: usize) -> c_int;

extern "C" {
    pub fn bnz_ctx_create(out: *mut *mut BnzCtx, n_gpus: c_int) -> c_int;
    pub fn bnz_ctx_destroy(ctx: *mut BnzCtx);
    pub fn bnz_strerror(code: c_int) -> *const c_char;
    pub fn bnz_last_error(ctx: *const BnzCtx) -> *const c_char;
    /// tunables, e.g. `bnz_ctx_set(ctx, c"verify".as_ptr(), 1)` (include/banzai_b200.h lists the keys)
    pub fn bnz_ctx_set(ctx: *mut BnzCtx, key: *const c_char, value: std::os::raw::c_long) -> c_int;
    pub fn bnz_encode(ctx: *mut BnzCtx, input: *const u8, in_len: usize, level: c_int,
                      out: *mut *mut u8, out_len: *mut usize, consumed: *mut usize) -> c_int;
    pub fn bnz_free(ctx: *mut BnzCtx, p: *mut u8);
    pub fn bnz_stream_open(ctx: *mut BnzCtx, level: c_int, sink: BnzSinkFn, user: *mut c_void,
                           out: *mut *mut BnzStream) -> c_int;
    pub fn bnz_stream_write(s: *mut BnzStream, data: *const u8, len: usize) -> c_int;
    pub fn bnz_stream_finish(s: *mut BnzStream, consumed: *mut usize) -> c_int;
    pub fn bnz_stream_close(s: *mut BnzStream);
}

fn text(p: *const c_char) -> String {
    if p.is_null() {
        String::new()
    } else {
        unsafe { CStr::from_ptr(p) }.to_string_lossy().into_owned()
    }
}

fn check(ctx: *mut BnzCtx, rc: c_int) -> io::Result<()> {
    if rc == BNZ_OK {
        return Ok(());
    }
    let detail = if ctx.is_null() { String::new() } else { text(unsafe { bnz_last_error(ctx) }) };
    Err(io::Error::new(io::ErrorKind::Other, format!("{}: {}", text(unsafe { bnz_strerror(rc) }), detail)))
}

/// What the sink hands the bytes to; remembers the first I/O error so that it can be returned
/// instead of the library's generic BNZ_EIO.
struct Sink<'a, W: io::Write> {
    writer: &'a mut io::BufWriter<W>,
    error: Option<io::Error>,
}

extern "C" fn sink<W: io::Write>(user: *mut c_void, data: *const u8, len: usize) -> c_int {
    let s = unsafe { &mut *(user as *mut Sink<W>) };
    match s.writer.write_all(unsafe { std::slice::from_raw_parts(data, len) }) {
        Ok(()) => 0,
        Err(e) => {
            s.error = Some(e);
            1
        }
    }
}

/// Same contract as `banzai::encode`: reads `reader` to its end, writes one `.bz2` stream to
/// `writer`, flushes it, returns the number of input bytes encoded; panics if `level` is not in
/// `1..=9` (banzai `lib/lib.rs:89`).  Streams: neither the input nor the output is held in memory.
pub fn encode<R, W>(mut reader: R, mut writer: io::BufWriter<W>, level: usize) -> io::Result<usize>
where
    R: io::BufRead,
    W: io::Write,
{
    assert!(1 <= level && level <= 9);
    let mut ctx: *mut BnzCtx = std::ptr::null_mut();
    check(std::ptr::null_mut(), unsafe { bnz_ctx_create(&mut ctx, 0) })?;      // 0 = all visible GPUs
    let result = (|| -> io::Result<usize> {
        let mut state = Sink { writer: &mut writer, error: None };
        let mut stream: *mut BnzStream = std::ptr::null_mut();
        check(ctx, unsafe {
            bnz_stream_open(ctx, level as c_int, sink::<W>, &mut state as *mut Sink<W> as *mut c_void, &mut stream)
        })?;
        let run = (|| -> io::Result<usize> {
            loop {
                let n = {
                    let buf = reader.fill_buf()?;                              // banzai's refill, rle.rs:62-79
                    if buf.is_empty() {
                        break;
                    }
                    check(ctx, unsafe { bnz_stream_write(stream, buf.as_ptr(), buf.len()) })?;
                    buf.len()
                };
                reader.consume(n);
            }
            let mut consumed = 0usize;
            check(ctx, unsafe { bnz_stream_finish(stream, &mut consumed) })?;  // footer + padding
            Ok(consumed)
        })();
        unsafe { bnz_stream_close(stream) };
        match (run, state.error.take()) {
            (_, Some(e)) => Err(e),                                            // the writer's own error
            (r, None) => r,
        }
    })();
    unsafe { bnz_ctx_destroy(ctx) };
    let consumed = result?;
    io::Write::flush(&mut writer)?;                                            // out.rs:22-28 close()
    Ok(consumed)
}
