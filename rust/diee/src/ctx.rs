//! One `diee_ctx` per GPU.  Calls on one context are serialised by `&mut`/interior locking of the caller; different
//! contexts may be used from different threads (`include/diee.h`, conventions).
use diee_sys as sys;
use std::ffi::CStr;
use std::fmt;
use std::ptr;

#[derive(Debug, Clone)]
pub struct DieeError {
    pub code: i32,
    pub message: String,
}

impl fmt::Display for DieeError {
    fn fmt(&self, f: &mut fmt::Formatter<'_>) -> fmt::Result {
        write!(f, "diee error {}: {}", self.code, self.message)
    }
}
impl std::error::Error for DieeError {}

pub struct Ctx {
    pub(crate) raw: *mut sys::diee_ctx,
}

// the library serialises nothing itself: a context must not be used from two threads at once
unsafe impl Send for Ctx {}

impl Ctx {
    /// `diee_ctx_create`: fails when no CUDA device is usable -- there is no CPU fallback.
    pub fn new(device: i32) -> Result<Self, DieeError> {
        let mut raw: *mut sys::diee_ctx = ptr::null_mut();
        let rc = unsafe { sys::diee_ctx_create(device, &mut raw) };
        if rc != sys::DIEE_OK {
            return Err(DieeError { code: rc, message: "diee_ctx_create failed: no usable CUDA device".into() });
        }
        Ok(Ctx { raw })
    }

    pub fn last_error(&self) -> String {
        unsafe { CStr::from_ptr(sys::diee_last_error(self.raw)).to_string_lossy().into_owned() }
    }

    pub(crate) fn check(&self, rc: i32) -> Result<(), DieeError> {
        if rc == sys::DIEE_OK { Ok(()) } else { Err(DieeError { code: rc, message: self.last_error() }) }
    }

    pub fn sync(&self) -> Result<(), DieeError> {
        self.check(unsafe { sys::diee_sync(self.raw) })
    }

    pub fn launch_count(&self) -> i64 {
        unsafe { sys::diee_launch_count(self.raw) }
    }

    // ---- multi-GPU exchange: one rank per context ----
    pub fn comm_unique_id() -> Result<[u8; sys::DIEE_COMM_ID_BYTES], DieeError> {
        let mut id = [0u8; sys::DIEE_COMM_ID_BYTES];
        let rc = unsafe { sys::diee_comm_unique_id(id.as_mut_ptr()) };
        if rc != sys::DIEE_OK {
            return Err(DieeError { code: rc, message: "NCCL is not available".into() });
        }
        Ok(id)
    }

    pub fn comm_init(&self, nranks: i32, rank: i32, id: &[u8; sys::DIEE_COMM_ID_BYTES]) -> Result<(), DieeError> {
        self.check(unsafe { sys::diee_comm_init(self.raw, nranks, rank, id.as_ptr()) })
    }

    pub fn comm_destroy(&self) -> Result<(), DieeError> {
        self.check(unsafe { sys::diee_comm_destroy(self.raw) })
    }
}

impl Drop for Ctx {
    fn drop(&mut self) {
        if !self.raw.is_null() {
            unsafe { sys::diee_ctx_destroy(self.raw) };
        }
    }
}

/// One block of the injected Philox4x32-10 stream (`include/diee.h`): what replaces `rand::thread_rng()`.
pub fn philox(seed: u64, c0: u32, c1: u32, c2: u32, c3: u32) -> [u32; 4] {
    let mut out = [0u32; 4];
    unsafe { sys::diee_philox(seed, c0, c1, c2, c3, out.as_mut_ptr()) };
    out
}

/// die face from a stream word: `1 + ((w * 6) >> 32)`
pub fn die_of(w: u32) -> u8 {
    (1 + (((w as u64) * 6) >> 32)) as u8
}

/// uniform index in `[0, n)` from a stream word
pub fn index_of(w: u32, n: u32) -> u32 {
    (((w as u64) * (n as u64)) >> 32) as u32
}
