//! UNVERIFIED (never compiled here).  Bit-identical Rust mirror of the device RNG (csrc/rtw_device.cuh):
//! Philox4x32-10, key = seed, counter = (pixel, sample, vertex, block); KATs in tests/test_oracle.py.
pub const M0: u32 = 0xD251_1F53;
pub const M1: u32 = 0xCD9E_8D57;
pub const W0: u32 = 0x9E37_79B9;
pub const W1: u32 = 0xBB67_AE85;

pub fn philox4x32_10(mut c: [u32; 4], mut k: [u32; 2]) -> [u32; 4] {
    for _ in 0..10 {
        let p0 = (M0 as u64) * (c[0] as u64);
        let p1 = (M1 as u64) * (c[2] as u64);
        c = [((p1 >> 32) as u32) ^ c[1] ^ k[0], p1 as u32, ((p0 >> 32) as u32) ^ c[3] ^ k[1], p0 as u32];
        k = [k[0].wrapping_add(W0), k[1].wrapping_add(W1)];
    }
    c
}

/// Uniform stream of one (pixel, sample, vertex); `exact` selects the W64 layout (53-bit, the f64 path),
/// otherwise W32 (24-bit, the FP32 path).
pub struct Stream { key: [u32; 2], pixel: u32, sample: u32, vertex: u32, k: u32, block: u32, buf: [u32; 4], exact: bool }

impl Stream {
    pub fn new(seed: u64, pixel: u32, sample: u32, vertex: u32, exact: bool) -> Self {
        Self { key: [seed as u32, (seed >> 32) as u32], pixel, sample, vertex, k: 0, block: u32::MAX, buf: [0; 4], exact }
    }
    fn word32(&mut self, idx: u32) -> u32 {
        let block = idx >> 2;
        if block != self.block {
            self.buf = philox4x32_10([self.pixel, self.sample, self.vertex, block], self.key);
            self.block = block;
        }
        self.buf[(idx & 3) as usize]
    }
    fn next64(&mut self) -> u64 { let k = self.k; self.k += 1; (self.word32(2 * k) as u64) | ((self.word32(2 * k + 1) as u64) << 32) }
    fn next32(&mut self) -> u32 { let k = self.k; self.k += 1; self.word32(k) }
    /// rand 0.8 `Standard`
    pub fn standard(&mut self) -> f64 {
        if self.exact { (self.next64() >> 11) as f64 * 2f64.powi(-53) } else { (self.next32() >> 8) as f64 * 2f64.powi(-24) }
    }
    /// rand 0.8 `Open01`
    pub fn open01(&mut self) -> f64 {
        if self.exact { (self.next64() >> 12) as f64 * 2f64.powi(-52) + 2f64.powi(-53) }
        else { (self.next32() >> 9) as f64 * 2f64.powi(-23) + 2f64.powi(-24) }
    }
    pub fn index(&mut self, n: u32) -> u32 {
        if self.exact { (((self.next64() as u128) * (n as u128)) >> 64) as u32 } else { (((self.next32() as u64) * (n as u64)) >> 32) as u32 }
    }
}
