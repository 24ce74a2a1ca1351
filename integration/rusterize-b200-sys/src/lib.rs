//! Raw bindings to `librz_b200.so` — one item per declaration of `include/rz_b200.h`, same order — and a thin safe
//! layer (`Geoms`, `Sparse`, `flatten`) on top.  The struct layouts are checked at start-up against the library's
//! own `rz_abi_layout()` (see `check_abi`).
//!
//! Reference seam this replaces: `geoms.rasterize::<DenseArray<N> | SparseArray<N>>(ctx)`
//! (rust/src/rasterize.rs:54-62).  `integration/rasterize_b200.rs` holds the `ArrayBuilder` bodies that call into
//! this crate from inside the `rusterize` core crate.
#![allow(non_camel_case_types)]

use std::ffi::{c_char, c_int, c_void, CStr};

pub mod flatten;
pub use flatten::{flatten, GeomSoa};

pub const RZ_OK: c_int = 0;
pub const RZ_VALUE_ERROR: c_int = 1; // RusterizeError::ValueError
pub const RZ_RUNTIME_ERROR: c_int = 2; // RusterizeError::RuntimeError

pub const RZ_PART_POLYGON: u8 = 0;
pub const RZ_PART_LINE: u8 = 1;
pub const RZ_PART_POINT: u8 = 2;

pub const RZ_FLAG_OUT_ON_DEVICE: u32 = 1;
pub const RZ_FLAG_FORCE_H2D: u32 = 2;
pub const RZ_FLAG_SYNC_STAGES: u32 = 4;
pub const RZ_FLAG_NO_TILE_ENGINE: u32 = 8;
pub const RZ_FLAG_FORCE_TILE_ENGINE: u32 = 16;
pub const RZ_FLAG_INPUTS_ON_DEVICE: u32 = 128;
pub const RZ_FLAG_OUT_ROW_COL_BAND: u32 = 256; // R's (row, col, band) column-major layout

/// `rz_dtype`, in the order of python/src/rusterize.rs:171-182.
pub trait RzDtype: Copy {
    const CODE: i32;
}
macro_rules! dtype { ($($t:ty => $c:expr),*) => { $(impl RzDtype for $t { const CODE: i32 = $c; })* } }
dtype!(u8 => 0, u16 => 1, u32 => 2, u64 => 3, i8 => 4, i16 => 5, i32 => 6, i64 => 7, f32 => 8, f64 => 9);

/// `rz_pixel_fn` (rust/src/rasterization/pixel_functions.rs:8-16).
#[repr(i32)]
#[derive(Clone, Copy, Debug, PartialEq, Eq)]
pub enum RzPixelFn {
    Sum = 0,
    First = 1,
    Last = 2,
    Min = 3,
    Max = 4,
    Count = 5,
    Any = 6,
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct RzRasterInfo {
    pub nrows: u64,
    pub ncols: u64,
    pub xmin: f64,
    pub ymin: f64,
    pub xmax: f64,
    pub ymax: f64,
    pub xres: f64,
    pub yres: f64,
    pub epsg: i32, // < 0: None
    pub _pad: i32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct RzRawRasterInfo {
    pub has_shape: i32,
    pub has_extent: i32,
    pub has_resolution: i32,
    pub tap: i32,
    pub nrows: u64,
    pub ncols: u64,
    pub extent: [f64; 4], // xmin, ymin, xmax, ymax
    pub xres: f64,
    pub yres: f64,
    pub epsg: i32,
    pub _pad: i32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct RzGeomSoa {
    pub n_geoms: u64,
    pub n_parts: u64,
    pub n_seqs: u64,
    pub n_coords: u64,
    pub geom_part_off: *const u64,
    pub part_kind: *const u8,
    pub part_seq_off: *const u64,
    pub seq_coord_off: *const u64,
    pub x: *const f64,
    pub y: *const f64,
}

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct RzContext {
    pub raster_info: RzRasterInfo,
    pub dtype: i32,
    pub pixel_fn: i32,
    pub field: *const c_void,
    pub field_is_scalar: i32,
    pub all_touched: i32,
    pub field_len: u64,
    pub field_valid: *const u8,
    pub band_of_geom: *const i32,
    pub by_len: u64,
    pub n_bands: i32,
    pub device: i32,
    pub background: *const c_void,
    pub row_begin: u64,
    pub row_end: u64,
    pub stream: *mut c_void,
    pub flags: u32,
    pub tile_bytes: u32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct RzStats {
    pub n_parts: u64,
    pub n_poly_vertices: u64,
    pub n_line_vertices: u64,
    pub n_points: u64,
    pub n_records: u64,
    pub n_crossings: u64,
    pub n_tasks: u64,
    pub key_bits: u32,
    pub sort_passes: u32,
    pub tile_width: u32,
    pub n_windows: u32,
    pub h2d_ms: f32,
    pub count_ms: f32,
    pub emit_ms: f32,
    pub sort_ms: f32,
    pub index_ms: f32,
    pub fill_ms: f32,
    pub d2h_ms: f32,
    pub total_ms: f32,
    pub h2d_bytes: u64,
    pub d2h_bytes: u64,
    pub out_bytes: u64,
    pub kernel_launches: u32,
    pub engine: u32,
    pub n_mask_words: u64,
    pub host_syncs: u32,
    pub plan_cached: u32,
    pub wall_ms: f32,
    pub shard_ms: f32,
}

pub enum rz_geoms {}
pub enum rz_sparse {}

unsafe extern "C" {
    pub fn rz_geoms_from_wkb(bufs: *const *const u8, lens: *const u64, n: u64, err: *mut c_char, errlen: usize) -> *mut rz_geoms;
    pub fn rz_geoms_from_wkt(strs: *const *const c_char, n: u64, err: *mut c_char, errlen: usize) -> *mut rz_geoms;
    pub fn rz_geoms_from_soa(soa: *const RzGeomSoa, err: *mut c_char, errlen: usize) -> *mut rz_geoms;
    pub fn rz_geoms_from_soa_to(soa: *const RzGeomSoa, device: c_int, err: *mut c_char, errlen: usize) -> *mut rz_geoms;
    pub fn rz_geoms_len(g: *const rz_geoms) -> u64;
    pub fn rz_geoms_n_parts(g: *const rz_geoms) -> u64;
    pub fn rz_geoms_n_coords(g: *const rz_geoms) -> u64;
    pub fn rz_geoms_bounds(g: *const rz_geoms, out_xmin_ymin_xmax_ymax: *mut f64) -> c_int;
    pub fn rz_geoms_upload(g: *mut rz_geoms, device: c_int, err: *mut c_char, errlen: usize) -> c_int;
    pub fn rz_geoms_evict(g: *mut rz_geoms);
    pub fn rz_geoms_free(g: *mut rz_geoms);
    pub fn rz_geoms_row_shard(g: *const rz_geoms, ri: *const RzRasterInfo, row_begin: u64, row_end: u64, all_touched: c_int,
                              err: *mut c_char, errlen: usize) -> *mut rz_geoms;
    pub fn rz_geoms_from_soa_rows(soa: *const RzGeomSoa, ri: *const RzRasterInfo, row_begin: u64, row_end: u64,
                                  all_touched: c_int, err: *mut c_char, errlen: usize) -> *mut rz_geoms;
    pub fn rz_geoms_part_kind(g: *const rz_geoms) -> *const u8;
    pub fn rz_geoms_part_geom(g: *const rz_geoms) -> *const u64;
    pub fn rz_geoms_pool_len(g: *const rz_geoms, kind: c_int) -> u64;
    pub fn rz_geoms_pool_x(g: *const rz_geoms, kind: c_int) -> *const f64;
    pub fn rz_geoms_pool_y(g: *const rz_geoms, kind: c_int) -> *const f64;
    pub fn rz_geoms_pool_tag(g: *const rz_geoms, kind: c_int) -> *const u32;
    pub fn rz_raster_info_build(raw: *const RzRawRasterInfo, g: *const rz_geoms, out: *mut RzRasterInfo, err: *mut c_char,
                                errlen: usize) -> c_int;
    pub fn rz_group_keys(keys: *const *const c_char, n: u64, band_of_geom: *mut i32, band_first: *mut u64) -> i64;
    pub fn rz_rasterize_dense(g: *mut rz_geoms, ctx: *const RzContext, out: *mut c_void, stats: *mut RzStats, err: *mut c_char,
                              errlen: usize) -> c_int;
    pub fn rz_rasterize_sparse(g: *mut rz_geoms, ctx: *const RzContext, out: *mut *mut rz_sparse, stats: *mut RzStats,
                               err: *mut c_char, errlen: usize) -> c_int;
    pub fn rz_rasterize_dense_multi(g: *mut rz_geoms, ctx: *const RzContext, devices: *const i32, n_devices: i32,
                                    out: *mut c_void, stats: *mut RzStats, per_device: *mut RzStats, err: *mut c_char,
                                    errlen: usize) -> c_int;
    pub fn rz_rasterize_sparse_multi(g: *mut rz_geoms, ctx: *const RzContext, devices: *const i32, n_devices: i32,
                                     out: *mut *mut rz_sparse, stats: *mut RzStats, per_device: *mut RzStats,
                                     err: *mut c_char, errlen: usize) -> c_int;
    pub fn rz_rasterize_dense_soa(soa: *const RzGeomSoa, ctx: *const RzContext, devices: *const i32, n_devices: i32,
                                  out: *mut c_void, stats: *mut RzStats, per_device: *mut RzStats, err: *mut c_char,
                                  errlen: usize) -> c_int;
    pub fn rz_sparse_len(s: *const rz_sparse) -> u64;
    pub fn rz_sparse_n_bands(s: *const rz_sparse) -> u64;
    pub fn rz_sparse_rows(s: *const rz_sparse) -> *const u64;
    pub fn rz_sparse_cols(s: *const rz_sparse) -> *const u64;
    pub fn rz_sparse_data(s: *const rz_sparse) -> *const c_void;
    pub fn rz_sparse_counts(s: *const rz_sparse) -> *const u64;
    pub fn rz_sparse_free(s: *mut rz_sparse);
    pub fn rz_sparse_build_array(ctx: *const RzContext, n_bands: u64, counts: *const u64, rows: *const u64, cols: *const u64,
                                 data: *const c_void, out: *mut c_void, stats: *mut RzStats, err: *mut c_char,
                                 errlen: usize) -> c_int;
    pub fn rz_host_alloc(bytes: usize, err: *mut c_char, errlen: usize) -> *mut c_void;
    pub fn rz_host_free(p: *mut c_void);
    pub fn rz_host_trim(keep_bytes: u64) -> u64;
    pub fn rz_device_count() -> c_int;
    pub fn rz_version() -> *const c_char;
    pub fn rz_abi_layout(out: *mut u64, n: c_int) -> c_int;
}

/// An error of the library: the reference's variant + its message string.
#[derive(Debug, Clone)]
pub struct RzError {
    pub code: c_int,
    pub message: String,
}

pub struct ErrBuf([c_char; 512]);
impl ErrBuf {
    pub fn new() -> Self {
        ErrBuf([0; 512])
    }
    pub fn ptr(&mut self) -> *mut c_char {
        self.0.as_mut_ptr()
    }
    pub const LEN: usize = 512;
    pub fn to_error(&self, code: c_int) -> RzError {
        let message = unsafe { CStr::from_ptr(self.0.as_ptr()) }.to_string_lossy().into_owned();
        RzError { code, message }
    }
}
impl Default for ErrBuf {
    fn default() -> Self {
        Self::new()
    }
}

/// Compares this file's `#[repr(C)]` layouts with the library's own (`rz_abi_layout`): sizes of the five structs,
/// then selected field offsets.  Call once before the first compute call; a mismatch means header and crate have
/// drifted apart.
pub fn check_abi() -> Result<(), String> {
    use std::mem::{offset_of, size_of};
    let mine: [u64; 16] = [
        size_of::<RzRasterInfo>() as u64,
        size_of::<RzRawRasterInfo>() as u64,
        size_of::<RzGeomSoa>() as u64,
        size_of::<RzContext>() as u64,
        size_of::<RzStats>() as u64,
        offset_of!(RzContext, field) as u64,
        offset_of!(RzContext, band_of_geom) as u64,
        offset_of!(RzContext, background) as u64,
        offset_of!(RzContext, row_begin) as u64,
        offset_of!(RzContext, stream) as u64,
        offset_of!(RzContext, flags) as u64,
        offset_of!(RzStats, h2d_ms) as u64,
        offset_of!(RzStats, h2d_bytes) as u64,
        offset_of!(RzStats, kernel_launches) as u64,
        offset_of!(RzStats, n_mask_words) as u64,
        offset_of!(RzStats, wall_ms) as u64,
    ];
    let mut theirs = [0u64; 16];
    let n = unsafe { rz_abi_layout(theirs.as_mut_ptr(), 16) };
    if n != 16 || mine != theirs {
        return Err(format!("rz_b200 ABI mismatch: crate {mine:?} vs library {theirs:?}"));
    }
    Ok(())
}

/// Owned `rz_geoms*`: the flattened geometry set (host pools in page-locked memory + cached device copies).
pub struct Geoms(*mut rz_geoms);
unsafe impl Send for Geoms {}
impl Geoms {
    /// `&[geo::Geometry<f64>]` -> SoA (pooling rules of burn_geometry.rs:24-210) -> library handle.
    pub fn from_geometries(geoms: &[geo_types::Geometry<f64>]) -> Result<Self, RzError> {
        Self::from_soa(&flatten(geoms))
    }
    /// Coordinate arrays -> library handle (`rz_geoms_from_soa`: parallel flattening into page-locked pools).
    pub fn from_soa(soa: &GeomSoa) -> Result<Self, RzError> {
        let raw = soa.as_raw();
        let mut err = ErrBuf::new();
        let h = unsafe { rz_geoms_from_soa(&raw, err.ptr(), ErrBuf::LEN) };
        if h.is_null() { Err(err.to_error(RZ_RUNTIME_ERROR)) } else { Ok(Geoms(h)) }
    }
    pub fn as_ptr(&self) -> *mut rz_geoms {
        self.0
    }
    pub fn len(&self) -> usize {
        unsafe { rz_geoms_len(self.0) as usize }
    }
    pub fn is_empty(&self) -> bool {
        self.len() == 0
    }
}
impl Drop for Geoms {
    fn drop(&mut self) {
        unsafe { rz_geoms_free(self.0) }
    }
}

/// Owned `rz_sparse*`: the triplet stream of one call, in page-locked host memory owned by the library.
pub struct Sparse(*mut rz_sparse);
unsafe impl Send for Sparse {}
impl Sparse {
    /// # Safety
    /// `p` must come from `rz_rasterize_sparse` / `rz_rasterize_sparse_multi`.
    pub unsafe fn from_raw(p: *mut rz_sparse) -> Self {
        Sparse(p)
    }
    pub fn len(&self) -> usize {
        unsafe { rz_sparse_len(self.0) as usize }
    }
    pub fn is_empty(&self) -> bool {
        self.len() == 0
    }
    pub fn rows(&self) -> &[u64] {
        let n = self.len();
        if n == 0 { &[] } else { unsafe { std::slice::from_raw_parts(rz_sparse_rows(self.0), n) } }
    }
    pub fn cols(&self) -> &[u64] {
        let n = self.len();
        if n == 0 { &[] } else { unsafe { std::slice::from_raw_parts(rz_sparse_cols(self.0), n) } }
    }
    /// # Safety
    /// `N` must be the dtype the call was made with.
    pub unsafe fn data<N: RzDtype>(&self) -> &[N] {
        let n = self.len();
        if n == 0 { &[] } else { unsafe { std::slice::from_raw_parts(rz_sparse_data(self.0) as *const N, n) } }
    }
    /// Per-band triplet counts (`offsets` of rust/src/encoding/arrays.rs:63-70).
    pub fn counts(&self) -> &[u64] {
        let n = unsafe { rz_sparse_n_bands(self.0) } as usize;
        if n == 0 { &[] } else { unsafe { std::slice::from_raw_parts(rz_sparse_counts(self.0), n) } }
    }
}
impl Drop for Sparse {
    fn drop(&mut self) {
        unsafe { rz_sparse_free(self.0) }
    }
}

/// Every visible CUDA device, or the ordinals of `RZ_DEVICES` (comma separated).
pub fn default_devices() -> Vec<i32> {
    if let Ok(v) = std::env::var("RZ_DEVICES") {
        let d: Vec<i32> = v.split(',').filter_map(|s| s.trim().parse().ok()).collect();
        if !d.is_empty() {
            return d;
        }
    }
    (0..unsafe { rz_device_count() }.max(1)).collect()
}
