//! rasterize_b200.rs — the `ArrayBuilder` bodies of the reference routed through librz_b200.so.
//!
//! SOURCE ONLY (no Rust toolchain in the image this repository is built in).  How a maintainer of ttrotto/rusterize
//! wires it in:
//!
//! 1. add `integration/rusterize-b200-sys` to the workspace and to `rust/Cargo.toml`:
//!        rusterize-b200-sys = { path = "../rusterize-b200-sys", optional = true }
//!        [features] b200 = ["dep:rusterize-b200-sys"]
//! 2. copy this file to `rust/src/rasterize_b200.rs`, declare `#[cfg(feature = "b200")] mod rasterize_b200;` in
//!    `rust/src/lib.rs`;
//! 3. in `rust/src/rasterize.rs`, first statement of `DenseArray::<N>::build` (:77) and `SparseArray::<N>::build`
//!    (:124):
//!        #[cfg(feature = "b200")]
//!        return crate::rasterize_b200::build_dense(geoms, ctx);     // resp. build_sparse
//!    (`N: RasterDtype` gains the bound `rusterize_b200_sys::RzDtype` under the feature; all ten dtypes implement it.)
//!
//! The Python (python/src/rusterize.rs:121-124) and R (R/rusterize/src/rust/src/rusterize.rs:52) bindings call
//! `geoms.rasterize::<A>(ctx)` and therefore route through the B200 library unchanged.
use std::ffi::{c_void, CString};

use geo::Geometry;
use ndarray::Array3;
use rusterize_b200_sys as sys;

use crate::{
    encoding::arrays::{DenseArray, SparseArray},
    error::{RusterizeError, RusterizeResult},
    prelude::{PixelFunction, RasterDtype, RasterizeContext},
    rasterize::FieldSource,
};

/// `RusterizeError` carries `&'static str`: the library's messages are the reference's own strings, matched back to
/// static ones; anything else (CUDA errors) is leaked once per distinct failure.
fn to_error(e: sys::RzError) -> RusterizeError {
    const KNOWN: [&str; 2] = ["Geometry and field lengths must match", "Geometry and by lengths must match"];
    let msg: &'static str = KNOWN
        .iter()
        .copied()
        .find(|k| *k == e.message)
        .unwrap_or_else(|| Box::leak(e.message.into_boxed_str()));
    if e.code == sys::RZ_VALUE_ERROR { RusterizeError::ValueError(msg) } else { RusterizeError::RuntimeError(msg) }
}

fn pixel_fn(f: &PixelFunction) -> sys::RzPixelFn {
    match f {
        PixelFunction::Sum => sys::RzPixelFn::Sum,
        PixelFunction::First => sys::RzPixelFn::First,
        PixelFunction::Last => sys::RzPixelFn::Last,
        PixelFunction::Min => sys::RzPixelFn::Min,
        PixelFunction::Max => sys::RzPixelFn::Max,
        PixelFunction::Count => sys::RzPixelFn::Count,
        PixelFunction::Any => sys::RzPixelFn::Any,
    }
}

/// `group_keys` (rust/src/rasterize.rs:199-205) done by the library: band ids per geometry + sorted band names.
fn bands(by: &[String]) -> (Vec<i32>, Vec<String>) {
    let keys: Vec<CString> = by.iter().map(|k| CString::new(k.as_bytes()).expect("NUL in `by` key")).collect();
    let ptrs: Vec<*const i8> = keys.iter().map(|k| k.as_ptr()).collect();
    let mut band = vec![0i32; by.len()];
    let mut first = vec![0u64; by.len().max(1)];
    let nb = unsafe { sys::rz_group_keys(ptrs.as_ptr().cast(), by.len() as u64, band.as_mut_ptr(), first.as_mut_ptr()) };
    let names = first[..nb as usize].iter().map(|&i| by[i as usize].clone()).collect();
    (band, names)
}

/// Everything a call borrows: the context points into these.
struct Call<N> {
    soa: sys::GeomSoa, // coordinate arrays in the library's input layout (pooling rules applied, not yet flattened)
    scalar: [N; 1],
    background: [N; 1],
    band: Option<Vec<i32>>,
    band_names: Vec<String>,
    #[cfg(feature = "polars")]
    column: Option<(Vec<N>, Vec<u8>)>, // values (nulls filled) + validity
}

fn prepare<N: RasterDtype + sys::RzDtype>(geoms: &[Geometry<f64>], ctx: &RasterizeContext<N>) -> RusterizeResult<(Call<N>, sys::RzContext)> {
    sys::check_abi().map_err(|m| RusterizeError::RuntimeError(Box::leak(m.into_boxed_str())))?;
    let soa = sys::flatten(geoms);
    let (band, band_names) = match ctx.by {
        Some(by) => {
            let (b, n) = bands(by);
            (Some(b), n)
        }
        None => (None, vec![String::from("band_1")]),
    };
    let mut call = Call {
        soa,
        scalar: [N::one()],
        background: [ctx.background],
        band,
        band_names,
        #[cfg(feature = "polars")]
        column: None,
    };
    let ri = &ctx.raster_info;
    let mut c = sys::RzContext {
        raster_info: sys::RzRasterInfo {
            nrows: ri.nrows as u64,
            ncols: ri.ncols as u64,
            xmin: ri.xmin,
            ymin: ri.ymin,
            xmax: ri.xmax,
            ymax: ri.ymax,
            xres: ri.xres,
            yres: ri.yres,
            epsg: ri.epsg.map(|e| e as i32).unwrap_or(-1),
            _pad: 0,
        },
        dtype: N::CODE,
        pixel_fn: pixel_fn(&ctx.pixel_fn) as i32,
        field: std::ptr::null(),
        field_is_scalar: 0,
        all_touched: ctx.all_touched as i32,
        field_len: 0,
        field_valid: std::ptr::null(),
        band_of_geom: std::ptr::null(),
        by_len: 0,
        n_bands: call.band_names.len() as i32,
        device: 0,
        background: std::ptr::null(),
        row_begin: 0,
        row_end: 0,
        stream: std::ptr::null_mut(),
        flags: 0,
        tile_bytes: 0,
    };
    match &ctx.field {
        FieldSource::Scalar(s) => {
            call.scalar = [*s];
            c.field_is_scalar = 1;
        }
        FieldSource::Array(arr) => {
            // (a non-contiguous view is copied; `as_slice` succeeds for the views the bindings build)
            c.field = match arr.as_slice() {
                Some(s) => s.as_ptr().cast(),
                None => return Err(RusterizeError::ValueError("`field` must be a contiguous array")),
            };
            c.field_len = arr.len() as u64;
        }
        #[cfg(feature = "polars")]
        FieldSource::Column(col) => {
            // null field => geometry skipped (rust/src/rasterize.rs:187-192): validity bytes for the library
            let ca = col.as_materialized_series().unpack::<N::ChunkedArrayType>().unwrap();
            let mut vals = Vec::with_capacity(ca.len());
            let mut valid = Vec::with_capacity(ca.len());
            for v in ca.iter() {
                valid.push(v.is_some() as u8);
                vals.push(v.unwrap_or_else(N::zero));
            }
            call.column = Some((vals, valid));
            c.field_len = ca.len() as u64;
        }
    }
    Ok((call, c))
}

/// Point the context at the buffers `call` owns (after `call` has reached its final place in memory).
fn bind<N>(call: &Call<N>, c: &mut sys::RzContext) {
    if c.field_is_scalar == 1 {
        c.field = call.scalar.as_ptr().cast();
    }
    #[cfg(feature = "polars")]
    if let Some((vals, valid)) = &call.column {
        c.field = vals.as_ptr().cast();
        c.field_valid = valid.as_ptr();
    }
    c.background = call.background.as_ptr().cast();
    if let Some(b) = &call.band {
        c.band_of_geom = b.as_ptr();
        c.by_len = b.len() as u64;
    }
}

/// `DenseArray::<N>::build` (rust/src/rasterize.rs:77-115): one owned `[band][row][col]` array, filled by every
/// visible GPU (row bands) in one library call.
pub(crate) fn build_dense<N: RasterDtype + sys::RzDtype>(geoms: &[Geometry<f64>], ctx: RasterizeContext<N>) -> RusterizeResult<DenseArray<N>> {
    crate::rasterize::assert_matching_len(geoms.len(), &ctx.field, ctx.by)?; // (made pub(crate))
    let (call, mut c) = prepare(geoms, &ctx)?;
    bind(&call, &mut c);
    let shape = (call.band_names.len(), ctx.raster_info.nrows, ctx.raster_info.ncols);
    // the library writes every element (background included: rust/src/geo/raster.rs:23-28 is fused into the fill)
    let mut raster = Array3::<N>::uninit(shape);
    let devices = sys::default_devices();
    let mut err = sys::ErrBuf::new();
    let mut stats = sys::RzStats::default();
    // one call: every device flattens the parts of its rows straight out of `soa`, uploads, burns, copies back
    let raw_soa = call.soa.as_raw();
    let rc = unsafe {
        sys::rz_rasterize_dense_soa(&raw_soa, &c, devices.as_ptr(), devices.len() as i32,
                                    raster.as_mut_ptr() as *mut c_void, &mut stats, std::ptr::null_mut(), err.ptr(),
                                    sys::ErrBuf::LEN)
    };
    if rc != sys::RZ_OK {
        return Err(to_error(err.to_error(rc)));
    }
    Ok(DenseArray::new(unsafe { raster.assume_init() }, call.band_names, ctx.raster_info))
}

/// `SparseArray::<N>::build` (rust/src/rasterize.rs:124-156): the triplet stream of every band, in burn order
/// (rust/src/encoding/writers.rs:101-131), produced by contiguous geometry ranges per GPU and concatenated by offset.
pub(crate) fn build_sparse<N: RasterDtype + sys::RzDtype>(geoms: &[Geometry<f64>], ctx: RasterizeContext<N>) -> RusterizeResult<SparseArray<N>> {
    crate::rasterize::assert_matching_len(geoms.len(), &ctx.field, ctx.by)?;
    let (call, mut c) = prepare(geoms, &ctx)?;
    bind(&call, &mut c);
    let devices = sys::default_devices();
    let handle = sys::Geoms::from_soa(&call.soa).map_err(to_error)?;
    let mut err = sys::ErrBuf::new();
    let mut stats = sys::RzStats::default();
    let mut raw: *mut sys::rz_sparse = std::ptr::null_mut();
    let rc = unsafe {
        sys::rz_rasterize_sparse_multi(handle.as_ptr(), &c, devices.as_ptr(), devices.len() as i32, &mut raw, &mut stats,
                                       std::ptr::null_mut(), err.ptr(), sys::ErrBuf::LEN)
    };
    if rc != sys::RZ_OK {
        return Err(to_error(err.to_error(rc)));
    }
    let sp = unsafe { sys::Sparse::from_raw(raw) };
    // SparseArray owns Vecs (rust/src/encoding/arrays.rs:63-70): one copy out of the library's page-locked block
    let rows = sp.rows().to_vec();
    let cols = sp.cols().to_vec();
    let data = unsafe { sp.data::<N>() }.to_vec();
    let offsets = sp.counts().iter().map(|&n| n as usize).collect();
    Ok(SparseArray::new(call.band_names, rows, cols, data, offsets, ctx))
}
