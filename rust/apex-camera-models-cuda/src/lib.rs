//! `CameraModel` over the B200 hot path.
//!
//! Source only: the image this repository is built in has no Rust toolchain, so this crate has
//! never been compiled.  It is deliberately mechanical: every method is one call into `acm-sys`
//! whose behaviour is covered by the Python twin (`apex_camera_models_b200/camera.py`) in
//! `tests/test_gpu_parity.py`.
//!
//! The wrapper keeps the reference's own struct (YAML loading, `validate_params`, getters stay the
//! reference's code) and routes the per-point work to the GPU:
//!   * `project` / `unproject`            -> `acm_project_host` / `acm_unproject_host` (1-point batch)
//!   * `project_batch` / `unproject_batch`-> the same calls on whole `Matrix3xX` / `Matrix2xX`
//!   * `linear_estimation`                -> `acm_linear_estimation`
//!   * `*OptimizationCost::optimize`      -> `acm_lm_solve` (replaces apex-solver's factor + LM)
use acm_sys as sys;
use apex_camera_models::camera::{CameraModel, CameraModelError, Intrinsics, Resolution};
use nalgebra::{DVector, Matrix2xX, Matrix3xX, Vector2, Vector3};
use std::ffi::CStr;
use std::ptr;

pub const PINHOLE: i32 = 0;
pub const RAD_TAN: i32 = 1;
pub const KANNALA_BRANDT: i32 = 2;
pub const UCM: i32 = 3;
pub const EUCM: i32 = 4;
pub const DOUBLE_SPHERE: i32 = 5;
pub const FOV: i32 = 6;

/// One CUDA device + stream (acm_ctx).
pub struct Context(*mut sys::acm_ctx);
unsafe impl Send for Context {}

impl Context {
    pub fn new(device: i32) -> Result<Self, CameraModelError> {
        let mut h = ptr::null_mut();
        let rc = unsafe { sys::acm_ctx_create(device, ptr::null_mut(), &mut h) };
        if rc != sys::ACM_OK {
            let msg = unsafe { CStr::from_ptr(sys::acm_last_error(ptr::null())) }.to_string_lossy().into_owned();
            return Err(CameraModelError::NumericalError(format!("acm_ctx_create: {msg}")));
        }
        Ok(Context(h))
    }
    fn err(&self, rc: i32) -> CameraModelError {
        let msg = unsafe { CStr::from_ptr(sys::acm_last_error(self.0)) }.to_string_lossy().into_owned();
        match rc {
            sys::ACM_ERR_INVALID_PARAMS => CameraModelError::InvalidParams(msg),
            sys::ACM_ERR_FOCAL_LENGTH => CameraModelError::FocalLengthMustBePositive,
            sys::ACM_ERR_PRINCIPAL_POINT => CameraModelError::PrincipalPointMustBeFinite,
            _ => CameraModelError::NumericalError(msg),
        }
    }
}
impl Drop for Context {
    fn drop(&mut self) { unsafe { sys::acm_ctx_destroy(self.0); } }
}

/// status byte -> the `Err(..)` the reference's scalar call returns (mod.rs:79-113)
pub fn status_to_error(status: u8, model: i32) -> Option<CameraModelError> {
    match status {
        0 => None,
        1 => Some(CameraModelError::PointIsOutSideImage),
        2 => Some(CameraModelError::PointAtCameraCenter),
        3 => Some(CameraModelError::ProjectionOutSideImage),
        _ => Some(CameraModelError::NumericalError(
            if model == RAD_TAN { "Jacobian is singular".into() } else { "Unprojection failed to converge".into() })),
    }
}

/// A reference model (`M: CameraModel`, e.g. `DoubleSphereModel`) whose per-point work runs on the GPU.
pub struct GpuCamera<'c, M: CameraModel> {
    pub inner: M,
    pub model_id: i32,
    ctx: &'c Context,
}

impl<'c, M: CameraModel> GpuCamera<'c, M> {
    pub fn new(ctx: &'c Context, inner: M, model_id: i32) -> Self { GpuCamera { inner, model_id, ctx } }

    fn block(&self) -> sys::acm_camera {
        let i = self.inner.get_intrinsics();
        let r = self.inner.get_resolution();
        let d = self.inner.get_distortion();
        let mut params = [0.0f64; sys::ACM_MAX_PARAMS];
        params[..4].copy_from_slice(&[i.fx, i.fy, i.cx, i.cy]);
        params[4..4 + d.len()].copy_from_slice(&d);
        sys::acm_camera { model: self.model_id, width: r.width, height: r.height, n_params: (4 + d.len()) as i32, params }
    }

    /// Batched `project`: returns (uv, status byte per point). Matrix3xX memory is xyzxyz.. as the ABI expects.
    pub fn project_batch(&self, points_3d: &Matrix3xX<f64>) -> Result<(Matrix2xX<f64>, Vec<u8>), CameraModelError> {
        let n = points_3d.ncols();
        let mut uv = Matrix2xX::<f64>::zeros(n);
        let mut st = vec![0u8; n];
        let cam = self.block();
        let rc = unsafe { sys::acm_project_host(self.ctx.0, &cam, points_3d.as_ptr(), n, uv.as_mut_ptr(), st.as_mut_ptr()) };
        if rc != sys::ACM_OK { return Err(self.ctx.err(rc)); }
        Ok((uv, st))
    }

    pub fn unproject_batch(&self, points_2d: &Matrix2xX<f64>) -> Result<(Matrix3xX<f64>, Vec<u8>), CameraModelError> {
        let n = points_2d.ncols();
        let mut xyz = Matrix3xX::<f64>::zeros(n);
        let mut st = vec![0u8; n];
        let cam = self.block();
        let rc = unsafe { sys::acm_unproject_host(self.ctx.0, &cam, points_2d.as_ptr(), n, xyz.as_mut_ptr(), st.as_mut_ptr()) };
        if rc != sys::ACM_OK { return Err(self.ctx.err(rc)); }
        Ok((xyz, st))
    }

    /// `CameraModel::project` (mod.rs:256) as a one-point batch.
    pub fn project(&self, p: &Vector3<f64>) -> Result<Vector2<f64>, CameraModelError> {
        let (uv, st) = self.project_batch(&Matrix3xX::from_columns(&[*p]))?;
        match status_to_error(st[0], self.model_id) { Some(e) => Err(e), None => Ok(uv.column(0).into_owned()) }
    }

    pub fn unproject(&self, p: &Vector2<f64>) -> Result<Vector3<f64>, CameraModelError> {
        let (xyz, st) = self.unproject_batch(&Matrix2xX::from_columns(&[*p]))?;
        match status_to_error(st[0], self.model_id) { Some(e) => Err(e), None => Ok(xyz.column(0).into_owned()) }
    }

    /// One fused pass: (H = J^T J, g = J^T r, cost, n_valid) -- what `Factor::linearize` + the solver's
    /// J^T J produce in apex-solver, without materialising J.
    pub fn linearize(&self, residual_kind: i32, points_3d: &Matrix3xX<f64>, points_2d: &Matrix2xX<f64>)
        -> Result<sys::acm_normal_equations, CameraModelError> {
        if points_3d.ncols() != points_2d.ncols() {
            return Err(CameraModelError::InvalidParams("Number of 2D and 3D points must match".into()));
        }
        let cam = self.block();
        let mut ne: sys::acm_normal_equations = unsafe { std::mem::zeroed() };
        let rc = unsafe { sys::acm_linearize_host(self.ctx.0, &cam, residual_kind, points_3d.as_ptr(), points_2d.as_ptr(), points_3d.ncols(), &mut ne) };
        if rc != sys::ACM_OK { return Err(self.ctx.err(rc)); }
        Ok(ne)
    }
}

/// Device point buffer owned for the duration of a call.
struct DevicePoints<'c> { ctx: &'c Context, h: *mut sys::acm_points }
impl<'c> DevicePoints<'c> {
    fn upload(ctx: &'c Context, dim: i32, host: *const f64, n: usize) -> Result<Self, CameraModelError> {
        let mut h = ptr::null_mut();
        let mut rc = unsafe { sys::acm_points_create(ctx.0, dim, n, sys::ACM_F64, &mut h) };
        if rc == sys::ACM_OK && n > 0 { rc = unsafe { sys::acm_points_upload_aos_f64(ctx.0, h, host, n) }; }
        if rc != sys::ACM_OK { unsafe { sys::acm_points_destroy(ctx.0, h); } return Err(ctx.err(rc)); }
        Ok(DevicePoints { ctx, h })
    }
    fn adopt(ctx: &'c Context, h: *mut sys::acm_points) -> Self { DevicePoints { ctx, h } }
    fn len(&self) -> usize { unsafe { sys::acm_points_len(self.h) } }
    fn download(&self, out: *mut f64) -> Result<(), CameraModelError> {
        let rc = unsafe { sys::acm_points_download_aos_f64(self.ctx.0, self.h, out, self.len()) };
        if rc != sys::ACM_OK { Err(self.ctx.err(rc)) } else { Ok(()) }
    }
}
impl<'c> Drop for DevicePoints<'c> { fn drop(&mut self) { unsafe { sys::acm_points_destroy(self.ctx.0, self.h); } } }

impl<'c, M: CameraModel> GpuCamera<'c, M> {
    /// `util::sample_points(Some(&model), n) -> (Matrix2xX, Matrix3xX)` (point_sampling.rs:46-120):
    /// grid of cell centres -> unproject -> keep Ok && z > 0, order preserved.
    pub fn sample_points(&self, n: usize) -> Result<(Matrix2xX<f64>, Matrix3xX<f64>), CameraModelError> {
        let cam = self.block();
        let (mut uv, mut xyz, mut kept) = (ptr::null_mut(), ptr::null_mut(), 0usize);
        let rc = unsafe { sys::acm_sample_points(self.ctx.0, &cam, n, &mut uv, &mut xyz, &mut kept) };
        if rc != sys::ACM_OK { return Err(self.ctx.err(rc)); }
        let (uv, xyz) = (DevicePoints::adopt(self.ctx, uv), DevicePoints::adopt(self.ctx, xyz));
        let mut p2 = Matrix2xX::<f64>::zeros(kept);
        let mut p3 = Matrix3xX::<f64>::zeros(kept);
        uv.download(p2.as_mut_ptr())?;
        xyz.download(p3.as_mut_ptr())?;
        Ok((p2, p3))
    }

    /// `util::compute_reprojection_error(Some(&model), &points3d, &points2d)` (error_metrics.rs:62-121).
    pub fn compute_reprojection_error(&self, points_3d: &Matrix3xX<f64>, points_2d: &Matrix2xX<f64>)
        -> Result<sys::acm_projection_error, CameraModelError> {
        if points_3d.ncols() != points_2d.ncols() {
            return Err(CameraModelError::InvalidParams("Number of 2D and 3D points must match".into()));
        }
        let n = points_3d.ncols();
        let x = DevicePoints::upload(self.ctx, 3, points_3d.as_ptr(), n)?;
        let u = DevicePoints::upload(self.ctx, 2, points_2d.as_ptr(), n)?;
        let cam = self.block();
        let mut out = sys::acm_projection_error::default();
        let rc = unsafe { sys::acm_reprojection_error(self.ctx.0, &cam, x.h, u.h, &mut out) };
        if rc != sys::ACM_OK { return Err(self.ctx.err(rc)); }   // -7 -> UtilError::ZeroProjectionPoints upstream
        Ok(out)
    }

    /// `util::undistort_image(&img, &model, target, method)` (undistort.rs:14-49) over the raw RGB8 buffer
    /// (`RgbImage::as_raw()`); `interpolation` = `sys::ACM_INTERP_NEAREST | ACM_INTERP_BILINEAR`.
    pub fn undistort_image(&self, image: &[u8], width: u32, height: u32, target: Option<Intrinsics>, interpolation: i32)
        -> Result<Vec<u8>, CameraModelError> {
        let r = self.inner.get_resolution();
        if width != r.width || height != r.height || image.len() != width as usize * height as usize * 3 {   // undistort.rs:23-28
            return Err(CameraModelError::InvalidParams(format!("Image {}x{} doesn't match model {}x{}", width, height, r.width, r.height)));
        }
        let cam = self.block();
        let t = target.map(|t| [t.fx, t.fy, t.cx, t.cy]);
        let mut out = vec![0u8; image.len()];
        let rc = unsafe {
            sys::acm_undistort_rgb8_host(self.ctx.0, &cam, t.as_ref().map_or(ptr::null(), |a| a.as_ptr()), image.as_ptr(), out.as_mut_ptr(), 1, interpolation)
        };
        if rc != sys::ACM_OK { Err(self.ctx.err(rc)) } else { Ok(out) }
    }

    /// README-era `project(&p, compute_jacobian = true)`: uv plus the 2xP Jacobian w.r.t. `[fx, fy, cx, cy, dist..]`
    /// (per-model docs, double_sphere.rs:326-332) and the 2x3 Jacobian w.r.t. the 3-D point (trait doc, mod.rs:246-252).
    pub fn project_with_jacobians(&self, p: &Vector3<f64>)
        -> Result<(Vector2<f64>, nalgebra::DMatrix<f64>, nalgebra::Matrix2x3<f64>), CameraModelError> {
        let cam = self.block();
        let np = cam.n_params as usize;
        let x = DevicePoints::upload(self.ctx, 3, p.as_ptr(), 1)?;
        let u = DevicePoints::upload(self.ctx, 2, [0.0f64; 2].as_ptr(), 1)?;
        let (mut dj, mut ds) = (ptr::null_mut(), ptr::null_mut());
        let mut rc = unsafe { sys::acm_device_alloc(self.ctx.0, (2 * np + 6) * 8, &mut dj) };
        if rc == sys::ACM_OK { rc = unsafe { sys::acm_device_alloc(self.ctx.0, 8, &mut ds) }; }
        let mut jp = vec![0.0f64; 2 * np];
        let mut jx = [0.0f64; 6];
        let mut st = [0u8; 1];
        let mut uv = [0.0f64; 2];
        unsafe {
            let djp = dj as *mut f64;
            let djx = djp.add(2 * np);
            if rc == sys::ACM_OK { rc = sys::acm_project_jacobian(self.ctx.0, &cam, x.h, u.h, djp, ds as *mut u8); }
            if rc == sys::ACM_OK { rc = sys::acm_project_point_jacobian(self.ctx.0, &cam, x.h, u.h, djx, ds as *mut u8); }
            if rc == sys::ACM_OK { rc = sys::acm_memcpy_d2h(self.ctx.0, jp.as_mut_ptr() as *mut _, djp as *const _, 2 * np * 8); }
            if rc == sys::ACM_OK { rc = sys::acm_memcpy_d2h(self.ctx.0, jx.as_mut_ptr() as *mut _, djx as *const _, 48); }
            if rc == sys::ACM_OK { rc = sys::acm_memcpy_d2h(self.ctx.0, st.as_mut_ptr() as *mut _, ds as *const _, 1); }
            if rc == sys::ACM_OK { rc = sys::acm_points_download_aos_f64(self.ctx.0, u.h, uv.as_mut_ptr(), 1); }
            if rc == sys::ACM_OK { rc = sys::acm_ctx_sync(self.ctx.0); }
            sys::acm_device_free(self.ctx.0, dj); sys::acm_device_free(self.ctx.0, ds);
        }
        if rc != sys::ACM_OK { return Err(self.ctx.err(rc)); }
        if let Some(e) = status_to_error(st[0], self.model_id) { return Err(e); }
        // one point: the rows of n doubles are the row-major matrices themselves
        Ok((Vector2::new(uv[0], uv[1]), nalgebra::DMatrix::from_row_slice(2, np, &jp), nalgebra::Matrix2x3::from_row_slice(&jx)))
    }
}

/// Resident correspondences + optimiser: the README-era `*OptimizationCost` and the converter's
/// `Problem` + `LevenbergMarquardt::with_config(cfg).optimize(..)` (camera_converter.rs:378-420).
pub struct OptimizationCost<'c> {
    ctx: &'c Context,
    pub camera: sys::acm_camera,
    xyz: *mut sys::acm_points,
    uv: *mut sys::acm_points,
    pub residual_kind: i32,
}

impl<'c> OptimizationCost<'c> {
    pub fn new(ctx: &'c Context, camera: sys::acm_camera, points_3d: &Matrix3xX<f64>, points_2d: &Matrix2xX<f64>, residual_kind: i32)
        -> Result<Self, CameraModelError> {
        assert_eq!(points_3d.ncols(), points_2d.ncols());
        let n = points_3d.ncols();
        let (mut xyz, mut uv) = (ptr::null_mut(), ptr::null_mut());
        unsafe {
            let mut rc = sys::acm_points_create(ctx.0, 3, n, sys::ACM_F64, &mut xyz);
            if rc == sys::ACM_OK { rc = sys::acm_points_create(ctx.0, 2, n, sys::ACM_F64, &mut uv); }
            if rc == sys::ACM_OK { rc = sys::acm_points_upload_aos_f64(ctx.0, xyz, points_3d.as_ptr(), n); }
            if rc == sys::ACM_OK { rc = sys::acm_points_upload_aos_f64(ctx.0, uv, points_2d.as_ptr(), n); }
            if rc == sys::ACM_OK { rc = sys::acm_ctx_sync(ctx.0); }
            if rc != sys::ACM_OK { return Err(ctx.err(rc)); }
        }
        Ok(OptimizationCost { ctx, camera, xyz, uv, residual_kind })
    }

    pub fn linear_estimation(&mut self) -> Result<(), CameraModelError> {
        let rc = unsafe { sys::acm_linear_estimation(self.ctx.0, &mut self.camera, self.xyz, self.uv) };
        if rc != sys::ACM_OK { Err(self.ctx.err(rc)) } else { Ok(()) }
    }

    /// `bounds`: (lower, upper) per parameter = `problem.set_variable_bounds("params", i, lo, hi)`.
    pub fn optimize(&mut self, bounds: Option<(&[f64], &[f64])>, cfg: Option<sys::acm_lm_config>)
        -> Result<(DVector<f64>, sys::acm_lm_result), CameraModelError> {
        let mut c = unsafe { std::mem::zeroed::<sys::acm_lm_config>() };
        unsafe { sys::acm_lm_default_config(&mut c); }
        let c = cfg.unwrap_or(c);
        let (lo, hi) = match bounds { Some((l, h)) => (l.as_ptr(), h.as_ptr()), None => (ptr::null(), ptr::null()) };
        let mut out = [0.0f64; sys::ACM_MAX_PARAMS];
        let mut res = sys::acm_lm_result::default();
        let rc = unsafe { sys::acm_lm_solve(self.ctx.0, &self.camera, self.residual_kind, self.xyz, self.uv, lo, hi, &c, out.as_mut_ptr(), &mut res) };
        if rc != sys::ACM_OK { return Err(self.ctx.err(rc)); }
        let p = self.camera.n_params as usize;
        self.camera.params[..p].copy_from_slice(&out[..p]);
        Ok((DVector::from_column_slice(&out[..p]), res))
    }

    pub fn get_intrinsics(&self) -> Intrinsics {
        Intrinsics { fx: self.camera.params[0], fy: self.camera.params[1], cx: self.camera.params[2], cy: self.camera.params[3] }
    }
    pub fn get_distortion(&self) -> Vec<f64> { self.camera.params[4..self.camera.n_params as usize].to_vec() }
    pub fn get_resolution(&self) -> Resolution { Resolution { width: self.camera.width, height: self.camera.height } }
}

impl<'c> Drop for OptimizationCost<'c> {
    fn drop(&mut self) {
        unsafe { sys::acm_points_destroy(self.ctx.0, self.xyz); sys::acm_points_destroy(self.ctx.0, self.uv); }
    }
}

/// `util::ImageQualityMetrics` (image_quality.rs:20-26).
#[derive(Debug, Clone)]
pub struct ImageQualityMetrics { pub psnr: f64, pub ssim: f64 }

/// RGB8 image staged in HBM for the duration of a call.
struct DeviceImage<'c> { ctx: &'c Context, ptr: *mut std::ffi::c_void, bytes: usize }
impl<'c> DeviceImage<'c> {
    fn new(ctx: &'c Context, bytes: usize, host: Option<&[u8]>) -> Result<Self, CameraModelError> {
        let mut p = ptr::null_mut();
        let mut rc = unsafe { sys::acm_device_alloc(ctx.0, bytes.max(4), &mut p) };
        if rc == sys::ACM_OK && bytes > 0 {
            rc = match host {
                Some(h) => unsafe { sys::acm_memcpy_h2d(ctx.0, p, h.as_ptr() as *const _, bytes) },
                None => unsafe { sys::acm_memset_d(ctx.0, p, 0, bytes) },
            };
        }
        if rc != sys::ACM_OK { return Err(ctx.err(rc)); }
        Ok(DeviceImage { ctx, ptr: p, bytes })
    }
    fn download(&self) -> Result<Vec<u8>, CameraModelError> {
        let mut out = vec![0u8; self.bytes];
        let mut rc = unsafe { sys::acm_memcpy_d2h(self.ctx.0, out.as_mut_ptr() as *mut _, self.ptr, self.bytes) };
        if rc == sys::ACM_OK { rc = unsafe { sys::acm_ctx_sync(self.ctx.0) }; }
        if rc != sys::ACM_OK { Err(self.ctx.err(rc)) } else { Ok(out) }
    }
}
impl<'c> Drop for DeviceImage<'c> { fn drop(&mut self) { unsafe { sys::acm_device_free(self.ctx.0, self.ptr); } } }

/// `util::calculate_psnr(&RgbImage, &RgbImage)` (image_quality.rs:45-89) over raw RGB8 buffers
/// (`RgbImage::as_raw()`), `width` x `height`.
pub fn calculate_psnr(ctx: &Context, img1: &[u8], img2: &[u8], width: u32, height: u32) -> Result<f64, CameraModelError> {
    if img1.len() != img2.len() { return Err(CameraModelError::InvalidParams("Images must have the same dimensions".into())); }
    let (a, b) = (DeviceImage::new(ctx, img1.len(), Some(img1))?, DeviceImage::new(ctx, img2.len(), Some(img2))?);
    let mut out = 0.0f64;
    let rc = unsafe { sys::acm_image_psnr(ctx.0, a.ptr as *const u8, b.ptr as *const u8, width, height, &mut out) };
    if rc != sys::ACM_OK { Err(ctx.err(rc)) } else { Ok(out) }
}

/// `util::calculate_ssim` (image_quality.rs:108-210).
pub fn calculate_ssim(ctx: &Context, img1: &[u8], img2: &[u8], width: u32, height: u32) -> Result<f64, CameraModelError> {
    if img1.len() != img2.len() { return Err(CameraModelError::InvalidParams("Images must have the same dimensions".into())); }
    let (a, b) = (DeviceImage::new(ctx, img1.len(), Some(img1))?, DeviceImage::new(ctx, img2.len(), Some(img2))?);
    let mut out = 0.0f64;
    let rc = unsafe { sys::acm_image_ssim(ctx.0, a.ptr as *const u8, b.ptr as *const u8, width, height, &mut out) };
    if rc != sys::ACM_OK { Err(ctx.err(rc)) } else { Ok(out) }
}

/// `util::compute_image_quality_metrics` (image_quality.rs:254-324) over camera blocks; returns the metrics and,
/// when `want_image`, the combined display image (green input / magenta output projections over `reference`).
pub fn compute_image_quality_metrics(ctx: &Context, input_model: &sys::acm_camera, output_model: &sys::acm_camera,
                                     points_3d: &Matrix3xX<f64>, width: u32, height: u32, reference: Option<&[u8]>, want_image: bool)
    -> Result<(ImageQualityMetrics, Option<Vec<u8>>), CameraModelError> {
    let n = points_3d.ncols();
    let bytes = width as usize * height as usize * 3;
    let mut xyz = ptr::null_mut();
    let mut rc = unsafe { sys::acm_points_create(ctx.0, 3, n, sys::ACM_F64, &mut xyz) };
    if rc == sys::ACM_OK { rc = unsafe { sys::acm_points_upload_aos_f64(ctx.0, xyz, points_3d.as_ptr(), n) }; }
    if rc != sys::ACM_OK { unsafe { sys::acm_points_destroy(ctx.0, xyz); } return Err(ctx.err(rc)); }
    let dref = match reference { Some(r) => Some(DeviceImage::new(ctx, bytes, Some(r))?), None => None };
    let dcomb = if want_image { Some(DeviceImage::new(ctx, bytes, None)?) } else { None };
    let mut out = sys::acm_image_quality::default();
    let rc = unsafe {
        sys::acm_image_quality_metrics(ctx.0, input_model, output_model, xyz, width, height,
                                       dref.as_ref().map_or(ptr::null(), |d| d.ptr as *const u8),
                                       dcomb.as_ref().map_or(ptr::null_mut(), |d| d.ptr as *mut u8), &mut out)
    };
    unsafe { sys::acm_points_destroy(ctx.0, xyz); }
    if rc != sys::ACM_OK { return Err(ctx.err(rc)); }   // ACM_ERR_ZERO_PROJECTION_POINTS -> UtilError::ZeroProjectionPoints upstream
    let img = match dcomb { Some(d) => Some(d.download()?), None => None };
    Ok((ImageQualityMetrics { psnr: out.psnr, ssim: out.ssim }, img))
}
