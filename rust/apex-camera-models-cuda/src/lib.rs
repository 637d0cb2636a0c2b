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
