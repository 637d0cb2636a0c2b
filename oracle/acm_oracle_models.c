/*
 * acm_oracle_models.c -- CPU ORACLE. TEST INFRASTRUCTURE ONLY (see acm_oracle.h).
 *
 * project / unproject of the seven camera models, restated from the reference Rust in the
 * reference's exact operation order (IEEE binary64, no FMA contraction), returning the
 * CameraModelError variant as a status code.  "ref:" comments give file:line under
 * /root/reference.
 */
#include "acm_oracle.h"
#include <math.h>
#include <string.h>

#define EPS 2.220446049250313e-16          /* f64::EPSILON */
#define SQRT_EPS 1.4901161193847656e-08    /* f64::EPSILON.sqrt() == 2^-26 exactly */
#define PRECISION 1e-3

int orc_n_params(int model) {
    static const int np[7] = {4, 9, 8, 5, 6, 6, 5};
    return (model >= 0 && model < 7) ? np[model] : -1;
}

static void set_nan2(double* o) { o[0] = NAN; o[1] = NAN; }
static void set_nan3(double* o) { o[0] = NAN; o[1] = NAN; o[2] = NAN; }

/* nalgebra Vector3::normalize: n = sqrt((x*x + y*y) + z*z); each component divided by n */
static void normalize3(double x, double y, double z, double* o) {
    double n = sqrt(x * x + y * y + z * z);
    o[0] = x / n; o[1] = y / n; o[2] = z / n;
}

static int out_of_image(const orc_model* m, double u, double v) {
    /* ref: src/camera/mod.rs:157-166 / :194-206 (half-open [0,W) x [0,H)) */
    return u < 0.0 || u >= (double)m->width || v < 0.0 || v >= (double)m->height;
}

/* ------------------------------------------------------------------ project ---------- */
/* bounds != 0 : the trait's project (Pinhole/RadTan test image bounds);
 * bounds == 0 : geometric validity only (what a resolution-less factor can test). */
static int project_impl(const orc_model* m, const double X[3], double uv[2], int bounds) {
    const double fx = m->p[0], fy = m->p[1], cx = m->p[2], cy = m->p[3];
    const double x = X[0], y = X[1], z = X[2];
    double u, v;
    switch (m->model) {
    case ORC_PINHOLE: { /* ref: src/camera/pinhole.rs:165-182 */
        if (z < SQRT_EPS) { set_nan2(uv); return ORC_POINT_AT_CENTER; }
        u = fx * x / z + cx;
        v = fy * y / z + cy;
        if (bounds && out_of_image(m, u, v)) { set_nan2(uv); return ORC_PROJECTION_OUTSIDE_IMAGE; }
        break;
    }
    case ORC_RADTAN: { /* ref: src/camera/rad_tan.rs:302-348 */
        if (z < SQRT_EPS) { set_nan2(uv); return ORC_POINT_AT_CENTER; }
        const double k1 = m->p[4], k2 = m->p[5], p1 = m->p[6], p2 = m->p[7], k3 = m->p[8];
        double xp = x / z, yp = y / z;
        double r2 = xp * xp + yp * yp;
        double r4 = r2 * r2;
        double r6 = r4 * r2;
        double xd = xp * (1.0 + k1 * r2 + k2 * r4 + k3 * r6) + 2.0 * p1 * xp * yp + p2 * (r2 + 2.0 * xp * xp);
        double yd = yp * (1.0 + k1 * r2 + k2 * r4 + k3 * r6) + p1 * (r2 + 2.0 * yp * yp) + 2.0 * p2 * xp * yp;
        u = fx * xd + cx;
        v = fy * yd + cy;
        if (bounds && out_of_image(m, u, v)) { set_nan2(uv); return ORC_PROJECTION_OUTSIDE_IMAGE; }
        break;
    }
    case ORC_KB: { /* ref: src/camera/kannala_brandt.rs:340-394 (no image-bounds test) */
        if (z < 0.0) { set_nan2(uv); return ORC_POINT_OUTSIDE_IMAGE; }
        else if (z < EPS) { set_nan2(uv); return ORC_POINT_AT_CENTER; }
        const double k1 = m->p[4], k2 = m->p[5], k3 = m->p[6], k4 = m->p[7];
        double r_sq = x * x + y * y;
        double r = sqrt(r_sq);
        double theta = atan2(r, z);
        double theta2 = theta * theta;
        double theta3 = theta2 * theta;
        double theta5 = theta3 * theta2;
        double theta7 = theta5 * theta2;
        double theta9 = theta7 * theta2;
        double theta_d = theta + k1 * theta3 + k2 * theta5 + k3 * theta7 + k4 * theta9;
        double x_r, y_r;
        if (r < EPS) { x_r = 0.0; y_r = 0.0; } else { x_r = x / r; y_r = y / r; }
        u = fx * theta_d * x_r + cx;
        v = fy * theta_d * y_r + cy;
        break;
    }
    case ORC_UCM: { /* ref: src/camera/ucm.rs:297-316, check_proj_condition :154-161 */
        const double alpha = m->p[4];
        double d = sqrt(x * x + y * y + z * z);
        double denom = alpha * d + (1.0 - alpha) * z;
        double w = (alpha <= 0.5) ? alpha / (1.0 - alpha) : (1.0 - alpha) / alpha;
        int cond = z > -w * d;
        if (denom < PRECISION || !cond) { set_nan2(uv); return ORC_POINT_OUTSIDE_IMAGE; }
        u = fx * (x / denom) + cx;
        v = fy * (y / denom) + cy;
        break;
    }
    case ORC_EUCM: { /* ref: src/camera/eucm.rs:328-347, check_proj_condition :167-177 */
        const double alpha = m->p[4], beta = m->p[5];
        double d = sqrt(beta * (x * x + y * y) + z * z);
        double denom = alpha * d + (1.0 - alpha) * z;
        int cond = 1;
        if (alpha > 0.5) {
            double c = (alpha - 1.0) / (2.0 * alpha - 1.0);
            if (z < denom * c) cond = 0;
        }
        if (denom < PRECISION || !cond) { set_nan2(uv); return ORC_POINT_OUTSIDE_IMAGE; }
        u = fx * (x / denom) + cx;
        v = fy * (y / denom) + cy;
        break;
    }
    case ORC_DS: { /* ref: src/camera/double_sphere.rs:361-390, check_projection_condition :177-184 */
        const double alpha = m->p[4], xi = m->p[5];
        double r_squared = (x * x) + (y * y);
        double d1 = sqrt(r_squared + (z * z));
        double gamma = xi * d1 + z;
        double d2 = sqrt(r_squared + gamma * gamma);
        double denom = alpha * d2 + (1.0 - alpha) * gamma;
        double w1 = (alpha <= 0.5) ? alpha / (1.0 - alpha) : (1.0 - alpha) / alpha;
        double w2 = (w1 + xi) / sqrt(2.0 * w1 * xi + xi * xi + 1.0);
        int cond = z > -w2 * d1;
        if (denom < PRECISION || !cond) { set_nan2(uv); return ORC_POINT_OUTSIDE_IMAGE; }
        double mx = x / denom, my = y / denom;
        u = fx * (mx) + cx;
        v = fy * (my) + cy;
        break;
    }
    case ORC_FOV: { /* ref: src/camera/fov.rs:284-316 (no image-bounds test) */
        const double w = m->p[4];
        if (z < SQRT_EPS) { set_nan2(uv); return ORC_POINT_AT_CENTER; }
        double r2 = x * x + y * y;
        double r = sqrt(r2);
        double tan_w_half = tan(w / 2.0);
        double atan_wrd = atan2(2.0 * tan_w_half * r, z);
        double rd = (r2 < SQRT_EPS) ? 2.0 * tan_w_half / w : atan_wrd / (r * w);
        double mx = x * rd, my = y * rd;
        u = fx * mx + cx;
        v = fy * my + cy;
        break;
    }
    default: set_nan2(uv); return ORC_NUMERICAL;
    }
    uv[0] = u; uv[1] = v;
    return ORC_OK;
}

int orc_project(const orc_model* m, const double X[3], double uv[2]) { return project_impl(m, X, uv, 1); }
int orc_project_nobounds(const orc_model* m, const double X[3], double uv[2]) { return project_impl(m, X, uv, 0); }

/* ------------------------------------------------------------------ unproject -------- */
int orc_unproject(const orc_model* m, const double uvp[2], double ray[3]) {
    const double fx = m->p[0], fy = m->p[1], cx = m->p[2], cy = m->p[3];
    const double u = uvp[0], v = uvp[1];
    switch (m->model) {
    case ORC_PINHOLE: { /* ref: src/camera/pinhole.rs:228-246 */
        if (out_of_image(m, u, v)) { set_nan3(ray); return ORC_POINT_OUTSIDE_IMAGE; }
        double mx = (u - cx) / fx, my = (v - cy) / fy;
        double r2 = mx * mx + my * my;
        double norm = sqrt(1.0 + r2);
        double norm_inv = 1.0 / norm;
        ray[0] = mx * norm_inv; ray[1] = my * norm_inv; ray[2] = norm_inv;
        return ORC_OK;
    }
    case ORC_RADTAN: { /* ref: src/camera/rad_tan.rs:401-524 */
        if (out_of_image(m, u, v)) { set_nan3(ray); return ORC_POINT_OUTSIDE_IMAGE; }
        const double k1 = m->p[4], k2 = m->p[5], p1 = m->p[6], p2 = m->p[7], k3 = m->p[8];
        const double tx = (u - cx) / fx, ty = (v - cy) / fy;
        double px = tx, py = ty;
        const double TOL = 1e-6;
        for (int it = 0; it < 100; ++it) {
            double x = px, y = py;
            double r2 = x * x + y * y;
            double r4 = r2 * r2;
            double r6 = r4 * r2;
            double radial = 1.0 + k1 * r2 + k2 * r4 + k3 * r6;
            double xe = x * radial + 2.0 * p1 * x * y + p2 * (r2 + 2.0 * x * x);
            double ye = y * radial + p1 * (r2 + 2.0 * y * y) + 2.0 * p2 * x * y;
            double ex = xe - tx, ey = ye - ty;
            if (sqrt(ex * ex + ey * ey) < TOL) break;
            double dr_dx = 2.0 * x, dr_dy = 2.0 * y;
            double drad_dx = (k1 + 2.0 * k2 * r2 + 3.0 * k3 * r4) * dr_dx;
            double drad_dy = (k1 + 2.0 * k2 * r2 + 3.0 * k3 * r4) * dr_dy;
            double j00 = radial + x * drad_dx + 2.0 * p1 * y + p2 * (dr_dx + 4.0 * x);
            double j01 = x * drad_dy + 2.0 * p1 * x + p2 * (dr_dy);
            double j10 = y * drad_dx + p1 * (dr_dx) + 2.0 * p2 * y;
            double j11 = radial + y * drad_dy + p1 * (dr_dy + 4.0 * y) + 2.0 * p2 * x;
            /* nalgebra Matrix2::try_inverse: det = m11*m22 - m21*m12; zero => None */
            double det = j00 * j11 - j10 * j01;
            if (det == 0.0) { set_nan3(ray); return ORC_NUMERICAL; }
            double i00 = j11 / det, i01 = -j01 / det, i10 = -j10 / det, i11 = j00 / det;
            /* nalgebra gemv: column-by-column accumulation */
            double dx = i00 * ex + i01 * ey;
            double dy = i10 * ex + i11 * ey;
            px -= dx; py -= dy;
            if (sqrt(dx * dx + dy * dy) < TOL) break;
            if (it == 99) { set_nan3(ray); return ORC_NUMERICAL; }
        }
        normalize3(px, py, 1.0, ray);
        return ORC_OK;
    }
    case ORC_KB: { /* ref: src/camera/kannala_brandt.rs:445-562 */
        if (m->width > 0 && m->height > 0 && out_of_image(m, u, v)) { set_nan3(ray); return ORC_POINT_OUTSIDE_IMAGE; }
        const double k1 = m->p[4], k2 = m->p[5], k3 = m->p[6], k4 = m->p[7];
        double mx = (u - cx) / fx, my = (v - cy) / fy;
        double ru = sqrt(mx * mx + my * my);
        const double half_pi = 3.14159265358979323846 / 2.0;
        ru = fmin(ru, half_pi); /* Rust f64::min returns the non-NaN operand, like fmin */
        double theta = ru;
        int converged = 1;
        if (ru > 1e-6) {
            for (int i = 0; i < 10; ++i) {
                double t2 = theta * theta, t4 = t2 * t2, t6 = t4 * t2, t8 = t4 * t4;
                double a1 = k1 * t2, a2 = k2 * t4, a3 = k3 * t6, a4 = k4 * t8;
                double f = theta * (1.0 + a1 + a2 + a3 + a4) - ru;
                double fp = 1.0 + (3.0 * a1) + (5.0 * a2) + (7.0 * a3) + (9.0 * a4);
                if (fabs(fp) < EPS) { converged = 0; break; }
                double delta = f / fp;
                theta -= delta;
                if (fabs(delta) < 1e-6) break;
                if (i == 9) converged = 0;
            }
        } else {
            if (ru > 0.0) converged = 0; else { theta = 0.0; converged = 1; }
        }
        if (!converged) { set_nan3(ray); return ORC_NUMERICAL; }
        double xc, yc;
        if (fabs(ru) < EPS) { xc = 0.0; yc = 0.0; } else { xc = mx / ru; yc = my / ru; }
        double st = sin(theta), ct = cos(theta);
        normalize3(st * xc, st * yc, ct, ray);
        return ORC_OK;
    }
    case ORC_UCM: { /* ref: src/camera/ucm.rs:337-367, check_unproj_condition :177-184 */
        const double alpha = m->p[4];
        double gamma = 1.0 - alpha;
        double xi = alpha / gamma;
        double mx = (u - cx) / fx * gamma, my = (v - cy) / fy * gamma;
        double r2 = mx * mx + my * my;
        double num = xi + sqrt(1.0 + (1.0 - xi * xi) * r2);
        double denom = 1.0 - r2;
        int cond = 1;
        if (alpha > 0.5) { double g = 1.0 - alpha; cond = r2 <= g * g / (2.0 * alpha - 1.0); }
        if (denom < PRECISION || !cond) { set_nan3(ray); return ORC_POINT_OUTSIDE_IMAGE; }
        double coeff = num / denom;
        normalize3(coeff * mx - 0.0, coeff * my - 0.0, coeff - xi, ray);
        return ORC_OK;
    }
    case ORC_EUCM: { /* ref: src/camera/eucm.rs:368-398, check_unproj_condition :194-200 */
        const double alpha = m->p[4], beta = m->p[5];
        double mx = (u - cx) / fx, my = (v - cy) / fy;
        double r2 = mx * mx + my * my;
        double gamma = 1.0 - alpha;
        double num = 1.0 - r2 * alpha * alpha * beta;
        double det = 1.0 - (alpha - gamma) * beta * r2;
        double denom = gamma + alpha * sqrt(det);
        int cond = 1;
        if (alpha > 0.5 && r2 > (1.0 / beta * (2.0 * alpha - 1.0))) cond = 0; /* precedence quirk kept */
        if (det < PRECISION || !cond) { set_nan3(ray); return ORC_POINT_OUTSIDE_IMAGE; }
        double mz = num / denom;
        double norm = sqrt(mx * mx + my * my + mz * mz);
        ray[0] = mx / norm; ray[1] = my / norm; ray[2] = mz / norm;
        return ORC_OK;
    }
    case ORC_DS: { /* ref: src/camera/double_sphere.rs:436-476, check_unprojection_condition :200-209 */
        const double alpha = m->p[4], xi = m->p[5];
        double gamma = 1.0 - alpha;
        double mx = (u - cx) / fx, my = (v - cy) / fy;
        double r2 = (mx * mx) + (my * my);
        int cond = 1;
        if (alpha > 0.5) { if (r2 > 1.0 / (2.0 * alpha - 1.0)) cond = 0; }
        if (alpha != 0.0 && !cond) { set_nan3(ray); return ORC_POINT_OUTSIDE_IMAGE; }
        double mz = (1.0 - alpha * alpha * r2) / (alpha * sqrt(1.0 - (2.0 * alpha - 1.0) * r2) + gamma);
        double mz2 = mz * mz;
        double num = mz * xi + sqrt(mz2 + (1.0 - xi * xi) * r2);
        double denom = mz2 + r2;
        if (denom < PRECISION) { set_nan3(ray); return ORC_POINT_OUTSIDE_IMAGE; }
        double coeff = num / denom;
        normalize3(coeff * mx, coeff * my, coeff * mz - xi, ray);
        return ORC_OK;
    }
    case ORC_FOV: { /* ref: src/camera/fov.rs:336-363 (never fails) */
        const double w = m->p[4];
        double tan_w_2 = tan(w / 2.0);
        double mul2 = tan_w_2 * 2.0;
        double mx = (u - cx) / fx, my = (v - cy) / fy;
        double r2 = mx * mx + my * my;
        double rd = sqrt(r2);
        double x, y;
        if (mul2 > SQRT_EPS && rd > SQRT_EPS) {
            double s = sin(rd * w), c = cos(rd * w);
            double ru = s / (rd * mul2);
            x = mx * ru / c; y = my * ru / c;
        } else { x = mx; y = my; }
        normalize3(x, y, 1.0, ray);
        return ORC_OK;
    }
    default: set_nan3(ray); return ORC_NUMERICAL;
    }
}

/* ------------------------------------------------------------------ batches ---------- */
void orc_project_batch(const orc_model* m, const double* xyz, size_t n, double* uv, uint8_t* status, int nthreads) {
    (void)nthreads;
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(nthreads > 0 ? nthreads : 1)
#endif
    for (long long i = 0; i < (long long)n; ++i) status[i] = (uint8_t)orc_project(m, xyz + 3 * i, uv + 2 * i);
}

void orc_unproject_batch(const orc_model* m, const double* uv, size_t n, double* xyz, uint8_t* status, int nthreads) {
    (void)nthreads;
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(nthreads > 0 ? nthreads : 1)
#endif
    for (long long i = 0; i < (long long)n; ++i) status[i] = (uint8_t)orc_unproject(m, uv + 2 * i, xyz + 3 * i);
}
