// 2x3 Jacobian of the projection w.r.t. the 3-D point -- the "Jacobian matrix (2x3)" of the README-era
// `project(&p, compute_jacobian)` (stale trait doc, reference src/camera/mod.rs:246-252).  The reference tree
// holds no code for it (third column); formulas derived from the model definitions and
// pinned by mpmath 50-digit differences (tests/golden/mpmath_point_jacobians.json).
// ju = d u / d(x, y, z), jv = d v / d(x, y, z).  Validity is decided by the caller (project<false>).
#pragma once
#include "acm_linearize.cuh"

template <int M> struct PointJac;

template <> struct PointJac<ACM_MODEL_PINHOLE> {
    static __device__ __forceinline__ void eval(const LinParams& p, double x, double y, double z, double* ju, double* jv) {
        const double iz = 1.0 / z;
        ju[0] = p.fx * iz; ju[1] = 0.0; ju[2] = -p.fx * x * iz * iz;
        jv[0] = 0.0; jv[1] = p.fy * iz; jv[2] = -p.fy * y * iz * iz;
    }
};

// chain rule through (a, b) = (x/z, y/z): the 2x2 distortion Jacobian is the one of rad_tan.rs:471-486
template <> struct PointJac<ACM_MODEL_RADTAN> {
    static __device__ __forceinline__ void eval(const LinParams& p, double x, double y, double z, double* ju, double* jv) {
        const double k1 = p.d[0], k2 = p.d[1], p1 = p.d[2], p2 = p.d[3], k3 = p.d[4];
        const double iz = 1.0 / z, a = x * iz, b = y * iz;
        const double rho = a * a + b * b;
        const double rad = 1.0 + rho * (k1 + rho * (k2 + rho * k3));
        const double c = 2.0 * (k1 + rho * (2.0 * k2 + 3.0 * k3 * rho));   // d rad / d a = c * a, d rad / d b = c * b
        const double d00 = rad + a * a * c + 2.0 * p1 * b + 6.0 * p2 * a;
        const double d01 = a * b * c + 2.0 * p1 * a + 2.0 * p2 * b;
        const double d11 = rad + b * b * c + 6.0 * p1 * b + 2.0 * p2 * a;   // d10 == d01
        const double fz = p.fx * iz, gz = p.fy * iz;
        ju[0] = fz * d00; ju[1] = fz * d01; ju[2] = -fz * (d00 * a + d01 * b);
        jv[0] = gz * d01; jv[1] = gz * d11; jv[2] = -gz * (d01 * a + d11 * b);
    }
};

template <> struct PointJac<ACM_MODEL_KANNALA_BRANDT> {
    static __device__ __forceinline__ void eval(const LinParams& p, double x, double y, double z, double* ju, double* jv) {
        const double r2 = x * x + y * y, r = sqrt(r2);
        if (r < LIN_EPS) {   // the reference returns the principal point here; analytic limit on the axis
            ju[0] = p.fx / z; ju[1] = ju[2] = 0.0; jv[0] = jv[2] = 0.0; jv[1] = p.fy / z;
            return;
        }
        const double th = atan2(r, z), t2 = th * th;
        const double thd = th * (1.0 + t2 * (p.d[0] + t2 * (p.d[1] + t2 * (p.d[2] + t2 * p.d[3]))));
        const double dthd = 1.0 + t2 * (3.0 * p.d[0] + t2 * (5.0 * p.d[1] + t2 * (7.0 * p.d[2] + t2 * 9.0 * p.d[3])));
        const double irho2 = 1.0 / (r2 + z * z), ir = 1.0 / r;
        const double cx_ = x * ir, cy_ = y * ir;                 // unit vector in the image plane
        // m = thd * (cx_, cy_):  dm/d(x,y) = dthd * th_r * c c^T + (thd / r) * (I - c c^T),  th_r = z / rho^2,  th_z = -r / rho^2
        const double radial = dthd * z * irho2, tang = thd * ir, dz = -dthd * r * irho2;
        ju[0] = p.fx * (radial * cx_ * cx_ + tang * cy_ * cy_);
        ju[1] = p.fx * (radial - tang) * cx_ * cy_;
        ju[2] = p.fx * dz * cx_;
        jv[0] = p.fy * (radial - tang) * cx_ * cy_;
        jv[1] = p.fy * (radial * cy_ * cy_ + tang * cx_ * cx_);
        jv[2] = p.fy * dz * cy_;
    }
};

// unified family: u - cx = fx * x / den  =>  grad u = fx * (e_x - mx * grad den) / den
template <int M> struct UnifiedPointJac {
    static __device__ __forceinline__ void eval(const LinParams& p, double x, double y, double z, double* ju, double* jv) {
        const double alpha = p.d[0], oma = 1.0 - alpha;
        double den, gx, gy, gz;
        if (M == ACM_MODEL_UCM) {
            const double d = sqrt(x * x + y * y + z * z), s = alpha / d;
            den = alpha * d + oma * z;
            gx = s * x; gy = s * y; gz = s * z + oma;
        } else if (M == ACM_MODEL_EUCM) {
            const double beta = p.d[1];
            const double d = sqrt(beta * (x * x + y * y) + z * z), s = alpha / d;
            den = alpha * d + oma * z;
            gx = s * beta * x; gy = s * beta * y; gz = s * z + oma;
        } else {
            const double xi = p.d[1];
            const double rr = x * x + y * y;
            const double d1 = sqrt(rr + z * z), g = xi * d1 + z, d2 = sqrt(rr + g * g);
            den = alpha * d2 + oma * g;
            const double e = xi / d1;                         // grad g = e * X + e_z
            const double w = alpha * g / d2 + oma;            // grad den = (alpha / d2) (x, y, 0) + w * grad g
            const double a2 = alpha / d2;
            gx = a2 * x + w * e * x; gy = a2 * y + w * e * y; gz = w * (e * z + 1.0);
        }
        const double id = 1.0 / den, mx = x * id, my = y * id;
        const double fu = p.fx * id, fv = p.fy * id;
        ju[0] = fu * (1.0 - mx * gx); ju[1] = -fu * mx * gy; ju[2] = -fu * mx * gz;
        jv[0] = -fv * my * gx; jv[1] = fv * (1.0 - my * gy); jv[2] = -fv * my * gz;
    }
};
template <> struct PointJac<ACM_MODEL_UCM> : UnifiedPointJac<ACM_MODEL_UCM> {};
template <> struct PointJac<ACM_MODEL_EUCM> : UnifiedPointJac<ACM_MODEL_EUCM> {};
template <> struct PointJac<ACM_MODEL_DOUBLE_SPHERE> : UnifiedPointJac<ACM_MODEL_DOUBLE_SPHERE> {};

template <> struct PointJac<ACM_MODEL_FOV> {
    static __device__ __forceinline__ void eval(const LinParams& p, double x, double y, double z, double* ju, double* jv) {
        const double w = p.d[0], t = p.k0;   // k0 = tan(w/2)
        const double r2 = x * x + y * y;
        if (r2 < LIN_SQRT_EPS) {   // the reference's near-axis branch mx = x * (2t / w): the derivative of what project() evaluates
            const double rd = 2.0 * t / w;
            ju[0] = p.fx * rd; ju[1] = ju[2] = 0.0; jv[0] = jv[2] = 0.0; jv[1] = p.fy * rd;
            return;
        }
        const double r = sqrt(r2), s = 2.0 * t * r;
        const double a = atan2(s, z), iq = 1.0 / (s * s + z * z), irw = 1.0 / (r * w);
        const double rd = a * irw;
        const double rd_r = (2.0 * t * z * iq - a / r) * irw;   // d rd / d r
        const double rd_z = -s * iq * irw;                      // d rd / d z
        const double gx = rd_r * x / r, gy = rd_r * y / r;
        ju[0] = p.fx * (rd + x * gx); ju[1] = p.fx * x * gy; ju[2] = p.fx * x * rd_z;
        jv[0] = p.fy * y * gx; jv[1] = p.fy * (rd + y * gy); jv[2] = p.fy * y * rd_z;
    }
};
