/*
 * acm_oracle_util.c -- CPU ORACLE. TEST INFRASTRUCTURE ONLY (see acm_oracle.h).
 *
 * sample_points, compute_reprojection_error, undistort_image and the deterministic synthetic
 * input generators shared (by specification, not by code) with the CUDA library.
 */
#include "acm_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ref: src/util/point_sampling.rs:46-78 */
size_t orc_sample_grid_size(const orc_model* m, size_t n_requested, int* ncx, int* ncy) {
    double width = (double)m->width, height = (double)m->height;
    int num_cells_x = (int)round(sqrt((double)n_requested * (width / height)));
    int num_cells_y = (int)round(sqrt((double)n_requested * (height / width)));
    if (ncx) *ncx = num_cells_x;
    if (ncy) *ncy = num_cells_y;
    return (size_t)((long long)num_cells_x * (long long)num_cells_y);
}

/* ref: src/util/point_sampling.rs:46-120 -- grid of cell centres, unproject, keep Ok && z > 0,
 * order preserved */
size_t orc_sample_points(const orc_model* m, size_t n_requested, double* uv_out, double* xyz_out) {
    int ncx, ncy;
    orc_sample_grid_size(m, n_requested, &ncx, &ncy);
    double width = (double)m->width, height = (double)m->height;
    double cell_width = width / (double)ncx;
    double cell_height = height / (double)ncy;
    size_t kept = 0;
    for (int i = 0; i < ncy; ++i) {
        for (int j = 0; j < ncx; ++j) {
            double p[2] = {((double)j + 0.5) * cell_width, ((double)i + 0.5) * cell_height};
            double ray[3];
            if (orc_unproject(m, p, ray) == ORC_OK && ray[2] > 0.0) {
                uv_out[2 * kept] = p[0]; uv_out[2 * kept + 1] = p[1];
                xyz_out[3 * kept] = ray[0]; xyz_out[3 * kept + 1] = ray[1]; xyz_out[3 * kept + 2] = ray[2];
                ++kept;
            }
        }
    }
    return kept;
}

static int cmp_double(const void* a, const void* b) {
    double x = *(const double*)a, y = *(const double*)b;
    return (x > y) - (x < y);
}

/* ref: src/util/error_metrics.rs:62-121 */
int orc_reprojection_error(const orc_model* m, const double* xyz, const double* uv, size_t n, orc_proj_error* out) {
    double* errors = (double*)malloc(sizeof(double) * (n ? n : 1));
    size_t cnt = 0;
    for (size_t i = 0; i < n; ++i) {
        double p[2];
        if (orc_project(m, xyz + 3 * i, p) == ORC_OK) {
            double dx = p[0] - uv[2 * i], dy = p[1] - uv[2 * i + 1];
            errors[cnt++] = sqrt(dx * dx + dy * dy);
        }
    }
    out->count = cnt;
    if (cnt == 0) { free(errors); return -1; } /* UtilError::ZeroProjectionPoints */
    double nn = (double)cnt, sum = 0.0;
    for (size_t i = 0; i < cnt; ++i) sum += errors[i];
    double mean = sum / nn, var = 0.0, sq = 0.0;
    for (size_t i = 0; i < cnt; ++i) { double d = errors[i] - mean; var += d * d; }
    var /= nn;
    for (size_t i = 0; i < cnt; ++i) sq += errors[i] * errors[i];
    double mn = INFINITY, mx = -INFINITY;
    for (size_t i = 0; i < cnt; ++i) { mn = fmin(mn, errors[i]); mx = fmax(mx, errors[i]); }
    qsort(errors, cnt, sizeof(double), cmp_double);
    double median = (cnt % 2 == 0) ? (errors[cnt / 2 - 1] + errors[cnt / 2]) / 2.0 : errors[cnt / 2];
    out->mean = mean; out->stddev = sqrt(var); out->rmse = sqrt(sq / nn); out->min = mn; out->max = mx; out->median = median;
    free(errors);
    return 0;
}

/* ref: src/util/undistort.rs:33-46 (ray through the target pinhole, then camera_model.project) */
void orc_undistort_map(const orc_model* m, const double target[4], double* src_xy) {
    const unsigned W = m->width, H = m->height;
    for (unsigned v_out = 0; v_out < H; ++v_out)
        for (unsigned u_out = 0; u_out < W; ++u_out) {
            double ray[3] = {((double)u_out - target[2]) / target[0], ((double)v_out - target[3]) / target[1], 1.0};
            double p[2];
            orc_project(m, ray, p); /* NaN on failure */
            src_xy[2 * ((size_t)v_out * W + u_out)] = p[0];
            src_xy[2 * ((size_t)v_out * W + u_out) + 1] = p[1];
        }
}

/* ref: src/util/undistort.rs:51-105 */
static int interpolate_pixel(const uint8_t* img, unsigned W, unsigned H, double x, double y, int interp, uint8_t rgb[3]) {
    if (interp == 0) {
        /* x.round() as i32: saturating cast, NaN -> 0 */
        double rx = round(x), ry = round(y);
        int u = isnan(rx) ? 0 : (rx >= 2147483647.0 ? 2147483647 : (rx <= -2147483648.0 ? (-2147483647 - 1) : (int)rx));
        int v = isnan(ry) ? 0 : (ry >= 2147483647.0 ? 2147483647 : (ry <= -2147483648.0 ? (-2147483647 - 1) : (int)ry));
        if (u >= 0 && u < (int)W && v >= 0 && v < (int)H) {
            const uint8_t* p = img + 3 * ((size_t)v * W + (size_t)u);
            rgb[0] = p[0]; rgb[1] = p[1]; rgb[2] = p[2];
            return 1;
        }
        return 0;
    }
    double x0 = floor(x), y0 = floor(y);
    double x1 = x0 + 1.0, y1 = y0 + 1.0;
    if (x0 < 0.0 || x1 >= (double)W || y0 < 0.0 || y1 >= (double)H) return 0;
    if (isnan(x0) || isnan(y0)) return 0; /* unreachable from undistort (project Ok => finite or handled) */
    unsigned x0u = (unsigned)x0, y0u = (unsigned)y0, x1u = (unsigned)x1, y1u = (unsigned)y1;
    const uint8_t* p00 = img + 3 * ((size_t)y0u * W + x0u);
    const uint8_t* p10 = img + 3 * ((size_t)y0u * W + x1u);
    const uint8_t* p01 = img + 3 * ((size_t)y1u * W + x0u);
    const uint8_t* p11 = img + 3 * ((size_t)y1u * W + x1u);
    double wx = x - x0, wy = y - y0, wx_inv = 1.0 - wx, wy_inv = 1.0 - wy;
    for (int c = 0; c < 3; ++c) {
        double val = (double)p00[c] * wx_inv * wy_inv + (double)p10[c] * wx * wy_inv + (double)p01[c] * wx_inv * wy + (double)p11[c] * wx * wy;
        double r = round(val);
        r = r < 0.0 ? 0.0 : (r > 255.0 ? 255.0 : r);
        rgb[c] = (uint8_t)r;
    }
    return 1;
}

int orc_undistort_rgb8(const orc_model* m, const double target[4], const uint8_t* in, uint8_t* out, int interp, int nthreads) {
    const unsigned W = m->width, H = m->height;
    (void)nthreads;
    memset(out, 0, (size_t)3 * W * H); /* RgbImage::new => black */
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(nthreads > 0 ? nthreads : 1)
#endif
    for (long long v_out = 0; v_out < (long long)H; ++v_out)
        for (unsigned u_out = 0; u_out < W; ++u_out) {
            double ray[3] = {((double)u_out - target[2]) / target[0], ((double)v_out - target[3]) / target[1], 1.0};
            double p[2];
            if (orc_project(m, ray, p) == ORC_OK) {
                uint8_t rgb[3];
                /* a NaN coordinate fails every comparison in the bilinear guard and would index
                 * garbage in Rust's `as u32` (saturates to 0); treat as no sample */
                if (isnan(p[0]) || isnan(p[1])) continue;
                if (interpolate_pixel(in, W, H, p[0], p[1], interp, rgb)) {
                    uint8_t* o = out + 3 * ((size_t)v_out * W + u_out);
                    o[0] = rgb[0]; o[1] = rgb[1]; o[2] = rgb[2];
                }
            }
        }
    return 0;
}

/* ------------------------------------------------------------------ synthetic inputs -- */
uint64_t orc_splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ULL;
    uint64_t z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
static double unit(uint64_t seed, uint64_t i, uint64_t k) {
    return (double)(orc_splitmix64(seed + 3ULL * i + k) >> 11) * 0x1.0p-53;
}

/* Direction in a cone about +z with cos(theta) uniform in [cos_max, 1]; azimuth from a rational
 * parametrisation (no sin/cos so that host and device agree bit for bit); depth in [0.5, 10).
 * adversarial != 0 replaces every 64th point by an edge case. */
void orc_synth_points3(uint64_t seed, size_t i0, size_t n, double cos_max, int adversarial, double* xyz) {
    for (size_t k = 0; k < n; ++k) {
        uint64_t i = (uint64_t)(i0 + k);
        double u0 = unit(seed, i, 0), u1 = unit(seed, i, 1), u2 = unit(seed, i, 2);
        double c = 1.0 - u0 * (1.0 - cos_max);
        double s = sqrt((1.0 - c) * (1.0 + c));
        double t = 4.0 * u1;
        int q = (int)t;
        double f = t - (double)q;
        double a = 1.0 - f, b = f;
        double nrm = sqrt(a * a + b * b);
        a = a / nrm; b = b / nrm;
        double ca, sa;
        switch (q & 3) { case 0: ca = a; sa = b; break; case 1: ca = -b; sa = a; break; case 2: ca = -a; sa = -b; break; default: ca = b; sa = -a; break; }
        double rho = 0.5 + 9.5 * u2;
        double rs = rho * s;
        double X = rs * ca, Y = rs * sa, Z = rho * c;
        if (adversarial && (i & 63ULL) == 63ULL) {
            switch ((i >> 6) % 6ULL) {
            case 0: X = 0.0; Y = 0.0; Z = 0.0; break;
            case 1: X = 0.1; Y = 0.2; Z = -1.0; break;
            case 2: X = 0.0; Y = 0.0; Z = 1e-9; break;
            case 3: X = 0.0; Y = 0.0; Z = 1.0; break;
            case 4: X = 1e-3; Y = 0.0; Z = 0x1.0p-26; break;              /* z == sqrt(EPS): valid */
            default: X = 1e-3; Y = 0.0; Z = 0x1.fffffffffffffp-27; break; /* one ulp below: centre */
            }
        }
        xyz[3 * k] = X; xyz[3 * k + 1] = Y; xyz[3 * k + 2] = Z;
    }
}

void orc_synth_pixels(uint64_t seed, size_t i0, size_t n, double W, double H, double* uv) {
    for (size_t k = 0; k < n; ++k) {
        uint64_t i = (uint64_t)(i0 + k);
        uv[2 * k] = W * unit(seed, i, 0);
        uv[2 * k + 1] = H * unit(seed, i, 1);
    }
}

void orc_synth_bytes(uint64_t seed, size_t i0, size_t n, uint8_t* out) {
    for (size_t k = 0; k < n; ++k) out[k] = (uint8_t)(orc_splitmix64(seed + (uint64_t)(i0 + k)) & 0xFF);
}
