/*
 * acm_oracle_image.c -- CPU ORACLE (test infrastructure only; see acm_oracle.h).
 *
 * Restatement of the reference's image-quality diagnostics:
 * reference src/util/image_quality.rs -- calculate_psnr (:45-89), calculate_ssim (:108-189),
 * rgb_to_grayscale (:194-210), compute_image_quality_metrics (:254-324), create_projection_image
 * (:338-373), create_combined_projection_image_on_reference (:389-437),
 * create_combined_projection_image (:453-505), model_projection_visualization (:553-616; the
 * drawing, not the PNG file).  Scalar loops in the reference's order; images are RGB8 interleaved
 * row-major (image::RgbImage).
 *
 * Pinning: the reference holds no test and no golden number for these functions (grep over
 * tests/ and src/: none).  The restatement is pinned by an independent float64 re-evaluation in
 * Python (tests/golden/image_quality.json, made by tests/golden/make_image_quality_golden.py) and
 * by closed-form cases (identical images, a single disc) in tests/test_oracle_image_quality.py.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "acm_oracle.h"

/* image_quality.rs:45-89.  Black pixels (all six channel values zero) are skipped; the sums run
 * over integers, so the f64 accumulation of the reference is exact below 2^53. */
double orc_image_psnr(const uint8_t* a, const uint8_t* b, uint32_t W, uint32_t H) {
    double mse = 0.0;
    uint64_t valid = 0;
    for (uint32_t y = 0; y < H; ++y)
        for (uint32_t x = 0; x < W; ++x) {
            const uint8_t* p = a + 3 * ((size_t)y * W + x);
            const uint8_t* q = b + 3 * ((size_t)y * W + x);
            if (p[0] != 0 || p[1] != 0 || p[2] != 0 || q[0] != 0 || q[1] != 0 || q[2] != 0) {
                for (int c = 0; c < 3; ++c) {
                    double diff = (double)p[c] - (double)q[c];
                    mse += diff * diff;
                }
                valid += 3;
            }
        }
    if (valid == 0) return INFINITY;
    mse /= (double)valid;
    if (mse <= 1e-10) return INFINITY;
    return 10.0 * log10(255.0 * 255.0 / mse);
}

/* image_quality.rs:194-210: (0.299 r + 0.587 g + 0.114 b) as u8 -- a truncating, saturating cast */
static uint8_t orc_luma(const uint8_t* p) {
    double g = 0.299 * (double)p[0] + 0.587 * (double)p[1] + 0.114 * (double)p[2];
    if (!(g > 0.0)) return 0;
    if (g >= 255.0) return 255;
    return (uint8_t)g;
}
void orc_rgb_to_grayscale(const uint8_t* rgb, uint32_t W, uint32_t H, uint8_t* gray) {
    for (size_t i = 0; i < (size_t)W * H; ++i) gray[i] = orc_luma(rgb + 3 * i);
}

/* image_quality.rs:108-189: 3x3 windows over the interior, sample covariance (divide by 8) */
double orc_image_ssim(const uint8_t* a, const uint8_t* b, uint32_t W, uint32_t H) {
    if (W < 3 || H < 3) return 1.0; /* empty loop ranges: count == 0 */
    uint8_t* g1 = (uint8_t*)malloc((size_t)W * H);
    uint8_t* g2 = (uint8_t*)malloc((size_t)W * H);
    orc_rgb_to_grayscale(a, W, H, g1);
    orc_rgb_to_grayscale(b, W, H, g2);
    const double t1 = 0.01 * 255.0, t2 = 0.03 * 255.0;
    const double c1 = t1 * t1, c2 = t2 * t2;
    double sum1 = 0.0;
    uint64_t count = 0;
    for (uint32_t y = 1; y < H - 1; ++y)
        for (uint32_t x = 1; x < W - 1; ++x) {
            double ls1 = 0.0, ls2 = 0.0;
            int lc = 0;
            for (int dy = -1; dy <= 1; ++dy)
                for (int dx = -1; dx <= 1; ++dx) {
                    size_t k = (size_t)(y + dy) * W + (size_t)(x + dx);
                    ls1 += (double)g1[k]; ls2 += (double)g2[k]; lc += 1;
                }
            double mu1 = ls1 / (double)lc, mu2 = ls2 / (double)lc;
            double s1 = 0.0, s2 = 0.0, s12 = 0.0;
            for (int dy = -1; dy <= 1; ++dy)
                for (int dx = -1; dx <= 1; ++dx) {
                    size_t k = (size_t)(y + dy) * W + (size_t)(x + dx);
                    double v1 = (double)g1[k], v2 = (double)g2[k];
                    s1 += (v1 - mu1) * (v1 - mu1);
                    s2 += (v2 - mu2) * (v2 - mu2);
                    s12 += (v1 - mu1) * (v2 - mu2);
                }
            s1 /= (double)(lc - 1); s2 /= (double)(lc - 1); s12 /= (double)(lc - 1);
            double numerator = (2.0 * mu1 * mu2 + c1) * (2.0 * s12 + c2);
            double denominator = (mu1 * mu1 + mu2 * mu2 + c1) * (s1 + s2 + c2);
            if (denominator > 0.0) { sum1 += numerator / denominator; count += 1; }
        }
    free(g1); free(g2);
    return count > 0 ? sum1 / (double)count : 1.0;
}

/* `projection.x.round() as i32` (image_quality.rs:353-354): round half away from zero, then a
 * saturating cast (NaN -> 0) */
static int32_t orc_round_i32(double v) {
    double r = round(v);
    if (r != r) return 0;
    if (r >= 2147483647.0) return INT32_MAX;
    if (r <= -2147483648.0) return INT32_MIN;
    return (int32_t)r;
}

/* the disc every drawing routine of the reference uses (radius 2, dx^2 + dy^2 <= 4), clipped */
void orc_draw_points_rgb8(const double* uv, size_t n, uint8_t r, uint8_t g, uint8_t b, uint8_t* img, uint32_t W, uint32_t H) {
    const int radius = 2;
    for (size_t i = 0; i < n; ++i) {
        const int64_t cx = orc_round_i32(uv[2 * i]), cy = orc_round_i32(uv[2 * i + 1]);
        for (int dy = -radius; dy <= radius; ++dy)
            for (int dx = -radius; dx <= radius; ++dx)
                if (dx * dx + dy * dy <= radius * radius) {
                    const int64_t x = cx + dx, y = cy + dy;
                    if (x >= 0 && x < (int64_t)W && y >= 0 && y < (int64_t)H) {
                        uint8_t* p = img + 3 * ((size_t)y * W + (size_t)x);
                        p[0] = r; p[1] = g; p[2] = b;
                    }
                }
    }
}

/* image_quality.rs:254-324.  `combined` (may be NULL) receives the display image: green input
 * projections, then magenta output projections, over `reference` (may be NULL = black).
 * Returns the number of kept points (0 => the reference returns ZeroProjectionPoints). */
size_t orc_image_quality_metrics(const orc_model* in, const orc_model* out, const double* xyz, size_t n, uint32_t W, uint32_t H,
                                 const uint8_t* reference, uint8_t* combined, double* psnr, double* ssim) {
    double* pin = (double*)malloc(2 * (n ? n : 1) * sizeof(double));
    double* pout = (double*)malloc(2 * (n ? n : 1) * sizeof(double));
    size_t m = 0;
    for (size_t i = 0; i < n; ++i) {
        double a[2], b[2];
        if (orc_project(in, xyz + 3 * i, a) != ORC_OK) continue;
        if (orc_project(out, xyz + 3 * i, b) != ORC_OK) continue;
        if (b[0] >= 0.0 && b[0] < (double)W && b[1] >= 0.0 && b[1] < (double)H) {
            pin[2 * m] = a[0]; pin[2 * m + 1] = a[1];
            pout[2 * m] = b[0]; pout[2 * m + 1] = b[1];
            ++m;
        }
    }
    *psnr = NAN; *ssim = NAN;
    if (m > 0) {
        const size_t bytes = (size_t)W * H * 3;
        if (combined) {
            if (reference) memcpy(combined, reference, bytes); else memset(combined, 0, bytes);
            orc_draw_points_rgb8(pin, m, 0, 255, 0, combined, W, H);
            orc_draw_points_rgb8(pout, m, 255, 0, 255, combined, W, H);
        }
        uint8_t* i1 = (uint8_t*)calloc(bytes ? bytes : 1, 1);
        uint8_t* i2 = (uint8_t*)calloc(bytes ? bytes : 1, 1);
        orc_draw_points_rgb8(pin, m, 255, 255, 255, i1, W, H);
        orc_draw_points_rgb8(pout, m, 255, 255, 255, i2, W, H);
        *psnr = orc_image_psnr(i1, i2, W, H);
        *ssim = orc_image_ssim(i1, i2, W, H);
        free(i1); free(i2);
    }
    free(pin); free(pout);
    return m;
}

/* util::validate_conversion_accuracy (reference src/util/validation.rs:93-213): five probe pixels
 * at 0.5 / 0.55 / 0.65 / 0.8 / 0.95 of (W, H) of the input model -> unproject(input) ->
 * project(both) -> |difference|.  errors[5] (NaN where a step fails); returns the number of valid
 * regions; *average = NaN when there is none. */
int orc_validate_conversion(const orc_model* out, const orc_model* in, double errors[5], double* average, double* max_error) {
    static const double frac[5] = {0.5, 0.55, 0.65, 0.8, 0.95};
    const double W = (double)in->width, H = (double)in->height;
    double total = 0.0, mx = 0.0;
    int valid = 0;
    for (int i = 0; i < 5; ++i) {
        errors[i] = NAN;
        double px[2] = {W * frac[i], H * frac[i]}, ray[3], a[2], b[2];
        if (orc_unproject(in, px, ray) != ORC_OK) continue;
        if (orc_project(in, ray, a) != ORC_OK || orc_project(out, ray, b) != ORC_OK) continue;
        const double dx = a[0] - b[0], dy = a[1] - b[1];
        const double e = sqrt(dx * dx + dy * dy);
        total += e; mx = fmax(mx, e); valid += 1; errors[i] = e;
    }
    *average = valid > 0 ? total / (double)valid : NAN;
    *max_error = mx;
    return valid;
}
