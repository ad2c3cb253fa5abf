/*
 * host/drt_spectra.c -- one-off f64 spectral setup on the host.
 *
 *   drt_load_csv_spectrum    <- load_csv_file_to_spectrum   read_scene.c:801-872
 *   drt_load_tables          <- init_spd_tables             spectrum.c:1-47
 *   drt_rgb_to_spectrum      <- rgb_f64_to_spectrum         spectrum.c:84-119
 *   drt_blackbody_spectrum   <- generate_blackbody_spectrum spectrum.c:245-273 (long double)
 *   drt_spectrum_to_xyz/rgb  <- spectrum_to_xyz/rgb_f64     spectrum.c:49-82
 *   drt_rgb_to_bgra8         <- rgb_f64_to_rgb_u8           win32_platform.c:136-147
 * These run once per scene; any drift here would tint every pixel, so the arithmetic keeps the
 * reference's operation order and types and is pinned bit-for-bit against oracle/_ref in tests/.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "drt_host.h"
#include "drt_host_internal.h"

void drt_join_path(char *dst, size_t cap, const char *root_dir, const char *path)
{
    if(root_dir && root_dir[0] && path[0] != '/') snprintf(dst, cap, "%s/%s", root_dir, path);
    else snprintf(dst, cap, "%s", path);
    for(char *c = dst; *c; c += 1) if(*c == '\\') *c = '/';
}

char *drt_read_text_file(const char *root_dir, const char *path, size_t *size)
{
    char full[1024];
    drt_join_path(full, sizeof(full), root_dir, path);
    FILE *f = fopen(full, "rb");
    if(!f) return NULL;
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    char *buf = (char *)calloc((size_t)n + 2, 1);
    if(!buf || fread(buf, 1, (size_t)n, f) != (size_t)n) { fclose(f); free(buf); return NULL; }
    fclose(f);
    if(size) *size = (size_t)n;
    return buf;
}

/* utils.c:1-4 */
static double lerp(double x, double x0, double x1, double y0, double y1)
{
    return y0 + ((x - x0) * ((y1 - y0) / (x1 - x0)));
}

/* The reference's scanners stop at NUL or at a byte equal to (char)EOF (read_scene.c:767-799). */
static int at_stop(char c) { return c == 0 || c == (char)EOF; }
static const char *seek_newline(const char *c) { for(;; c += 1) { if(*c == '\n') return c; if(at_stop(*c)) return NULL; } }
static const char *seek_digit(const char *c)   { for(;; c += 1) { if(*c >= '0' && *c <= '9') return c; if(at_stop(*c)) return NULL; } }
static const char *seek_char(const char *c, char want) { for(;; c += 1) { if(*c == want) return c; if(at_stop(*c)) return NULL; } }

int drt_load_csv_spectrum(const char *root_dir, const char *path, int n, double min_wl, double interval, double *dst)
{
    size_t size = 0;
    char *text = drt_read_text_file(root_dir, path, &size);
    if(!text) return drt_fail(DRT_E_IO, "cannot read spectrum csv '%s'", path);

    double wl[DRT_MAX_WAVELENGTHS + 1], val[DRT_MAX_WAVELENGTHS + 1];
    memset(wl, 0, sizeof(wl));
    memset(val, 0, sizeof(val));
    /* one sample per newline that still has a digit somewhere after it; the heading line is skipped */
    int rows = 0;
    for(const char *c = seek_newline(text); c != NULL; c = seek_newline(c))
    {
        c += 1;
        if(seek_digit(c)) rows += 1;
    }
    if(rows > DRT_MAX_WAVELENGTHS) { free(text); return drt_fail(DRT_E_LIMIT, "'%s' has %d rows, more than %d", path, rows, DRT_MAX_WAVELENGTHS); }
    const char *c = seek_newline(text);
    c = c ? c + 1 : NULL;
    for(int i = 0; i < rows; i += 1)
    {
        /* sign and a leading '.' are lost, exactly as in the reference: numbers start at the next digit */
        if(!(c = seek_digit(c))) break;
        wl[i] = atof(c);
        if(!(c = seek_char(c, ','))) { free(text); return drt_fail(DRT_E_PARSE, "'%s' row %d has no comma", path, i + 1); }
        if(!(c = seek_digit(c))) { free(text); return drt_fail(DRT_E_PARSE, "'%s' row %d has no value", path, i + 1); }
        val[i] = atof(c);
        c = seek_newline(c);
        if(!c) { rows = i + 1; break; }
    }
    free(text);

    if(wl[0] < 10.0) for(int i = 0; i < rows; i += 1) wl[i] *= 1000.0;   /* looks like micrometres, :847-853 */

    int at = 0;
    for(int s = 0; s < n; s += 1)
    {
        double sample_wl = min_wl + ((double)s) * interval;
        while(wl[at + 1] < sample_wl)
        {
            at += 1;
            if(at + 1 >= rows) return drt_fail(DRT_E_LIMIT, "'%s' ends at %g nm, before %g nm", path, wl[rows > 0 ? rows - 1 : 0], sample_wl);
        }
        dst[s] = lerp(sample_wl, wl[at], wl[at + 1], val[at], val[at + 1]);
    }
    return DRT_OK;
}

int drt_load_tables(const drt_config *cfg, const char *root_dir, drt_tables *t)
{
    memset(t, 0, sizeof(*t));
    uint32_t n = (uint32_t)(((cfg->max_wl - cfg->min_wl) / cfg->wl_interval) + 1.0);   /* spectrum.c:3 */
    if(n == 0 || n > DRT_MAX_WAVELENGTHS) return drt_fail(DRT_E_LIMIT, "%u wavelengths requested, limit is %d", n, DRT_MAX_WAVELENGTHS);
    t->num_wavelengths = (int32_t)n;
    t->min_wl = cfg->min_wl;
    t->wl_interval = cfg->wl_interval;
    struct { double *dst; const char *path; } jobs[] =
    {
        { t->ref_white, cfg->white_spd }, { t->cmf_x, cfg->cmf_x }, { t->cmf_y, cfg->cmf_y }, { t->cmf_z, cfg->cmf_z },
        { t->rgb_basis[0], cfg->white_spd }, { t->rgb_basis[1], cfg->red_spd }, { t->rgb_basis[2], cfg->green_spd },
        { t->rgb_basis[3], cfg->blue_spd }, { t->rgb_basis[4], cfg->cyan_spd }, { t->rgb_basis[5], cfg->magenta_spd },
        { t->rgb_basis[6], cfg->yellow_spd },
    };
    for(size_t i = 0; i < sizeof(jobs) / sizeof(jobs[0]); i += 1)
    {
        int rc = drt_load_csv_spectrum(root_dir, jobs[i].path, (int)n, cfg->min_wl, cfg->wl_interval, jobs[i].dst);
        if(rc != DRT_OK) return rc;
    }
    return DRT_OK;
}

void drt_rgb_to_spectrum(const drt_tables *t, const double rgb[3], double *dst)
{
    const double *primary[3]   = { t->rgb_basis[1], t->rgb_basis[2], t->rgb_basis[3] };   /* red, green, blue */
    const double *secondary[3] = { t->rgb_basis[4], t->rgb_basis[5], t->rgb_basis[6] };   /* cyan, magenta, yellow */
    int order[3] = { 0, 1, 2 }, tmp;
    /* three-compare sort of the channel indices, smallest first (spectrum.c:92-109) */
    if(rgb[order[0]] > rgb[order[1]]) { tmp = order[1]; order[1] = order[0]; order[0] = tmp; }
    if(rgb[order[1]] > rgb[order[2]]) { tmp = order[2]; order[2] = order[1]; order[1] = tmp; }
    if(rgb[order[0]] > rgb[order[1]]) { tmp = order[1]; order[1] = order[0]; order[0] = tmp; }
    int lo = order[0], mid = order[1], hi = order[2];
    double mid_minus_lo = rgb[mid] - rgb[lo];
    double hi_minus_mid = rgb[hi] - rgb[mid];
    int n = t->num_wavelengths;
    for(int i = 0; i < n; i += 1) dst[i]  = t->rgb_basis[0][i] * rgb[lo];
    for(int i = 0; i < n; i += 1) dst[i] += secondary[lo][i] * mid_minus_lo;
    for(int i = 0; i < n; i += 1) dst[i] += primary[hi][i] * hi_minus_mid;
}

void drt_blackbody_spectrum(const drt_tables *t, double temperature, double *dst)
{
    const long double c = 2.99792458e8L, h = 6.626176e-34L, k = 1.380662e-23L;
    long double temp = (long double)temperature;
    for(int i = 0; i < t->num_wavelengths; i += 1)
    {
        long double nm = (long double)(t->min_wl + (i * t->wl_interval));
        long double m = nm * 1e-9L;
        long double numerator = 2.0L * DRT_PI_L * h * c * c;
        long double lambda_5 = powl(m, 5.0L);
        long double e_power = ((h * c) / k) / (temp * m);
        long double denominator = lambda_5 * (expl(e_power) - 1.0L);
        dst[i] = (double)((numerator / denominator) * 1e9L);
    }
}

void drt_spectrum_to_xyz(const drt_tables *t, const double *spd, double xyz[3])
{
    int n = t->num_wavelengths;
    double norm = 0.0;
    for(int i = 0; i < n; i += 1) norm += (t->cmf_y[i] * t->ref_white[i]);
    norm *= t->wl_interval;
    double x = 0.0, y = 0.0, z = 0.0;
    for(int i = 0; i < n; i += 1)
    {
        x += (t->cmf_x[i] * spd[i] * t->ref_white[i]);
        y += (t->cmf_y[i] * spd[i] * t->ref_white[i]);
        z += (t->cmf_z[i] * spd[i] * t->ref_white[i]);
    }
    xyz[0] = x * (t->wl_interval / norm);
    xyz[1] = y * (t->wl_interval / norm);
    xyz[2] = z * (t->wl_interval / norm);
}

void drt_spectrum_to_rgb(const drt_tables *t, const double *spd, double rgb[3])
{
    double q[3];
    drt_spectrum_to_xyz(t, spd, q);
    rgb[0] = (2.3706743 * q[0]) - (0.9000405 * q[1]) - (0.4706338 * q[2]);
    rgb[1] = (-0.5138850 * q[0]) + (1.4253036 * q[1]) + (0.0885814 * q[2]);
    rgb[2] = (0.0052982 * q[0]) - (0.0146949 * q[1]) + (1.0093968 * q[2]);
}

uint32_t drt_rgb_to_bgra8(const double rgb[3])
{
    uint32_t out = 0;
    for(int ch = 0; ch < 3; ch += 1)
    {
        double v = rgb[ch];
        if(v != v) v = 0.0;                               /* NaN (0/0 variance pixels, Q17): x86 cvttsd2si truncates to byte 0 */
        v = (v < 0.0) ? 0.0 : v;
        v = (v > 1.0) ? 1.0 : v;
        uint32_t q = (uint32_t)(uint8_t)(v * 255.0);     /* truncation, win32_platform.c:142-144 */
        out |= q << (8 * (2 - ch));                      /* memory order b,g,r,a (types.h:37-47) */
    }
    return out;
}
