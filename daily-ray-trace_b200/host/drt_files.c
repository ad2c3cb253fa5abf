/*
 * host/drt_files.c -- the output formats of the render path.
 *
 *   .spd  40-byte spd_file_header (daily_ray_trace.h:59-68) + W*H records of f64, row-major from the film's
 *         bottom-left: output.spd N+1 values (SPD sum, filter sum; has_filter=1), average.spd / variance.spd
 *         N values (writer daily_ray_trace.c:667-680,758-770; variance is divided by its per-pixel maximum,
 *         :766-769, giving NaN for zero-variance pixels, Q17)
 *   .bmp  14+40-byte headers, 32 bpp BI_RGB, bottom-up, BGRA, 3780 px/m, biPlanes left 0
 *         (win32_platform.c:11-41)
 * The device hands back f32 film planes; they are widened to the file's f64 here.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "drt_host.h"
#include "drt_host_internal.h"

static FILE *open_out(const char *path)
{
    char full[1024];
    drt_join_path(full, sizeof(full), NULL, path);
    return fopen(full, "wb");
}

static void fill_header(drt_spd_header *h, const drt_tables *t, uint32_t w, uint32_t hgt, uint32_t has_filter)
{
    memset(h, 0, sizeof(*h));
    h->id = DRT_SPD_FILE_ID;
    h->width = w; h->height = hgt;
    h->num_wavelengths = (uint32_t)t->num_wavelengths;
    h->has_filter = has_filter;
    h->min_wl = t->min_wl;
    h->wl_interval = t->wl_interval;
}

int drt_write_spd_sum(const char *path, const drt_tables *t, uint32_t w, uint32_t h, const float *sum, const float *filter)
{
    FILE *f = open_out(path);
    if(!f) return drt_fail(DRT_E_IO, "cannot write '%s'", path);
    drt_spd_header hdr; fill_header(&hdr, t, w, h, 1);
    size_t n = (size_t)t->num_wavelengths, pixels = (size_t)w * h;
    double *rec = (double *)malloc((n + 1) * sizeof(double));
    int ok = fwrite(&hdr, sizeof(hdr), 1, f) == 1;
    for(size_t p = 0; ok && p < pixels; p += 1)
    {
        for(size_t i = 0; i < n; i += 1) rec[i] = (double)sum[p * n + i];
        rec[n] = (double)filter[p];
        ok = fwrite(rec, sizeof(double), n + 1, f) == n + 1;
    }
    free(rec);
    fclose(f);
    return ok ? DRT_OK : drt_fail(DRT_E_IO, "short write to '%s'", path);
}

int drt_write_spd_plain(const char *path, const drt_tables *t, uint32_t w, uint32_t h, const float *values, int normalise_per_pixel)
{
    FILE *f = open_out(path);
    if(!f) return drt_fail(DRT_E_IO, "cannot write '%s'", path);
    drt_spd_header hdr; fill_header(&hdr, t, w, h, 0);
    size_t n = (size_t)t->num_wavelengths, pixels = (size_t)w * h;
    double *rec = (double *)malloc(n * sizeof(double));
    int ok = fwrite(&hdr, sizeof(hdr), 1, f) == 1;
    for(size_t p = 0; ok && p < pixels; p += 1)
    {
        for(size_t i = 0; i < n; i += 1) rec[i] = (double)values[p * n + i];
        if(normalise_per_pixel)
        {
            double peak = 0.0;
            for(size_t i = 0; i < n; i += 1) if(rec[i] > peak) peak = rec[i];
            for(size_t i = 0; i < n; i += 1) rec[i] /= peak;
        }
        ok = fwrite(rec, sizeof(double), n, f) == n;
    }
    free(rec);
    fclose(f);
    return ok ? DRT_OK : drt_fail(DRT_E_IO, "short write to '%s'", path);
}

int drt_spd_to_rgb(const char *path, const drt_tables *t, uint32_t *w, uint32_t *h, double **rgb_out)
{
    char full[1024];
    drt_join_path(full, sizeof(full), NULL, path);
    FILE *f = fopen(full, "rb");
    if(!f) return drt_fail(DRT_E_IO, "cannot read '%s'", path);
    drt_spd_header hdr;
    if(fread(&hdr, sizeof(hdr), 1, f) != 1 || hdr.id != DRT_SPD_FILE_ID || (int32_t)hdr.num_wavelengths != t->num_wavelengths)
    {
        fclose(f);
        return drt_fail(DRT_E_PARSE, "'%s' is not an .spd file with %d wavelengths", path, t->num_wavelengths);
    }
    size_t n = hdr.num_wavelengths, pixels = (size_t)hdr.width * hdr.height;
    size_t rec_len = hdr.has_filter ? n + 1 : n;
    double *rec = (double *)malloc(rec_len * sizeof(double));
    double *rgb = (double *)malloc(pixels * 3 * sizeof(double));
    int ok = 1;
    for(size_t p = 0; ok && p < pixels; p += 1)
    {
        ok = fread(rec, sizeof(double), rec_len, f) == rec_len;
        if(hdr.has_filter) for(size_t i = 0; i < n; i += 1) rec[i] = rec[i] / rec[n];
        drt_spectrum_to_rgb(t, rec, &rgb[p * 3]);
    }
    free(rec);
    fclose(f);
    if(!ok) { free(rgb); return drt_fail(DRT_E_IO, "'%s' is truncated", path); }
    *w = hdr.width; *h = hdr.height; *rgb_out = rgb;
    return DRT_OK;
}

int drt_write_bmp(const char *path, uint32_t w, uint32_t h, const uint32_t *bgra)
{
    FILE *f = open_out(path);
    if(!f) return drt_fail(DRT_E_IO, "cannot write '%s'", path);
    uint32_t pixel_bytes = w * h * 4u, off = 14u + 40u, total = off + pixel_bytes;
    unsigned char hdr[54];
    memset(hdr, 0, sizeof(hdr));
    hdr[0] = 'B'; hdr[1] = 'M';
    memcpy(hdr + 2, &total, 4);
    memcpy(hdr + 10, &off, 4);
    uint32_t info = 40, ppm = 3780; uint16_t bits = 32;
    memcpy(hdr + 14, &info, 4);
    memcpy(hdr + 18, &w, 4);
    memcpy(hdr + 22, &h, 4);
    memcpy(hdr + 28, &bits, 2);      /* biPlanes (offset 26) stays 0 as in the reference, Q23 */
    memcpy(hdr + 38, &ppm, 4);
    memcpy(hdr + 42, &ppm, 4);
    int ok = fwrite(hdr, 1, sizeof(hdr), f) == sizeof(hdr) && fwrite(bgra, 4, (size_t)w * h, f) == (size_t)w * h;
    fclose(f);
    return ok ? DRT_OK : drt_fail(DRT_E_IO, "short write to '%s'", path);
}

int drt_write_bmp_rgb(const char *path, uint32_t w, uint32_t h, const double *rgb)
{
    size_t pixels = (size_t)w * h;
    uint32_t *q = (uint32_t *)malloc(pixels * 4);
    for(size_t p = 0; p < pixels; p += 1) q[p] = drt_rgb_to_bgra8(&rgb[p * 3]);
    int rc = drt_write_bmp(path, w, h, q);
    free(q);
    return rc;
}
