/*
 * oracle/ref_perpath.c -- TEST INFRASTRUCTURE (oracle), not product code.
 *
 * Per-path driver around the UNMODIFIED reference. It #includes the reference's unity
 * header where it lies under /root/reference/src (nothing is copied into this repo) and
 * calls the reference's own functions:
 *     parse_config            read_scene.c:604
 *     init_spd_tables         spectrum.c:1
 *     load_scene              daily_ray_trace.c:30
 *     sample_scene            daily_ray_trace.c:571     <- the per-path oracle
 *     spectral_* / Welford    the sequence of daily_ray_trace.c:732-743
 *     spectrum_to_rgb_f64     spectrum.c:72
 *     rgb_f64_to_spectrum     spectrum.c:84
 * The only change in behaviour is the random source: the build passes -Drand=drt_rand
 * (the single libc call site is rng.c:4) and drt_rand() below serves the counter-based
 * per-path stream of include/drt_rng.h, re-keyed before every sample_scene call
 * (SURVEY.md Appendix C).  RAND_MAX stays glibc's 2147483647.
 *
 * Built by oracle/Makefile into oracle/_ref/libdrt_ref.so; driven from tests/ and from
 * bench.py's cpu_baseline / --impl reference legs through ctypes.
 */
#include <unistd.h>
#include "daily_ray_trace.h"
#include "drt_rng.h"

static drt_rng_stream   g_stream;
static unsigned long long g_seed        = 0;
static unsigned long long g_draw_count  = 0;
static config_arguments g_config;
static camera_data      g_camera;
static scene_data       g_scene;
static int              g_ready = 0;

int drt_rand(void)
{
    g_draw_count += 1;
    return (int)drt_rng_next31(&g_stream);
}

/* Reads the whole file NUL-terminated (is_word_keyword scans past the buffer end, read_scene.c:89). */
static char *slurp(const char *path, u32 *size)
{
    FILE *f = fopen(path, "rb");
    if(!f) return NULL;
    fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
    char *buf = calloc((size_t)n + 16, 1);
    if(fread(buf, 1, (size_t)n, f) != (size_t)n) { fclose(f); free(buf); return NULL; }
    fclose(f);
    *size = (u32)n;
    return buf;
}

/* root_dir: directory holding spectra/ and scenes/ (init_spd hard-codes "spectra\\", daily_ray_trace.c:93).
 * config_path: a config.cfg in the reference's own format ('\\' separators), relative to root_dir or absolute. */
int ref_setup(const char *root_dir, const char *config_path)
{
    if(chdir(root_dir) != 0) return -1;
    u32 size = 0;
    char *cfg = slurp(config_path, &size);
    if(!cfg) return -2;
    memset(&g_config, 0, sizeof(g_config));
    parse_config(cfg, size, &g_config);
    free(cfg);

    spd_tables_csvs csvs;
    csvs.white       = g_config.white_spd;
    csvs.cmf_x       = g_config.cmf_x;
    csvs.cmf_y       = g_config.cmf_y;
    csvs.cmf_z       = g_config.cmf_z;
    csvs.rgb_red     = g_config.red_spd;
    csvs.rgb_green   = g_config.green_spd;
    csvs.rgb_blue    = g_config.blue_spd;
    csvs.rgb_cyan    = g_config.cyan_spd;
    csvs.rgb_magenta = g_config.magenta_spd;
    csvs.rgb_yellow  = g_config.yellow_spd;
    init_spd_tables(csvs, 32, g_config.min_wl, g_config.max_wl, g_config.wl_interval);

    memset(&g_camera, 0, sizeof(g_camera));
    memset(&g_scene, 0, sizeof(g_scene));
    load_scene(g_config.input_scene, &g_camera, &g_scene, g_config.output_width, g_config.output_height);
    g_ready = 1;
    return 0;
}

void ref_set_seed(unsigned long long seed) { g_seed = seed; }
unsigned long long ref_draw_count(void) { return g_draw_count; }

unsigned ref_num_wavelengths(void) { return number_of_spectrum_samples; }
unsigned ref_width(void)           { return g_config.output_width; }
unsigned ref_height(void)          { return g_config.output_height; }
unsigned ref_spp(void)             { return g_config.num_pixel_samples; }
unsigned ref_max_depth(void)       { return g_config.max_cast_depth; }
unsigned ref_pixel_scheme(void)    { return (unsigned)g_config.pixel_scheme; }

/* One camera path with its own stream: the reference's sample_scene, verbatim. */
void ref_sample(unsigned x, unsigned y, unsigned sample, double *out_spd, double *out_filter)
{
    spectrum c; c.samples = out_spd;
    drt_rng_begin(&g_stream, g_seed, y * g_config.output_width + x, sample);
    sample_scene(c, out_filter, x, y, &g_scene, &g_camera, g_config.max_cast_depth, g_config.pixel_scheme);
}

/* Renders pixels [x0,x1)x[y0,y1), samples [s0,s1), accumulating exactly as render_image does
 * (daily_ray_trace.c:720-743).  Buffers are tile-local row-major, zero-initialised by the caller:
 *   sum  : (N+1) f64 per pixel (SPD sum, filter sum)      avg, m2 : N f64 per pixel
 *   paths: optional, N f64 per (pixel, sample) in [pixel][sample] order, or NULL */
void ref_render_tile(unsigned x0, unsigned y0, unsigned x1, unsigned y1, unsigned s0, unsigned s1,
                     double *sum, double *avg, double *m2, double *paths)
{
    u32 n = number_of_spectrum_samples;
    u32 tw = x1 - x0;
    f64 *cbuf = calloc(n + 1, sizeof(f64));
    spectrum contribution; contribution.samples = cbuf;
    f64 *filter = &cbuf[n];
    spectrum tmp_0 = alloc_spd();
    spectrum tmp_1 = alloc_spd();
    for(u32 s = s0; s < s1; s += 1)
    for(u32 y = y0; y < y1; y += 1)
    for(u32 x = x0; x < x1; x += 1)
    {
        u32 p = (y - y0) * tw + (x - x0);
        spectrum d_sum, d_avg, d_m2;
        d_sum.samples = sum + (size_t)(n + 1) * p;
        d_avg.samples = avg + (size_t)n * p;
        d_m2.samples  = m2  + (size_t)n * p;

        drt_rng_begin(&g_stream, g_seed, y * g_config.output_width + x, s);
        sample_scene(contribution, filter, x, y, &g_scene, &g_camera, g_config.max_cast_depth, g_config.pixel_scheme);
        if(paths) memcpy(paths + ((size_t)p * (s1 - s0) + (s - s0)) * n, cbuf, n * sizeof(f64));

        spectral_sum(d_sum, d_sum, contribution);
        d_sum.samples[n] += *filter;

        copy_spectrum(tmp_0, d_avg);
        spectral_sub(tmp_0, contribution, tmp_0);
        copy_spectrum(tmp_1, tmp_0);
        spectral_div_by_scalar(tmp_0, tmp_0, (f64)(s + 1));
        spectral_sum(d_avg, d_avg, tmp_0);
        spectral_sub(tmp_0, contribution, d_avg);
        spectral_mul_by_spectrum(tmp_0, tmp_1, tmp_0);
        spectral_sum(d_m2, d_m2, tmp_0);
    }
    free_spd(tmp_1);
    free_spd(tmp_0);
    free(cbuf);
}

/* ---- introspection used to pin the product's host-side parser/setup against the reference ---- */

/* forward, right, up, aperture_position (3 each), aperture_radius, focal_depth, focal_length,
 * film_bottom_left (3), pixel_width, pixel_height  = 20 doubles */
void ref_get_camera(double *out)
{
    double *o = out;
    for(int i = 0; i < 3; i += 1) *o++ = g_camera.forward.xyz[i];
    for(int i = 0; i < 3; i += 1) *o++ = g_camera.right.xyz[i];
    for(int i = 0; i < 3; i += 1) *o++ = g_camera.up.xyz[i];
    for(int i = 0; i < 3; i += 1) *o++ = g_camera.aperture_position.xyz[i];
    *o++ = g_camera.aperture_radius; *o++ = g_camera.focal_depth; *o++ = g_camera.focal_length;
    for(int i = 0; i < 3; i += 1) *o++ = g_camera.film_bottom_left.xyz[i];
    *o++ = g_camera.pixel_width; *o++ = g_camera.pixel_height;
}

unsigned ref_num_surfaces(void)  { return g_scene.num_surfaces; }
unsigned ref_num_materials(void) { return g_scene.num_scene_materials; }
int ref_base_material(void)   { return g_scene.base_material   ? (int)(g_scene.base_material   - g_scene.scene_materials) : -1; }
int ref_escape_material(void) { return g_scene.escape_material ? (int)(g_scene.escape_material - g_scene.scene_materials) : -1; }

/* type, material index; position(3), radius | normal(3), u(3), v(3) -> 13 doubles */
void ref_get_surface(unsigned i, int *type, int *material, double *geom)
{
    object_geometry *s = &g_scene.surfaces[i];
    *type = (int)s->type;
    *material = (int)g_scene.surface_material_indices[i];
    memset(geom, 0, 13 * sizeof(double));
    for(int k = 0; k < 3; k += 1) geom[k] = s->position.xyz[k];
    if(s->type == GEO_TYPE_SPHERE) geom[3] = s->radius;
    if(s->type == GEO_TYPE_PLANE)
    {
        for(int k = 0; k < 3; k += 1) { geom[4 + k] = s->normal.xyz[k]; geom[7 + k] = s->u.xyz[k]; geom[10 + k] = s->v.xyz[k]; }
    }
}

/* flags: [is_black_body, is_emissive, num_bdsfs, dir_func index (-1 none), has_spd bitmask (emission..extinct = bits 0..5)]
 * scalars: [shininess, roughness];  bdsf_ids: indices into bdsf_list.h order;  spds: 6*N doubles */
void ref_get_material(unsigned i, char *name32, int *flags, double *scalars, int *bdsf_ids, double *spds)
{
    object_material *m = &g_scene.scene_materials[i];
    memcpy(name32, m->name, 32);
    flags[0] = (int)m->is_black_body; flags[1] = (int)m->is_emissive; flags[2] = (int)m->num_bdsfs;
    flags[3] = -1;
    for(u32 k = 0; k < num_dir_funcs_defined; k += 1) if(m->sample_direction == dir_func_list[k]) flags[3] = (int)k;
    scalars[0] = m->shininess; scalars[1] = m->roughness;
    for(u32 j = 0; j < m->num_bdsfs && j < 16; j += 1)
    {
        bdsf_ids[j] = -1;
        for(u32 k = 0; k < num_bdsfs_defined; k += 1) if(m->bdsfs[j] == bdsf_list[k]) bdsf_ids[j] = (int)k;
    }
    spectrum *all[6] = { &m->emission_spd, &m->diffuse_spd, &m->glossy_spd, &m->mirror_spd, &m->refract_spd, &m->extinct_spd };
    u32 n = number_of_spectrum_samples;
    flags[4] = 0;
    for(int k = 0; k < 6; k += 1)
    {
        if(all[k]->samples) { flags[4] |= 1 << k; memcpy(spds + (size_t)k * n, all[k]->samples, n * sizeof(double)); }
        else memset(spds + (size_t)k * n, 0, n * sizeof(double));
    }
}

/* rw, x, y, z, then white, red, green, blue, cyan, magenta, yellow: 11*N doubles */
void ref_get_tables(double *out)
{
    u32 n = number_of_spectrum_samples;
    spectrum all[11] = { cmfs.rw, cmfs.x, cmfs.y, cmfs.z, rgb_spds.white, rgb_spds.red, rgb_spds.green,
                         rgb_spds.blue, rgb_spds.cyan, rgb_spds.magenta, rgb_spds.yellow };
    for(int k = 0; k < 11; k += 1) memcpy(out + (size_t)k * n, all[k].samples, n * sizeof(double));
}

void ref_spectrum_to_rgb(const double *spd, double *rgb)
{
    spectrum s; s.samples = (f64 *)spd;
    rgb_f64 c = spectrum_to_rgb_f64(s);
    rgb[0] = c.r; rgb[1] = c.g; rgb[2] = c.b;
}

void ref_rgb_to_spectrum(const double *rgb, double *spd)
{
    rgb_f64 c; c.r = rgb[0]; c.g = rgb[1]; c.b = rgb[2];
    spectrum s; s.samples = spd;
    rgb_f64_to_spectrum(c, s);
}

unsigned ref_rgb_to_u8(const double *rgb)
{
    rgb_f64 c; c.r = rgb[0]; c.g = rgb[1]; c.b = rgb[2];
    rgb_u8 q = rgb_f64_to_rgb_u8(c);
    return (unsigned)q.r | ((unsigned)q.g << 8) | ((unsigned)q.b << 16);
}

void ref_blackbody(double temperature, double *spd)
{
    spectrum s; s.samples = spd;
    generate_blackbody_spectrum(s, temperature);
}

/* The reference's bdsf() on a hand-built scene_point; used for the Q7 stale-lobe known answers. */
void ref_bdsf(unsigned surface_material, unsigned incident_material, unsigned transmit_material,
              const double *normal, const double *out_dir, const double *in_dir, double *result)
{
    scene_point p; memset(&p, 0, sizeof(p));
    for(int k = 0; k < 3; k += 1) { p.normal.xyz[k] = normal[k]; p.out.xyz[k] = out_dir[k]; }
    p.on_dot = vec3_dot(p.normal, p.out);
    p.trans_wl = 630.0;
    p.surface_material  = &g_scene.scene_materials[surface_material];
    p.incident_material = &g_scene.scene_materials[incident_material];
    p.transmit_material = &g_scene.scene_materials[transmit_material];
    vec3 in; for(int k = 0; k < 3; k += 1) in.xyz[k] = in_dir[k];
    spectrum r; r.samples = result;
    bdsf(r, &p, in);
}
