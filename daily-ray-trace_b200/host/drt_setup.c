/*
 * host/drt_setup.c -- turns parsed input into the flattened drt_scene / drt_camera.
 *
 *   camera      <- init_camera   daily_ray_trace.c:49-77   (find_rotation_between_vectors geometry.c:263-295,
 *                                                           rotation_about_axis geometry.c:297-313 incl. its
 *                                                           axis.z*axis.z term at [0].z)
 *   materials   <- init_scene    daily_ray_trace.c:125-176 (+ init_spd :79-123); processes num_materials+1
 *                                                           entries, the last one zero-filled (Q18)
 *   surfaces    <- init_scene    daily_ray_trace.c:178-210 (create_plane_from_points geometry.c:203-209;
 *                                                           material = first name match, else index 0)
 * One-off f64 work; operation order follows the reference so the result can be compared bit-for-bit.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "drt_host.h"
#include "drt_host_internal.h"

typedef struct { double c[3][3]; } m33;   /* c[column][row], as geometry.h:32-35 */

static double dot3(const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static void cross3(const double *a, const double *b, double *o)
{
    double x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
    o[0] = x; o[1] = y; o[2] = z;
}
static void normalise3(const double *v, double *o)
{
    double len = sqrt(dot3(v, v));
    o[0] = v[0] / len; o[1] = v[1] / len; o[2] = v[2] / len;
}
static void m33_row(const m33 *m, int r, double *o) { o[0] = m->c[0][r]; o[1] = m->c[1][r]; o[2] = m->c[2][r]; }
static void m33_apply(const m33 *m, const double *v, double *o)
{
    double w[3], row[3];
    for(int i = 0; i < 3; i += 1) { m33_row(m, i, row); w[i] = dot3(row, v); }
    o[0] = w[0]; o[1] = w[1]; o[2] = w[2];
}

/* geometry.c:263-295, including the quirk that mat3x3_mul writes element (i,j) into column i (harmless:
 * the operand is skew-symmetric) and that antiparallel inputs give -I (Q14). */
static m33 rotation_between(const double *v, const double *w)
{
    double n[3]; cross3(v, w, n);
    double c = dot3(v, w);
    m33 r = {{{1.0, 0.0, 0.0}, {0.0, 1.0, 0.0}, {0.0, 0.0, 1.0}}};
    if(dot3(n, n) == 0.0 && c <= 0.0)
    {
        r.c[0][0] = -1.0; r.c[1][1] = -1.0; r.c[2][2] = -1.0;
        return r;
    }
    m33 m;
    m.c[0][0] = 0.0;   m.c[0][1] = n[2];  m.c[0][2] = -n[1];
    m.c[1][0] = -n[2]; m.c[1][1] = 0.0;   m.c[1][2] = n[0];
    m.c[2][0] = n[1];  m.c[2][1] = -n[0]; m.c[2][2] = 0.0;
    m33 mm;
    for(int i = 0; i < 3; i += 1)
        for(int j = 0; j < 3; j += 1)
        {
            double row[3]; m33_row(&m, i, row);
            mm.c[i][j] = dot3(row, m.c[j]);
        }
    double f = (1.0 / (1.0 + c));
    for(int i = 0; i < 3; i += 1) for(int j = 0; j < 3; j += 1) mm.c[i][j] = f * mm.c[i][j];
    m33 id = {{{1.0, 0.0, 0.0}, {0.0, 1.0, 0.0}, {0.0, 0.0, 1.0}}};
    for(int i = 0; i < 3; i += 1) for(int j = 0; j < 3; j += 1) r.c[i][j] = (id.c[i][j] + m.c[i][j]) + mm.c[i][j];
    return r;
}

static m33 rotation_about(const double *a, double angle)
{
    double ct = cos(angle), st = sin(angle);
    m33 r;
    r.c[0][0] = ct + (a[0] * a[0]) * (1 - ct);
    r.c[0][1] = a[1] * a[0] * (1 - ct) + a[2] * st;
    r.c[0][2] = a[2] * a[2] * (1 - ct) - a[1] * st;       /* sic: the reference multiplies axis.z by axis.z here */
    r.c[1][0] = a[0] * a[1] * (1 - ct) - a[2] * st;
    r.c[1][1] = ct + (a[1] * a[1]) * (1 - ct);
    r.c[1][2] = a[2] * a[1] * (1 - ct) + a[0] * st;
    r.c[2][0] = a[0] * a[2] * (1 - ct) + a[1] * st;
    r.c[2][1] = a[1] * a[2] * (1 - ct) - a[0] * st;
    r.c[2][2] = ct + a[2] * a[2] * (1 - ct);
    return r;
}

static void build_camera(const drt_camera_input *in, uint32_t width_px, uint32_t height_px, drt_camera *cam)
{
    const double ref_forward[3] = {0.0, 0.0, -1.0}, ref_up[3] = {0.0, 1.0, 0.0};
    memset(cam, 0, sizeof(*cam));
    double d[3] = { in->target[0] - in->position[0], in->target[1] - in->position[1], in->target[2] - in->position[2] };
    normalise3(d, cam->forward);
    m33 orient = rotation_between(ref_forward, cam->forward);
    double roll_rad = in->roll * (DRT_PI_L / 180.0);
    m33 roll = rotation_about(cam->forward, roll_rad);
    double up0[3];
    m33_apply(&orient, ref_up, up0);
    m33_apply(&roll, up0, cam->up);
    double fr[3]; cross3(cam->forward, cam->up, fr);
    normalise3(fr, cam->right);

    cam->focal_depth = in->fdepth;
    cam->focal_length = in->flength;
    double aperture_distance = (in->flength * in->fdepth) / (in->flength + in->fdepth);
    for(int i = 0; i < 3; i += 1) cam->aperture_position[i] = in->position[i] + aperture_distance * cam->forward[i];
    cam->aperture_radius = in->aperture;

    double fov_rad = in->fov * (DRT_PI_L / 180.0);
    double aspect = (double)width_px / (double)height_px;
    double film_w = 2.0 * aperture_distance * tan(fov_rad / 2.0);
    double film_h = film_w / aspect;
    for(int i = 0; i < 3; i += 1)
    {
        double right_i = (0.5 * film_w) * cam->right[i];
        double top_i   = (0.5 * film_h) * cam->up[i];
        cam->film_bottom_left[i] = (in->position[i] - right_i) - top_i;
    }
    cam->pixel_width  = film_w / (double)width_px;
    cam->pixel_height = film_h / (double)height_px;

    const double plus_z[3] = {0.0, 0.0, 1.0};     /* lens disc frame, daily_ray_trace.c:591,595 */
    m33 lens = rotation_between(plus_z, cam->forward);
    for(int i = 0; i < 3; i += 1) for(int j = 0; j < 3; j += 1) cam->lens_rotation[i * 3 + j] = lens.c[i][j];
}

/* init_spd, daily_ray_trace.c:79-123 */
static int build_spd(const drt_spd_input *in, const drt_tables *t, const char *root_dir, double *dst, int *given)
{
    int n = t->num_wavelengths;
    *given = 1;
    switch(in->method)
    {
        case DRT_SPD_METHOD_RGB: drt_rgb_to_spectrum(t, in->rgb, dst); break;
        case DRT_SPD_METHOD_CSV:
        {
            char path[128];
            snprintf(path, sizeof(path), "spectra/%s", in->csv);
            int rc = drt_load_csv_spectrum(root_dir, path, n, t->min_wl, t->wl_interval, dst);
            if(rc != DRT_OK) return rc;
            break;
        }
        case DRT_SPD_METHOD_BLACKBODY:
        {
            drt_blackbody_spectrum(t, in->value, dst);
            double peak = 0.0;                                    /* spectrum_normalise, spectrum.c:182-187 */
            for(int i = 0; i < n; i += 1) if(dst[i] > peak) peak = dst[i];
            for(int i = 0; i < n; i += 1) dst[i] /= peak;
            break;
        }
        case DRT_SPD_METHOD_CONST: for(int i = 0; i < n; i += 1) dst[i] = in->value; break;
        default: *given = 0; return DRT_OK;
    }
    if(in->has_scale) for(int i = 0; i < n; i += 1) dst[i] = dst[i] * in->scale;
    return DRT_OK;
}

int drt_build_scene(const drt_scene_input *in, const drt_tables *tables, const char *root_dir,
                    uint32_t width_px, uint32_t height_px, drt_scene *scene, drt_camera *camera)
{
    if(width_px == 0 || height_px == 0) return drt_fail(DRT_E_ARG, "image size %ux%u", width_px, height_px);
    build_camera(&in->camera, width_px, height_px, camera);

    memset(scene, 0, sizeof(*scene));
    scene->num_wavelengths = tables->num_wavelengths;
    scene->min_wl = tables->min_wl;
    scene->wl_interval = tables->wl_interval;
    scene->max_wl = tables->min_wl + (tables->num_wavelengths - 1) * tables->wl_interval;
    scene->num_surfaces = in->num_surfaces;
    scene->num_materials = in->num_materials + 1;
    scene->base_material = -1;
    scene->escape_material = -1;

    static const drt_material_input phantom = { .dir_func = DRT_DIR_NONE };
    for(int i = 0; i < scene->num_materials; i += 1)
    {
        const drt_material_input *src = (i < in->num_materials) ? &in->materials[i] : &phantom;
        drt_material *dst = &scene->materials[i];
        memcpy(dst->name, src->name, sizeof(dst->name));
        dst->is_black_body = src->is_escape ? 1 : src->is_black_body;
        dst->is_emissive = src->is_emissive;
        dst->shininess = src->shininess;
        dst->roughness = src->roughness;
        dst->dir_func = (i < in->num_materials) ? src->dir_func : DRT_DIR_NONE;
        dst->num_lobes = src->num_lobes;
        for(int j = 0; j < src->num_lobes; j += 1) dst->lobes[j] = src->lobes[j];
        for(int k = 0; k < DRT_SPD_COUNT; k += 1)
        {
            int given = 0;
            int rc = build_spd(&src->spd[k], tables, root_dir, dst->spd[k], &given);
            if(rc != DRT_OK) return rc;
            if(given) dst->spd_mask |= 1 << k;
        }
        if(src->is_escape) scene->escape_material = i;
        if(src->is_base)   scene->base_material = i;
    }

    for(int i = 0; i < scene->num_surfaces; i += 1)
    {
        const drt_surface_input *src = &in->surfaces[i];
        drt_surface *dst = &scene->surfaces[i];
        memcpy(dst->name, src->name, sizeof(dst->name));
        dst->type = src->type;
        memcpy(dst->position, src->position, sizeof(dst->position));
        if(src->type == DRT_GEO_SPHERE) dst->radius = src->radius;
        if(src->type == DRT_GEO_PLANE)
        {
            for(int k = 0; k < 3; k += 1) { dst->u[k] = src->pointu[k] - src->position[k]; dst->v[k] = src->pointv[k] - src->position[k]; }
            double n[3]; cross3(dst->u, dst->v, n);
            normalise3(n, dst->normal);
        }
        dst->material = 0;
        for(int j = 0; j < scene->num_materials; j += 1)
            if(strcmp(src->material_name, scene->materials[j].name) == 0) { dst->material = j; break; }
    }

    /* What the reference leaves as a crash (NULL sample_direction, wild base/escape pointers) is an error here. */
    if(scene->base_material < 0)   return drt_fail(DRT_E_SCENE, "scene has no base_material");
    if(scene->escape_material < 0) return drt_fail(DRT_E_SCENE, "scene has no escape_material");
    for(int i = 0; i < scene->num_surfaces; i += 1)
    {
        const drt_surface *s = &scene->surfaces[i];
        const drt_material *m = &scene->materials[s->material];
        if(s->type == DRT_GEO_NONE) return drt_fail(DRT_E_SCENE, "surface '%s' has no type", s->name);
        if(s->type != DRT_GEO_POINT && !m->is_black_body && m->dir_func == DRT_DIR_NONE)
            return drt_fail(DRT_E_SCENE, "material '%s' of surface '%s' has no dir_func", m->name, s->name);
    }
    return DRT_OK;
}

int drt_load_scene_file(const char *root_dir, const char *scene_path, const drt_tables *tables, int flags,
                        uint32_t width_px, uint32_t height_px, drt_scene *scene, drt_camera *camera)
{
    size_t size = 0;
    char *text = drt_read_text_file(root_dir, scene_path, &size);
    if(!text) return drt_fail(DRT_E_IO, "cannot read scene '%s'", scene_path);
    drt_scene_input *input = (drt_scene_input *)calloc(1, sizeof(drt_scene_input));
    int rc = drt_parse_scene(text, size, flags, input);
    free(text);
    if(rc == DRT_OK) rc = drt_build_scene(input, tables, root_dir, width_px, height_px, scene, camera);
    free(input);
    return rc;
}
