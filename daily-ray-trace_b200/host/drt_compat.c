/*
 * host/drt_compat.c -- legacy scene grammar support.
 *
 * Four of the five shipped legacy scenes (init_cornell, cornell_large_box, cornell_downward,
 * first_scene, example_scene) predate the reference's current parser: parse_camera
 * (read_scene.c:345-396) has no case for up/right/forward, their materials carry no
 * bdsfs/dir_func (-> NULL sample_direction at daily_ray_trace.c:465) and they name no
 * base_material / escape_material (-> uninitialised pointers, daily_ray_trace.c:662).
 * The reference itself cannot render them.  The rules below turn such a scene into the
 * equivalent current-grammar scene; drt_scene_write() emits it as text so that the
 * unmodified reference (oracle/_ref) renders exactly the same scene as the CUDA path.
 */
#include <math.h>
#include <stdio.h>
#include <stdarg.h>
#include <string.h>
#include "drt_host.h"
#include "drt_host_internal.h"

static double dot3(const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static void cross3(const double *a, const double *b, double *o)
{
    o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0];
}

/* The 'up' that init_camera derives at roll 0 (daily_ray_trace.c:51-57): the shortest rotation taking
 * (0,0,-1) to forward, applied to (0,1,0). */
static void derived_up_at_roll0(const double *fwd, double *up)
{
    const double ref_f[3] = {0.0, 0.0, -1.0}, ref_u[3] = {0.0, 1.0, 0.0};
    double n[3]; cross3(ref_f, fwd, n);
    double c = dot3(ref_f, fwd);
    if(dot3(n, n) == 0.0 && c <= 0.0) { up[0] = -ref_u[0]; up[1] = -ref_u[1]; up[2] = -ref_u[2]; return; }
    /* Rodrigues: u + n x u + n x (n x u) / (1 + c) */
    double nu[3], nnu[3];
    cross3(n, ref_u, nu); cross3(n, nu, nnu);
    for(int i = 0; i < 3; i += 1) up[i] = ref_u[i] + nu[i] + nnu[i] / (1.0 + c);
}

int drt_scene_apply_compat(drt_scene_input *s)
{
    drt_camera_input *cam = &s->camera;
    if(cam->has_legacy_axes && !cam->has_target)
    {
        double f[3] = { cam->forward[0], cam->forward[1], cam->forward[2] };
        double len = sqrt(dot3(f, f));
        if(len == 0.0) return drt_fail(DRT_E_SCENE, "legacy camera has a zero 'forward'");
        for(int i = 0; i < 3; i += 1) { f[i] /= len; cam->target[i] = cam->position[i] + cam->forward[i]; }
        cam->has_target = 1;
        /* roll (degrees about forward) that carries the derived up onto the file's up */
        if(dot3(cam->up, cam->up) > 0.0)
        {
            double up0[3], cr[3];
            derived_up_at_roll0(f, up0);
            cross3(up0, cam->up, cr);
            double deg = atan2(dot3(cr, f), dot3(up0, cam->up)) * (180.0 / 3.14159265358979323846);
            /* shipped files need exactly 0 or 180; keep those free of rounding noise */
            if(fabs(deg) < 1e-9) deg = 0.0;
            if(fabs(fabs(deg) - 180.0) < 1e-9) deg = 180.0;
            cam->roll = deg;
        }
        s->used_legacy = 1;
    }
    /* second-generation legacy files (example_scene.scn) give the camera no optics at all */
    if(cam->fov == 0.0 && cam->fdepth == 0.0 && cam->flength == 0.0)
    {
        cam->fov = 90.0; cam->fdepth = 8.0; cam->flength = 0.5; cam->aperture = 0.0;
        s->used_legacy = 1;
    }

    int has_base = 0, has_escape = 0;
    for(int i = 0; i < s->num_materials; i += 1) { has_base |= s->materials[i].is_base; has_escape |= s->materials[i].is_escape; }
    int inject = (has_base ? 0 : 1) + (has_escape ? 0 : 1);
    if(inject)
    {
        if(s->num_materials + inject > 16) return drt_fail(DRT_E_LIMIT, "no room to add the vacuum/escape materials a legacy scene needs");
        memmove(&s->materials[inject], &s->materials[0], sizeof(drt_material_input) * (size_t)s->num_materials);
        memset(&s->materials[0], 0, sizeof(drt_material_input) * (size_t)inject);
        int at = 0;
        if(!has_base)
        {
            drt_material_input *m = &s->materials[at++];
            strcpy(m->name, "vacuum");
            m->spd[DRT_SPD_REFRACT].method = DRT_SPD_METHOD_CONST;
            m->spd[DRT_SPD_REFRACT].value = 1.0;
            m->is_base = 1; m->dir_func = DRT_DIR_NONE;
        }
        if(!has_escape)
        {
            drt_material_input *m = &s->materials[at++];
            strcpy(m->name, "escape");
            m->is_escape = 1; m->dir_func = DRT_DIR_NONE;
        }
        s->num_materials += inject;
        s->used_legacy = 1;
    }
    for(int i = 0; i < s->num_materials; i += 1)
    {
        drt_material_input *m = &s->materials[i];
        if(m->has_lobes_key || m->is_black_body || m->is_base || m->is_escape) continue;
        m->num_lobes = 2;
        m->lobes[0] = DRT_LOBE_BP_DIFFUSE;
        m->lobes[1] = DRT_LOBE_BP_GLOSSY;
        m->dir_func = DRT_DIR_COS_WEIGHTED_HEMISPHERE;
        m->has_lobes_key = 1;
        s->used_legacy = 1;
    }
    return DRT_OK;
}

/* ---- writer: current grammar, numbers printed with %.17g so that atof() returns the same f64 ---- */

typedef struct { char *buf; size_t cap, len; int overflow; } sink;

static void put(sink *o, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    int n = vsnprintf(o->buf + o->len, o->len < o->cap ? o->cap - o->len : 0, fmt, ap);
    va_end(ap);
    if(n < 0 || o->len + (size_t)n >= o->cap) { o->overflow = 1; return; }
    o->len += (size_t)n;
}

/* The tokeniser has no exponent or '+' (read_scene.c:81-84): print plain decimals only. */
static void put_num(sink *o, double v)
{
    char tmp[64];
    snprintf(tmp, sizeof(tmp), "%.17g", v);
    if(strchr(tmp, 'e') || strchr(tmp, 'E')) snprintf(tmp, sizeof(tmp), "%.25f", v);
    if(!strchr(tmp, '.')) strcat(tmp, ".0");
    put(o, "%s", tmp);
}

static void put_vec(sink *o, const char *key, const double *v)
{
    put(o, "%s ", key); put_num(o, v[0]); put(o, ", "); put_num(o, v[1]); put(o, ", "); put_num(o, v[2]); put(o, "\n");
}

static void put_spd(sink *o, const char *key, const drt_spd_input *s)
{
    if(s->method == DRT_SPD_METHOD_NONE) return;
    put(o, "%s ", key);
    switch(s->method)
    {
        case DRT_SPD_METHOD_RGB:       put(o, "rgb "); put_num(o, s->rgb[0]); put(o, ", "); put_num(o, s->rgb[1]); put(o, ", "); put_num(o, s->rgb[2]); break;
        case DRT_SPD_METHOD_CSV:       put(o, "csv %s", s->csv); break;
        case DRT_SPD_METHOD_BLACKBODY: put(o, "blackbody "); put_num(o, s->value); break;
        case DRT_SPD_METHOD_CONST:     put(o, "constant "); put_num(o, s->value); break;
    }
    if(s->has_scale) { put(o, " scale "); put_num(o, s->scale); }
    put(o, "\n");
}

int drt_scene_write(const drt_scene_input *s, char *buf, size_t cap, size_t *written)
{
    sink o = { buf, cap, 0, 0 };
    const drt_camera_input *cam = &s->camera;
    put(&o, "Camera\n");
    put_vec(&o, "position", cam->position);
    put_vec(&o, "target", cam->target);
    put(&o, "roll "); put_num(&o, cam->roll); put(&o, "\n");
    put(&o, "fov "); put_num(&o, cam->fov); put(&o, "\n");
    put(&o, "fdepth "); put_num(&o, cam->fdepth); put(&o, "\n");
    put(&o, "flength "); put_num(&o, cam->flength); put(&o, "\n");
    put(&o, "aperture "); put_num(&o, cam->aperture); put(&o, "\n\n");
    static const char *spd_keys[DRT_SPD_COUNT] = { "emission", "diffuse", "glossy", "mirror", "refract", "extinct" };
    for(int i = 0; i < s->num_materials; i += 1)
    {
        const drt_material_input *m = &s->materials[i];
        put(&o, "Material\nname %s\n", m->name);
        for(int k = 0; k < DRT_SPD_COUNT; k += 1) put_spd(&o, spd_keys[k], &m->spd[k]);
        if(m->is_black_body) put(&o, "is_black_body true\n");
        if(m->shininess != 0.0) { put(&o, "shininess "); put_num(&o, m->shininess); put(&o, "\n"); }
        if(m->roughness != 0.0) { put(&o, "roughness "); put_num(&o, m->roughness); put(&o, "\n"); }
        if(m->num_lobes > 0 || m->dir_func != DRT_DIR_NONE)
        {
            put(&o, "bdsfs");
            for(int k = 0; k < m->num_lobes; k += 1) put(&o, "%s %s", k ? "," : "", drt_lobe_name(m->lobes[k]));
            put(&o, "\ndir_func %s\n", drt_dir_name(m->dir_func));
        }
        if(m->is_base) put(&o, "base_material\n");
        if(m->is_escape) put(&o, "escape_material\n");
        put(&o, "\n");
    }
    static const char *type_names[4] = { "none", "point", "sphere", "plane" };
    for(int i = 0; i < s->num_surfaces; i += 1)
    {
        const drt_surface_input *f = &s->surfaces[i];
        put(&o, "Surface\nname %s\ntype %s\n", f->name, type_names[f->type & 3]);
        put_vec(&o, "position", f->position);
        if(f->type == DRT_GEO_SPHERE) { put(&o, "radius "); put_num(&o, f->radius); put(&o, "\n"); }
        if(f->type == DRT_GEO_PLANE)  { put_vec(&o, "pointu", f->pointu); put_vec(&o, "pointv", f->pointv); }
        put(&o, "material %s\n\n", f->material_name);
    }
    if(o.overflow) return drt_fail(DRT_E_LIMIT, "scene text does not fit in %zu bytes", cap);
    if(written) *written = o.len;
    return DRT_OK;
}
