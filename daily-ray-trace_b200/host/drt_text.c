/*
 * host/drt_text.c -- tokeniser and the two text front-ends (config.cfg and .scn).
 *
 * Same language as the reference accepts, restated as a
 * table-driven reader instead of the reference's nested switches:
 *   tokeniser      read_scene.c:62-155   FLOAT = run of [0-9.-] valued by atof; WORD starts with a
 *                                         letter and continues with [A-Za-z0-9_.\]; anything else
 *                                         (spaces, commas, '#', ...) separates tokens.
 *   config keys    read_scene.c:604-765 / keywords.h:39-69
 *   scene blocks   read_scene.c:345-602
 * Differences, all additive: '/' is a word character (Linux paths); input need not be
 * NUL-terminated; fixed-size fields are bounds-checked; errors are returned, not exit(-1);
 * DRT_PARSE_LEGACY_COMPAT accepts the older scene grammar (drt_compat.c).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdarg.h>
#include "drt_host.h"
#include "drt_host_internal.h"

static char g_error[512];

const char *drt_host_last_error(void) { return g_error; }

int drt_fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
    return code;
}

/* ---------------------------------------------------------------- tokeniser */

typedef enum { TK_END, TK_FLOAT, TK_WORD } tk_kind;

typedef struct
{
    tk_kind     kind;
    const char *text;
    size_t      len;
    double      value;
} token;

typedef struct
{
    const char *at, *end;
    int         line;
} cursor;

static int is_digit(char c)  { return c >= '0' && c <= '9'; }
static int is_alpha(char c)  { return (c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z'); }
static int is_numch(char c)  { return is_digit(c) || c == '.' || c == '-'; }
static int is_wordch(char c) { return is_alpha(c) || is_digit(c) || c == '_' || c == '.' || c == '\\' || c == '/'; }

/* atof on a bounded, not necessarily NUL-terminated run (the reference calls atof(src), read_scene.c:129,
 * which may read an exponent beyond the [0-9.-] run; the same bytes are handed to atof here). */
static double bounded_atof(const char *s, const char *end)
{
    char tmp[64];
    size_t n = 0;
    while(s + n < end && n < sizeof(tmp) - 1 && (is_numch(s[n]) || s[n] == 'e' || s[n] == 'E' || s[n] == '+')) n += 1;
    memcpy(tmp, s, n);
    tmp[n] = 0;
    return atof(tmp);
}

static token next_token(cursor *c)
{
    token t; memset(&t, 0, sizeof(t));
    for(;;)
    {
        if(c->at >= c->end) { t.kind = TK_END; t.text = c->end; return t; }
        char ch = *c->at;
        if(is_numch(ch))
        {
            t.kind = TK_FLOAT; t.text = c->at; t.value = bounded_atof(c->at, c->end);
            while(c->at < c->end && is_numch(*c->at)) c->at += 1;
            t.len = (size_t)(c->at - t.text);
            return t;
        }
        if(is_alpha(ch))
        {
            t.kind = TK_WORD; t.text = c->at;
            while(c->at < c->end && is_wordch(*c->at)) c->at += 1;
            t.len = (size_t)(c->at - t.text);
            return t;
        }
        if(ch == '\n') c->line += 1;
        c->at += 1;   /* separator */
    }
}

static token peek_token(const cursor *c) { cursor k = *c; return next_token(&k); }

static int word_is(const token *t, const char *s)
{
    return t->kind == TK_WORD && strlen(s) == t->len && memcmp(t->text, s, t->len) == 0;
}

static int bad_token(const cursor *c, const token *t, const char *expect)
{
    if(t->kind == TK_END) return drt_fail(DRT_E_PARSE, "line %d: unexpected end of input, expected %s", c->line + 1, expect);
    return drt_fail(DRT_E_PARSE, "line %d: unexpected '%.*s', expected %s", c->line + 1, (int)t->len, t->text, expect);
}

static int read_float(cursor *c, double *dst)
{
    token t = next_token(c);
    if(t.kind != TK_FLOAT) return bad_token(c, &t, "a number");
    *dst = t.value;
    return DRT_OK;
}

static int read_vec3(cursor *c, double *dst)
{
    int rc;
    for(int i = 0; i < 3; i += 1) if((rc = read_float(c, &dst[i])) != DRT_OK) return rc;
    return DRT_OK;
}

/* parse_uint, read_scene.c:213-218: a float truncated to u32 */
static int read_uint(cursor *c, uint32_t *dst)
{
    double f = 0.0; int rc = read_float(c, &f);
    if(rc == DRT_OK) *dst = (uint32_t)f;
    return rc;
}

static int read_word(cursor *c, char *dst, size_t cap)
{
    token t = next_token(c);
    if(t.kind == TK_END) return bad_token(c, &t, "a word");
    if(t.len >= cap) return drt_fail(DRT_E_LIMIT, "line %d: '%.*s' is longer than %zu characters", c->line + 1, (int)t.len, t.text, cap - 1);
    memset(dst, 0, cap);
    memcpy(dst, t.text, t.len);
    return DRT_OK;
}

static int read_bool(cursor *c, int32_t *dst)
{
    token t = next_token(c);
    if(word_is(&t, "true"))  { *dst = 1; return DRT_OK; }
    if(word_is(&t, "false")) { *dst = 0; return DRT_OK; }
    return bad_token(c, &t, "true or false");
}

/* ---------------------------------------------------------------- config.cfg */

typedef enum { F_UINT, F_FLOAT, F_PATH, F_SCHEME } field_kind;
typedef struct { const char *key; field_kind kind; size_t offset; } config_field;

#define CFG(name, kind) { #name, kind, offsetof(drt_config, name) }
static const config_field config_fields[] =
{
    CFG(num_pixel_samples, F_UINT), CFG(max_cast_depth, F_UINT), CFG(output_width, F_UINT), CFG(output_height, F_UINT),
    CFG(min_wl, F_FLOAT), CFG(max_wl, F_FLOAT), CFG(wl_interval, F_FLOAT),
    CFG(input_scene, F_PATH),
    CFG(output_spd, F_PATH), CFG(average_spd, F_PATH), CFG(variance_spd, F_PATH),
    CFG(output_bmp, F_PATH), CFG(average_bmp, F_PATH), CFG(variance_bmp, F_PATH),
    CFG(white_spd, F_PATH), CFG(cmf_x, F_PATH), CFG(cmf_y, F_PATH), CFG(cmf_z, F_PATH),
    CFG(red_spd, F_PATH), CFG(green_spd, F_PATH), CFG(blue_spd, F_PATH),
    CFG(cyan_spd, F_PATH), CFG(magenta_spd, F_PATH), CFG(yellow_spd, F_PATH),
    CFG(pixel_scheme, F_SCHEME),
};
#undef CFG

int drt_parse_config(const char *text, size_t size, drt_config *out)
{
    cursor c = { text, text + size, 0 };
    memset(out, 0, sizeof(*out));
    for(;;)
    {
        token key = next_token(&c);
        if(key.kind == TK_END) return DRT_OK;
        const config_field *f = NULL;
        for(size_t i = 0; i < sizeof(config_fields) / sizeof(config_fields[0]); i += 1)
            if(word_is(&key, config_fields[i].key)) { f = &config_fields[i]; break; }
        if(!f) return bad_token(&c, &key, "a config key");
        char *dst = (char *)out + f->offset;
        int rc = DRT_OK;
        switch(f->kind)
        {
            case F_UINT:  rc = read_uint(&c, (uint32_t *)dst); break;
            case F_FLOAT: rc = read_float(&c, (double *)dst);  break;
            case F_PATH:  rc = read_word(&c, dst, 64);         break;
            case F_SCHEME:
            {
                token v = next_token(&c);
                if(word_is(&v, "pixel_random"))      *(int32_t *)dst = DRT_PIXEL_RANDOM;
                else if(word_is(&v, "pixel_center")) *(int32_t *)dst = DRT_PIXEL_CENTER;
                else rc = bad_token(&c, &v, "pixel_random or pixel_center");
                break;
            }
        }
        if(rc != DRT_OK) return rc;
    }
}

int drt_parse_config_file(const char *path, drt_config *out)
{
    size_t size = 0;
    char *text = drt_read_text_file(NULL, path, &size);
    if(!text) return drt_fail(DRT_E_IO, "cannot read config '%s'", path);
    int rc = drt_parse_config(text, size, out);
    free(text);
    return rc;
}

/* ---------------------------------------------------------------- .scn */

static const char *lobe_names[DRT_LOBE_COUNT] =
{
    "bp_diffuse_bdsf", "bp_glossy_bdsf", "mirror_bdsf", "fs_conductor_bdsf",
    "fs_dielectric_reflectance_bdsf", "fs_dielectric_transmittance_bdsf", "ct_conductor_bdsf",
};
static const char *dir_names[DRT_DIR_COUNT] =
{
    "cos_weighted_sample_hemisphere", "uniform_sample_hemisphere", "sample_specular_direction",
    "sample_transmit_direction", "sample_reflect_or_transmit_direction", "sample_ct_direction",
};

const char *drt_lobe_name(int id) { return (id >= 0 && id < DRT_LOBE_COUNT) ? lobe_names[id] : "?"; }
const char *drt_dir_name(int id)  { return (id >= 0 && id < DRT_DIR_COUNT) ? dir_names[id] : "?"; }

static int is_block_start(const token *t)
{
    return t->kind == TK_END || word_is(t, "Camera") || word_is(t, "Material") || word_is(t, "Surface");
}

/* parse_spd_method, read_scene.c:263-305 */
static int read_spd(cursor *c, drt_spd_input *dst)
{
    token m = next_token(c);
    int rc;
    if(word_is(&m, "rgb"))            { dst->method = DRT_SPD_METHOD_RGB;       rc = read_vec3(c, dst->rgb); }
    else if(word_is(&m, "csv"))       { dst->method = DRT_SPD_METHOD_CSV;       rc = read_word(c, dst->csv, sizeof(dst->csv)); }
    else if(word_is(&m, "blackbody")) { dst->method = DRT_SPD_METHOD_BLACKBODY; rc = read_float(c, &dst->value); }
    else if(word_is(&m, "constant"))  { dst->method = DRT_SPD_METHOD_CONST;     rc = read_float(c, &dst->value); }
    else return bad_token(c, &m, "rgb, csv, blackbody or constant");
    if(rc != DRT_OK) return rc;
    token l = peek_token(c);
    if(word_is(&l, "scale"))
    {
        next_token(c);
        dst->has_scale = 1;
        rc = read_float(c, &dst->scale);
    }
    return rc;
}

static int read_camera(cursor *c, drt_scene_input *scene, int flags)
{
    drt_camera_input *cam = &scene->camera;
    for(;;)
    {
        token l = peek_token(c);
        if(is_block_start(&l)) return DRT_OK;
        token k = next_token(c);
        int rc;
        if(word_is(&k, "position"))      rc = read_vec3(c, cam->position);
        else if(word_is(&k, "target"))   { rc = read_vec3(c, cam->target); cam->has_target = 1; }
        else if(word_is(&k, "roll"))     rc = read_float(c, &cam->roll);
        else if(word_is(&k, "fov"))      rc = read_float(c, &cam->fov);
        else if(word_is(&k, "fdepth"))   rc = read_float(c, &cam->fdepth);
        else if(word_is(&k, "flength"))  rc = read_float(c, &cam->flength);
        else if(word_is(&k, "aperture")) rc = read_float(c, &cam->aperture);
        else if((flags & DRT_PARSE_LEGACY_COMPAT) && (word_is(&k, "up") || word_is(&k, "right") || word_is(&k, "forward")))
        {
            double *dst = word_is(&k, "up") ? cam->up : word_is(&k, "right") ? cam->right : cam->forward;
            rc = read_vec3(c, dst);
            cam->has_legacy_axes = 1;
            scene->used_legacy = 1;
        }
        else return bad_token(c, &k, "a Camera key");
        if(rc != DRT_OK) return rc;
    }
}

/* parse_bdsfs, read_scene.c:308-328: names up to (not including) the dir_func key */
static int read_lobes(cursor *c, drt_material_input *m)
{
    m->num_lobes = 0;
    m->has_lobes_key = 1;
    for(;;)
    {
        token l = peek_token(c);
        if(word_is(&l, "dir_func")) return DRT_OK;
        if(l.kind != TK_WORD) return bad_token(c, &l, "a bdsf name or dir_func");
        if(m->num_lobes == DRT_MAX_LOBES) return drt_fail(DRT_E_LIMIT, "line %d: more than %d bdsfs", c->line + 1, DRT_MAX_LOBES);
        token w = next_token(c);
        int id = -1;
        for(int i = 0; i < DRT_LOBE_COUNT; i += 1) if(word_is(&w, lobe_names[i])) id = i;
        if(id < 0) return drt_fail(DRT_E_SCENE, "line %d: '%.*s' is not in bdsf_list.h", c->line + 1, (int)w.len, w.text);
        m->lobes[m->num_lobes++] = id;
    }
}

static int read_material(cursor *c, drt_scene_input *scene, int flags)
{
    if(scene->num_materials == 16) return drt_fail(DRT_E_LIMIT, "line %d: more than 16 materials", c->line + 1);
    drt_material_input *m = &scene->materials[scene->num_materials++];
    m->dir_func = DRT_DIR_NONE;
    for(;;)
    {
        token l = peek_token(c);
        if(is_block_start(&l)) return DRT_OK;
        token k = next_token(c);
        int rc = DRT_OK;
        if(word_is(&k, "name"))               rc = read_word(c, m->name, sizeof(m->name));
        else if(word_is(&k, "diffuse"))       rc = read_spd(c, &m->spd[DRT_SPD_DIFFUSE]);
        else if(word_is(&k, "glossy"))        rc = read_spd(c, &m->spd[DRT_SPD_GLOSSY]);
        else if(word_is(&k, "emission"))      { rc = read_spd(c, &m->spd[DRT_SPD_EMISSION]); m->is_emissive = 1; }
        else if(word_is(&k, "mirror"))        rc = read_spd(c, &m->spd[DRT_SPD_MIRROR]);
        else if(word_is(&k, "refract"))       rc = read_spd(c, &m->spd[DRT_SPD_REFRACT]);
        else if(word_is(&k, "extinct"))       rc = read_spd(c, &m->spd[DRT_SPD_EXTINCT]);
        else if(word_is(&k, "is_black_body")) rc = read_bool(c, &m->is_black_body);
        else if(word_is(&k, "shininess"))     rc = read_float(c, &m->shininess);
        else if(word_is(&k, "roughness"))     rc = read_float(c, &m->roughness);
        else if(word_is(&k, "bdsfs"))         rc = read_lobes(c, m);
        else if(word_is(&k, "dir_func"))
        {
            token w = next_token(c);
            m->has_lobes_key = 1;
            m->dir_func = DRT_DIR_NONE;
            for(int i = 0; i < DRT_DIR_COUNT; i += 1) if(word_is(&w, dir_names[i])) m->dir_func = i;
            if(m->dir_func == DRT_DIR_NONE) rc = drt_fail(DRT_E_SCENE, "line %d: '%.*s' is not a dir_func of bdsf_list.h", c->line + 1, (int)w.len, w.text);
        }
        else if(word_is(&k, "base_material"))   m->is_base = 1;
        else if(word_is(&k, "escape_material")) m->is_escape = 1;
        else if((flags & DRT_PARSE_LEGACY_COMPAT) && word_is(&k, "is_blackbody")) { rc = read_bool(c, &m->is_black_body); scene->used_legacy = 1; }
        else return bad_token(c, &k, "a Material key");
        if(rc != DRT_OK) return rc;
    }
}

static int read_surface(cursor *c, drt_scene_input *scene, int flags)
{
    if(scene->num_surfaces == 16) return drt_fail(DRT_E_LIMIT, "line %d: more than 16 surfaces", c->line + 1);
    drt_surface_input *s = &scene->surfaces[scene->num_surfaces++];
    int compat = flags & DRT_PARSE_LEGACY_COMPAT;
    for(;;)
    {
        token l = peek_token(c);
        if(is_block_start(&l)) return DRT_OK;
        token k = next_token(c);
        int rc = DRT_OK;
        if(word_is(&k, "name")) rc = read_word(c, s->name, sizeof(s->name));
        else if(word_is(&k, "type"))
        {
            token v = next_token(c);
            if(word_is(&v, "point"))       s->type = DRT_GEO_POINT;
            else if(word_is(&v, "sphere")) s->type = DRT_GEO_SPHERE;
            else if(word_is(&v, "plane"))  s->type = DRT_GEO_PLANE;
            else rc = bad_token(c, &v, "point, sphere or plane");
        }
        else if(word_is(&k, "position")) rc = read_vec3(c, s->position);
        else if(word_is(&k, "radius"))   rc = read_float(c, &s->radius);
        else if(word_is(&k, "pointu"))   rc = read_vec3(c, s->pointu);
        else if(word_is(&k, "pointv"))   rc = read_vec3(c, s->pointv);
        else if(word_is(&k, "material")) rc = read_word(c, s->material_name, sizeof(s->material_name));
        else if(compat && (word_is(&k, "center") || word_is(&k, "origin"))) { rc = read_vec3(c, s->position); scene->used_legacy = 1; }
        else if(compat && word_is(&k, "point_u")) { rc = read_vec3(c, s->pointu); scene->used_legacy = 1; }
        else if(compat && word_is(&k, "point_v")) { rc = read_vec3(c, s->pointv); scene->used_legacy = 1; }
        else return bad_token(c, &k, "a Surface key");
        if(rc != DRT_OK) return rc;
    }
}

int drt_parse_scene(const char *text, size_t size, int flags, drt_scene_input *out)
{
    cursor c = { text, text + size, 0 };
    memset(out, 0, sizeof(*out));
    for(;;)
    {
        token t = next_token(&c);
        int rc;
        if(t.kind == TK_END) break;
        if(word_is(&t, "Camera"))        rc = read_camera(&c, out, flags);
        else if(word_is(&t, "Material")) rc = read_material(&c, out, flags);
        else if(word_is(&t, "Surface"))  rc = read_surface(&c, out, flags);
        else return bad_token(&c, &t, "Camera, Material or Surface");
        if(rc != DRT_OK) return rc;
    }
    if(flags & DRT_PARSE_LEGACY_COMPAT) return drt_scene_apply_compat(out);
    return DRT_OK;
}
