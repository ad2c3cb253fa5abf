/*
 * include/drt_host.h -- host-side C front-end of the render path (no CUDA here).
 *
 * Replaces, for Linux, the reference's scene/config front-end and file formats:
 *   parse_config               read_scene.c:604-765   -> drt_parse_config
 *   parse_scene                read_scene.c:569-602   -> drt_parse_scene
 *   load_csv_file_to_spectrum  read_scene.c:801-872   -> drt_load_csv_spectrum
 *   init_spd_tables            spectrum.c:1-47        -> drt_load_tables
 *   init_camera / init_scene   daily_ray_trace.c:49-77, 125-211 -> drt_build_scene
 *   rgb_f64_to_spectrum        spectrum.c:84-119      -> drt_rgb_to_spectrum
 *   generate_blackbody_spectrum spectrum.c:245-273    -> drt_blackbody_spectrum
 *   spectrum_to_xyz/rgb        spectrum.c:49-82       -> drt_spectrum_to_rgb
 *   .spd writer / reader       daily_ray_trace.c:667-680,758-770 / :1-28 -> drt_write_spd / drt_spd_to_rgb
 *   32-bpp BMP writer          win32_platform.c:11-41,136-161 -> drt_write_bmp
 * Every function returns 0 on success and a negative DRT_E_* code otherwise (the reference
 * printf()s and exit(-1)s, read_scene.c:196-203); drt_host_last_error() holds the message.
 */
#ifndef DRT_HOST_H
#define DRT_HOST_H

#include <stddef.h>
#include <stdint.h>
#include "drt_scene.h"

#ifdef __cplusplus
extern "C" {
#endif

enum
{
    DRT_OK = 0,
    DRT_E_IO = -1,        /* file could not be opened / read / written */
    DRT_E_PARSE = -2,     /* token the grammar does not allow here (reference: parse_error) */
    DRT_E_LIMIT = -3,     /* more than 16 materials / surfaces / lobes, >128 wavelengths or CSV rows */
    DRT_E_SCENE = -4,     /* scene cannot be rendered: no base/escape material, unknown lobe name, ... */
    DRT_E_ARG = -5
};

/* config_arguments, daily_ray_trace.h:28-55 (paths keep the reference's 64-byte fields) */
typedef struct
{
    uint32_t num_pixel_samples, max_cast_depth, output_width, output_height;
    double   min_wl, max_wl, wl_interval;
    char     input_scene[64];
    char     output_spd[64], average_spd[64], variance_spd[64];
    char     output_bmp[64], average_bmp[64], variance_bmp[64];
    char     white_spd[64], cmf_x[64], cmf_y[64], cmf_z[64];
    char     red_spd[64], green_spd[64], blue_spd[64], cyan_spd[64], magenta_spd[64], yellow_spd[64];
    int32_t  pixel_scheme;   /* DRT_PIXEL_* */
} drt_config;

/* spd_input_data, read_scene.h:14-38 */
enum { DRT_SPD_METHOD_NONE = 0, DRT_SPD_METHOD_RGB = 1, DRT_SPD_METHOD_CSV = 2, DRT_SPD_METHOD_BLACKBODY = 3, DRT_SPD_METHOD_CONST = 4 };
typedef struct
{
    int32_t method;
    int32_t has_scale;
    double  scale;
    double  rgb[3];
    char    csv[64];
    double  value;          /* blackbody temperature or constant */
} drt_spd_input;

/* material_input_data / surface_input_data / camera_input_data, read_scene.h:1-87 */
typedef struct
{
    char    name[32];
    int32_t is_base, is_escape, is_black_body, is_emissive;
    double  shininess, roughness;
    drt_spd_input spd[DRT_SPD_COUNT];
    int32_t num_lobes;
    int32_t lobes[DRT_MAX_LOBES];
    int32_t dir_func;
    int32_t has_lobes_key;  /* a bdsfs/dir_func key was present (legacy files have none) */
} drt_material_input;

typedef struct
{
    char    name[32];
    int32_t type;
    double  position[3];
    double  radius;
    double  pointu[3], pointv[3];
    char    material_name[32];
} drt_surface_input;

typedef struct
{
    double  position[3], target[3];
    double  roll, fov, fdepth, flength, aperture;
    /* legacy grammar (keywords.h:6-8 exist, parse_camera has no case for them) */
    int32_t has_target, has_legacy_axes;
    double  up[3], right[3], forward[3];
} drt_camera_input;

typedef struct
{
    drt_camera_input   camera;
    int32_t            num_materials, num_surfaces;
    drt_material_input materials[16];
    drt_surface_input  surfaces[16];
    int32_t            used_legacy;   /* any legacy-compat rule fired while parsing / fixing up */
} drt_scene_input;

/* flags for drt_parse_scene */
enum
{
    DRT_PARSE_STRICT = 0,        /* exactly the reference grammar: legacy files are DRT_E_PARSE */
    DRT_PARSE_LEGACY_COMPAT = 1  /* up/right/forward, alias keys, default lobes, injected vacuum/escape */
};

const char *drt_host_last_error(void);

int drt_parse_config(const char *text, size_t size, drt_config *out);
int drt_parse_config_file(const char *path, drt_config *out);

int drt_parse_scene(const char *text, size_t size, int flags, drt_scene_input *out);
/* Applies the legacy-compat fix-ups in place (no-op on a current-grammar scene). */
int drt_scene_apply_compat(drt_scene_input *scene);
/* Writes the scene back in the CURRENT grammar, parseable by the unmodified reference. */
int drt_scene_write(const drt_scene_input *scene, char *buf, size_t cap, size_t *written);

/* root_dir is prepended to relative paths; '\\' in a path is read as '/'. */
int drt_load_csv_spectrum(const char *root_dir, const char *path, int n, double min_wl, double interval, double *dst);
int drt_load_tables(const drt_config *cfg, const char *root_dir, drt_tables *out);

void drt_rgb_to_spectrum(const drt_tables *t, const double rgb[3], double *dst);
void drt_blackbody_spectrum(const drt_tables *t, double temperature, double *dst);
void drt_spectrum_to_xyz(const drt_tables *t, const double *spd, double xyz[3]);
void drt_spectrum_to_rgb(const drt_tables *t, const double *spd, double rgb[3]);
uint32_t drt_rgb_to_bgra8(const double rgb[3]);

int drt_build_scene(const drt_scene_input *in, const drt_tables *tables, const char *root_dir,
                    uint32_t width_px, uint32_t height_px, drt_scene *scene, drt_camera *camera);
/* parse + compat + build in one call: the equivalent of load_scene, daily_ray_trace.c:30-47 */
int drt_load_scene_file(const char *root_dir, const char *scene_path, const drt_tables *tables, int flags,
                        uint32_t width_px, uint32_t height_px, drt_scene *scene, drt_camera *camera);

/* spd_file_header, daily_ray_trace.h:59-68 (40 bytes on disk) */
typedef struct
{
    uint32_t id, width, height, num_wavelengths, has_filter, pad_;
    double   min_wl, wl_interval;
} drt_spd_header;
#define DRT_SPD_FILE_ID 0xedfeefbeu

/* Film planes as the device produces them: f32, pixel-major, N values per pixel. */
int drt_write_spd_sum(const char *path, const drt_tables *t, uint32_t w, uint32_t h, const float *sum, const float *filter);
int drt_write_spd_plain(const char *path, const drt_tables *t, uint32_t w, uint32_t h, const float *values, int normalise_per_pixel);
/* .spd -> linear RGB (f64 triplets), as spd_file_to_rgb_f64_pixels does */
int drt_spd_to_rgb(const char *path, const drt_tables *t, uint32_t *w, uint32_t *h, double **rgb_out);
int drt_write_bmp(const char *path, uint32_t w, uint32_t h, const uint32_t *bgra);
int drt_write_bmp_rgb(const char *path, uint32_t w, uint32_t h, const double *rgb);

#ifdef __cplusplus
}
#endif
#endif
