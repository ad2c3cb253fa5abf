/*
 * include/drt_scene.h -- flattened, pointer-free scene description shared by the host C
 * front-end (daily-ray-trace_b200/host), the CUDA path (csrc) and the CPU oracle (oracle/).
 *
 * It restates the reference's AoS data model as plain-old-data with enum ids in place of
 * function pointers and SPD pointers:
 *   object_geometry   daily_ray_trace.h:78-93    -> drt_surface
 *   object_material   daily_ray_trace.h:95-111   -> drt_material  (bdsfs[]/sample_direction -> ids)
 *   scene_data        daily_ray_trace.h:146-156  -> drt_scene
 *   camera_data       daily_ray_trace.h:158-170  -> drt_camera
 *   cmfs / rgb_spds   spectrum.h:17-33           -> drt_tables
 * Lobe and sampler ids are the positions in src/bdsf_list.h (lobes lines 1-7, samplers 9-14).
 * All reals are f64 exactly as the reference computes them on the host; the device narrows.
 */
#ifndef DRT_SCENE_H
#define DRT_SCENE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DRT_MAX_WAVELENGTHS 128   /* MAX_NUM_SPECTRUM_VALUES, spectrum.h:1 */
#define DRT_MAX_SURFACES    16    /* read_scene.h:85-86 */
#define DRT_MAX_MATERIALS   17    /* 16 parsed + the phantom entry of init_scene (daily_ray_trace.c:128) */
#define DRT_MAX_LOBES       16    /* daily_ray_trace.h:109 */
#define DRT_TRANS_WL        630.0 /* daily_ray_trace.c:381 */
#define DRT_RAY_FUDGE       0.0001 /* vis_fudge, daily_ray_trace.c:237 */

/* geometry_type, daily_ray_trace.h:7-14 */
enum { DRT_GEO_NONE = 0, DRT_GEO_POINT = 1, DRT_GEO_SPHERE = 2, DRT_GEO_PLANE = 3 };

/* film_sample_scheme, daily_ray_trace.h:20-26 */
enum { DRT_PIXEL_NONE = 0, DRT_PIXEL_CENTER = 1, DRT_PIXEL_RANDOM = 2 };

/* src/bdsf_list.h:1-7 */
enum
{
    DRT_LOBE_BP_DIFFUSE = 0,
    DRT_LOBE_BP_GLOSSY = 1,
    DRT_LOBE_MIRROR = 2,
    DRT_LOBE_FS_CONDUCTOR = 3,
    DRT_LOBE_FS_DIELECTRIC_REFLECTANCE = 4,
    DRT_LOBE_FS_DIELECTRIC_TRANSMITTANCE = 5,
    DRT_LOBE_CT_CONDUCTOR = 6,
    DRT_LOBE_COUNT = 7
};

/* src/bdsf_list.h:9-14 */
enum
{
    DRT_DIR_COS_WEIGHTED_HEMISPHERE = 0,
    DRT_DIR_UNIFORM_HEMISPHERE = 1,
    DRT_DIR_SPECULAR = 2,
    DRT_DIR_TRANSMIT = 3,
    DRT_DIR_REFLECT_OR_TRANSMIT = 4,
    DRT_DIR_CT = 5,
    DRT_DIR_COUNT = 6,
    DRT_DIR_NONE = -1
};

/* which of the six material SPDs were given in the scene file (init_spd leaves the rest NULL) */
enum { DRT_SPD_EMISSION = 0, DRT_SPD_DIFFUSE = 1, DRT_SPD_GLOSSY = 2, DRT_SPD_MIRROR = 3, DRT_SPD_REFRACT = 4, DRT_SPD_EXTINCT = 5, DRT_SPD_COUNT = 6 };

typedef struct
{
    int32_t type;          /* DRT_GEO_* */
    int32_t material;      /* index into drt_scene.materials */
    double  position[3];
    double  radius;        /* sphere */
    double  normal[3];     /* plane: normalise(u x v), geometry.c:203-209 */
    double  u[3];          /* plane bounds vector pointu - position */
    double  v[3];          /* plane bounds vector pointv - position */
    char    name[32];
} drt_surface;

typedef struct
{
    char    name[32];
    int32_t is_black_body;
    int32_t is_emissive;
    double  shininess;
    double  roughness;
    int32_t num_lobes;
    int32_t lobes[DRT_MAX_LOBES];   /* DRT_LOBE_*, -1 = name not found in bdsf_list.h */
    int32_t dir_func;               /* DRT_DIR_* */
    int32_t spd_mask;               /* bit k set: spd[k] was given */
    double  spd[DRT_SPD_COUNT][DRT_MAX_WAVELENGTHS];
} drt_material;

typedef struct
{
    int32_t      num_wavelengths;   /* N, spectrum.c:3 */
    double       min_wl, max_wl, wl_interval;
    int32_t      num_surfaces;
    int32_t      num_materials;     /* includes the phantom last entry */
    int32_t      base_material;     /* -1 when the scene names none */
    int32_t      escape_material;
    drt_surface  surfaces[DRT_MAX_SURFACES];
    drt_material materials[DRT_MAX_MATERIALS];
} drt_scene;

typedef struct
{
    double forward[3], right[3], up[3];
    double aperture_position[3];
    double aperture_radius, focal_depth, focal_length;
    double film_bottom_left[3];
    double pixel_width, pixel_height;
    /* thin lens: find_rotation_between_vectors((0,0,1), forward), column-major as geometry.h mat3x3 */
    double lens_rotation[9];
} drt_camera;

/* cmfs {rw,x,y,z} (spectrum.h:17-23) and rgb_spds {white,red,green,blue,cyan,magenta,yellow} (:25-33) */
typedef struct
{
    int32_t num_wavelengths;
    double  min_wl, wl_interval;
    double  ref_white[DRT_MAX_WAVELENGTHS];
    double  cmf_x[DRT_MAX_WAVELENGTHS];
    double  cmf_y[DRT_MAX_WAVELENGTHS];
    double  cmf_z[DRT_MAX_WAVELENGTHS];
    double  rgb_basis[7][DRT_MAX_WAVELENGTHS];
} drt_tables;

/* one render request: the arguments of render_image's triple loop (daily_ray_trace.c:710-718) */
typedef struct
{
    uint32_t width, height;
    uint32_t sample_begin, sample_end;  /* global sample indices [begin,end) rendered by this call */
    uint32_t max_depth;                 /* max_cast_depth */
    int32_t  pixel_scheme;              /* DRT_PIXEL_* */
    uint64_t seed;                      /* per-path stream seed, include/drt_rng.h */
} drt_render_params;

#ifdef __cplusplus
}
#endif
#endif
