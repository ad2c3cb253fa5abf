/* host/drt_host_internal.h -- helpers shared by the host C files (not part of the public ABI). */
#ifndef DRT_HOST_INTERNAL_H
#define DRT_HOST_INTERNAL_H
#include <stddef.h>

/* The reference's PI is a long double literal (types.h:1); expressions that mix it with f64 are
 * evaluated in x87 extended precision, so the host setup keeps the same type to stay bit-identical. */
#define DRT_PI_L 3.1415926535897932385L

int         drt_fail(int code, const char *fmt, ...);
/* Reads root/path ('\\' -> '/'); returns a malloc'd NUL-terminated buffer or NULL. */
char       *drt_read_text_file(const char *root_dir, const char *path, size_t *size);
void        drt_join_path(char *dst, size_t cap, const char *root_dir, const char *path);
const char *drt_lobe_name(int id);
const char *drt_dir_name(int id);

#endif
