/*
 * oracle/shim/Windows.h -- TEST INFRASTRUCTURE, not product code.
 *
 * A minimal Linux stand-in for <Windows.h> so that the UNMODIFIED reference
 * translation unit (/root/reference/src/win32_main.c, which #includes every
 * other file through daily_ray_trace.h:1,70-76,127-133,179) compiles with gcc.
 * Only the Win32 names the reference actually touches are provided:
 *   win32_platform.c:53-62   VirtualAlloc / VirtualFree  -> zero-filled calloc
 *   win32_platform.c:64-134  CreateFile/ReadFile/WriteFile/GetFileSize/...  -> stdio
 *   win32_platform.c:180-195 QueryPerformanceCounter    -> clock_gettime (1 GHz)
 *   win32_platform.c:11-41   BITMAPFILEHEADER/BITMAPINFOHEADER (2-byte packed)
 *   win32_platform.c:43-51   SYSTEM_INFO / GetSystemInfo
 * Path separators: the reference writes "spectra\\x.csv"; the shim maps '\\' to '/'.
 */
#ifndef DRT_ORACLE_WINDOWS_SHIM_H
#define DRT_ORACLE_WINDOWS_SHIM_H

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <time.h>

typedef FILE    *HANDLE;
typedef uint32_t DWORD;
typedef uint16_t WORD;
typedef int32_t  LONG;
typedef int      BOOL;
typedef void    *LPVOID;

typedef union { struct { DWORD LowPart; LONG HighPart; }; long long QuadPart; } LARGE_INTEGER;

#pragma pack(push, 2)
typedef struct { WORD bfType; DWORD bfSize; WORD bfReserved1; WORD bfReserved2; DWORD bfOffBits; } BITMAPFILEHEADER;
#pragma pack(pop)
typedef struct
{
    DWORD biSize; LONG biWidth; LONG biHeight; WORD biPlanes; WORD biBitCount; DWORD biCompression;
    DWORD biSizeImage; LONG biXPelsPerMeter; LONG biYPelsPerMeter; DWORD biClrUsed; DWORD biClrImportant;
} BITMAPINFOHEADER;
#define BI_RGB 0

typedef struct { DWORD dwPageSize; } SYSTEM_INFO;
static inline void GetSystemInfo(SYSTEM_INFO *s) { s->dwPageSize = 4096; }

#define MEM_COMMIT     0x1000
#define MEM_RESERVE    0x2000
#define MEM_RELEASE    0x8000
#define PAGE_READWRITE 0x04
/* VirtualAlloc returns zeroed pages; the reference depends on it
 * (names memcpy'd without NUL daily_ray_trace.c:146-147; films :689-691).
 * +4096 slack: load_csv_file_to_spectrum writes buffer[size] (read_scene.c:810). */
static inline void *VirtualAlloc(void *addr, size_t size, DWORD type, DWORD prot)
{
    (void)addr; (void)type; (void)prot;
    return calloc(size + 4096, 1);
}
static inline BOOL VirtualFree(void *p, size_t size, DWORD type) { (void)size; (void)type; free(p); return 1; }

#define GENERIC_READ          0x80000000u
#define GENERIC_WRITE         0x40000000u
#define FILE_SHARE_READ       1
#define FILE_SHARE_WRITE      2
#define CREATE_ALWAYS         2
#define OPEN_EXISTING         3
#define FILE_ATTRIBUTE_NORMAL 0x80
#define FILE_BEGIN            0

static inline HANDLE CreateFile(const char *path, DWORD access, DWORD share, void *sec, DWORD disp, DWORD attr, void *tmpl)
{
    (void)share; (void)sec; (void)attr; (void)tmpl;
    char fixed[512];
    size_t n = strlen(path);
    if(n >= sizeof(fixed)) n = sizeof(fixed) - 1;
    for(size_t i = 0; i < n; i += 1) fixed[i] = (path[i] == '\\') ? '/' : path[i];
    fixed[n] = 0;
    const char *mode = (disp == CREATE_ALWAYS) ? "w+b" : ((access & GENERIC_WRITE) ? "r+b" : "rb");
    FILE *f = fopen(fixed, mode);
    if(!f) { fprintf(stderr, "shim CreateFile: cannot open '%s'\n", fixed); exit(-2); }
    return f;
}
static inline BOOL  CloseHandle(HANDLE h) { return fclose(h) == 0; }
static inline DWORD GetFileSize(HANDLE h, DWORD *hi)
{
    (void)hi;
    long at = ftell(h); fseek(h, 0, SEEK_END);
    long sz = ftell(h); fseek(h, at, SEEK_SET);
    return (DWORD)sz;
}
static inline BOOL ReadFile(HANDLE h, void *dst, DWORD n, DWORD *got, void *ov)   { (void)ov; *got = (DWORD)fread(dst, 1, n, h);  return 1; }
static inline BOOL WriteFile(HANDLE h, const void *src, DWORD n, DWORD *put, void *ov) { (void)ov; *put = (DWORD)fwrite(src, 1, n, h); return 1; }
static inline DWORD SetFilePointer(HANDLE h, LONG loc, LONG *hi, DWORD how) { (void)hi; (void)how; fseek(h, loc, SEEK_SET); return (DWORD)loc; }

static inline BOOL QueryPerformanceFrequency(LARGE_INTEGER *f) { f->QuadPart = 1000000000LL; return 1; }
static inline BOOL QueryPerformanceCounter(LARGE_INTEGER *c)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    c->QuadPart = (long long)ts.tv_sec * 1000000000LL + ts.tv_nsec;
    return 1;
}

#endif
