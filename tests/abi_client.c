/* A plain-C99 client of include/csmri_dc.h: what a cgo / JNI / ctypes-free
 * consumer sees.  Built and run by tests/test_abi_and_host.py (no GPU needed):
 * the header must compile as C, the library must load with dlopen, and the
 * argument checks must answer before anything touches a device. */
#include <dlfcn.h>
#include <stdio.h>
#include <string.h>

#include "csmri_dc.h"

#define GET(name)                                              \
  do {                                                         \
    *(void**)(&p_##name) = dlsym(h, #name);                    \
    if (!p_##name) {                                           \
      fprintf(stderr, "missing symbol %s\n", #name);           \
      return 2;                                                \
    }                                                          \
  } while (0)

int main(int argc, char** argv) {
  int (*p_csmri_version)(void);
  const char* (*p_csmri_last_error)(void);
  size_t (*p_csmri_dc_workspace_bytes)(int, int, int);
  int (*p_csmri_fft2)(const float*, float*, int, int, int, int, void*, void*);
  int (*p_csmri_dc_forward_cartesian)(const float*, const float*, const float*, const float*,
                                      float*, int, int, int, void*);
  int (*p_csmri_plane_scale)(const float*, const float*, const float*, float*, int, int, long long,
                             long long, int, void*);
  void* h;
  float buf[8];
  if (argc != 2) return 64;
  h = dlopen(argv[1], RTLD_NOW | RTLD_LOCAL);
  if (!h) {
    fprintf(stderr, "dlopen: %s\n", dlerror());
    return 1;
  }
  GET(csmri_version);
  GET(csmri_last_error);
  GET(csmri_dc_workspace_bytes);
  GET(csmri_fft2);
  GET(csmri_dc_forward_cartesian);
  GET(csmri_plane_scale);
  /* the prototypes above must be the header's: taking the address of the
   * declared functions with these types is a compile-time check */
  (void)sizeof(p_csmri_fft2 == &csmri_fft2);
  (void)sizeof(p_csmri_dc_forward_cartesian == &csmri_dc_forward_cartesian);
  (void)sizeof(p_csmri_plane_scale == &csmri_plane_scale);
  if (p_csmri_version() < 100) return 3;
  if (p_csmri_dc_workspace_bytes(2, 128, 128) != (size_t)2 * 2 * 128 * 128 * 4) return 4;
  if (p_csmri_fft2(buf, buf, 1, 96, 96, 0, buf, NULL) != CSMRI_E_SHAPE) return 5;
  if (!strstr(p_csmri_last_error(), "unsupported")) return 6;
  if (p_csmri_dc_forward_cartesian(NULL, NULL, NULL, NULL, NULL, 1, 256, 256, NULL) !=
      CSMRI_E_NULLPTR)
    return 7;
  if (p_csmri_plane_scale(buf, buf, buf, buf, 1, 8, 8, 8, 7, NULL) != CSMRI_E_ARG) return 8;
  printf("abi ok, version %d\n", p_csmri_version());
  dlclose(h);
  return 0;
}
