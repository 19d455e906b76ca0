// ls_host.h -- host-side helpers shared by the translation units of libls_raster.so.
#pragma once
#include "ls_raster.h"
int ls_fail(const char* fmt, ...);         // sets the thread-local error string, returns -1
int ls_check_cuda(const char* what);       // cudaGetLastError -> ls_fail
int ls_validate_scene(const LsRasterScene* sc);
