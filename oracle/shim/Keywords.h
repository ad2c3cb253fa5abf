/* oracle/shim/Keywords.h -- TEST INFRASTRUCTURE. read_scene.c:4 includes "Keywords.h" with a
 * capital K (Windows file systems are case-insensitive); the file is src/keywords.h. */
#include "keywords.h"
