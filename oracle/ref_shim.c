/*
 * ORACLE -- TEST INFRASTRUCTURE ONLY.
 * Link shim for the ONE reference source that compiles unmodified under gcc:
 * /root/reference/tg/src/util/tg_amanatides_woo.c. It needs three helpers of math/tg_math.c
 * (which itself is MSVC-only: `1ui32` literals etc.), restated here against the reference's own
 * header types so the reference object links: tgm_v3_sub (math/tg_math.c v3 section),
 * tgm_v3_min (:1133-1140), tgm_v3_floor (:1076-1083). Built only into oracle/_ref/ (git-ignored).
 */
#include "math/tg_math.h"
#include <math.h>

v3 tgm_v3_sub(v3 v0, v3 v1) { v3 r; r.x = v0.x - v1.x; r.y = v0.y - v1.y; r.z = v0.z - v1.z; return r; }
v3 tgm_v3_min(v3 v0, v3 v1) { v3 r; r.x = v0.x < v1.x ? v0.x : v1.x; r.y = v0.y < v1.y ? v0.y : v1.y; r.z = v0.z < v1.z ? v0.z : v1.z; return r; }
v3 tgm_v3_floor(v3 v) { v3 r; r.x = floorf(v.x); r.y = floorf(v.y); r.z = floorf(v.z); return r; }
