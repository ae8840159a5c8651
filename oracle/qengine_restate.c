/* qengine_restate.c — TEST INFRASTRUCTURE (oracle), NOT product code.  See qengine_restate_impl.h.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load the library built from this file. */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* FP_NORM_EPSILON = epsilon/4 (reference include/common/qrack_types.hpp:263) */
#define REAL float
#define SUF _f32
#define FP_NORM_EPS (1.1920928955078125e-07f / 4)
#include "qengine_restate_impl.h"
#include "qalu_restate_impl.h"
#undef REAL
#undef SUF
#undef FP_NORM_EPS

#define REAL double
#define SUF _f64
#define FP_NORM_EPS (2.220446049250313e-16 / 4)
#include "qengine_restate_impl.h"
#include "qalu_restate_impl.h"
#undef REAL
#undef SUF
#undef FP_NORM_EPS
