/*
 * TEST INFRASTRUCTURE — NOT PRODUCT CODE.  See msda_oracle_impl.h.
 *
 * Builds the CPU oracle for multi-scale deformable attention in two
 * precisions.  The f64 instance is the parity checker (the reference's own
 * test compares against a float64 run, ops/test.py:34-47); the f32 instance
 * shows what plain float arithmetic gives and is timed by bench.py as a
 * "port" CPU baseline.
 *
 * Build: make -C oracle     (gcc -O2 -fopenmp -shared -fPIC)
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#define REAL double
#define SUFFIX _f64
#define FLOOR floor
#include "msda_oracle_impl.h"
#undef REAL
#undef SUFFIX
#undef FLOOR

#define REAL float
#define SUFFIX _f32
#define FLOOR floorf
#include "msda_oracle_impl.h"
#undef REAL
#undef SUFFIX
#undef FLOOR

int msda_oracle_abi_version(void) { return 1; }
