// Build shim only (no reference code): HW1 includes "Ray.h" but ships "ray.h"
// (case-sensitive file systems). Used by oracle/build_ref.sh with -I oracle/shim.
#include "ray.h"
