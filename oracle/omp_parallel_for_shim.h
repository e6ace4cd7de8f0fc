/* oracle/omp_parallel_for_shim.h -- TEST INFRASTRUCTURE. Force-included (-include) in front of the reference's
 * lib/nnc and lib/nnc/cmd sources when building oracle/_ref/libccv_ref.so.
 *
 * The reference's OpenMP form of parallel_for (lib/ccv_internal.h:28-31) writes the canonical loop as
 * `for ((x) = 0; ...)`; gcc >= 12 rejects a parenthesised iteration variable in `omp parallel for`, which is why
 * lib/configure reports OpenMP "unsupported" on this box.  The sources stay untouched: this shim pulls the
 * header in first (its include guard then keeps the later #include inert) and re-states the same macro without
 * the parentheses, so CPU_REF's loops run on all host cores for the cpu_baseline / --impl reference timings.
 * Loop bodies, scheduling clause and iteration space are exactly the reference's. */
#ifndef ORACLE_OMP_PARALLEL_FOR_SHIM_H
#define ORACLE_OMP_PARALLEL_FOR_SHIM_H
#include "ccv.h"
#include "ccv_internal.h"
#undef parallel_for
#undef parallel_endfor
#undef FOR_IS_PARALLEL
#define parallel_for(x, n) { int x; _Pragma("omp parallel for schedule(dynamic)") for (x = 0; x < (n); x++) {
#define parallel_endfor } }
#define FOR_IS_PARALLEL (1)
#endif
