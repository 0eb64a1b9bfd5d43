/* TEST INFRASTRUCTURE ONLY. Force-included AFTER <unistd.h> when compiling the reference.
 *
 * The reference sizes its builder thread pool from sysconf(_SC_NPROCESSORS_ONLN)
 * (RayAccelerator/Threading.cpp:83-98). Its bounds pass spawns that many copies of one task and
 * each copy OVERWRITES its executor's slot with the bounds it accumulated (Bvh2.cpp:697-698); when
 * one pool thread happens to run two copies, the second (which finds no work left) replaces the
 * first's bounds with -inf, so the root box -- and with it the root's split decision -- depends on
 * thread timing (always wrong for scenes of a few hundred triangles). Running the reference with
 * ONE builder thread removes the race without touching its source: the tree it then builds is the
 * one its algorithm defines. RACC_REF_THREADS overrides (e.g. to reproduce the race).
 */
#ifndef RACC_REF_SHIM_CPU_COUNT_H
#define RACC_REF_SHIM_CPU_COUNT_H
#include <stdlib.h>
static inline long racc_ref_shim_cpu_count(void) {
	const char* v = getenv("RACC_REF_THREADS");
	long n = v ? atol(v) : 1;
	return n > 0 ? n : 1;
}
#define sysconf(name) racc_ref_shim_cpu_count()
#endif
