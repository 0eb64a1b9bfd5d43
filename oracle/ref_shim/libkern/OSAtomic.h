/* TEST INFRASTRUCTURE ONLY. Stand-in for the macOS header used by the reference's
 * Threading.cpp:81 (OSAtomicIncrement32 returns the NEW value). */
#ifndef RACC_REF_SHIM_OSATOMIC_H
#define RACC_REF_SHIM_OSATOMIC_H
static inline int OSAtomicIncrement32(volatile int* p) { return __sync_add_and_fetch(p, 1); }
#endif
