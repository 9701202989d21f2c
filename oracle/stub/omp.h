/* Stub <omp.h> for the serial-semantics build of the reference (oracle/_ref/ag_ref).
 * TEST INFRASTRUCTURE ONLY.  The reference (Tree.cpp:44-48, Node.cpp:420,457) only needs
 * omp_get_max_threads() for its `cores*100` bulk/serial insertion switch; with OpenMP
 * pragmas ignored the code is single-threaded and bit-reproducible. */
#pragma once
#ifdef __cplusplus
extern "C" {
#endif
extern int ag_stub_cores;
static inline int  omp_get_max_threads(void) { return ag_stub_cores; }
static inline int  omp_get_thread_num(void)  { return 0; }
static inline int  omp_get_num_threads(void) { return 1; }
static inline void omp_set_num_threads(int n) { (void)n; }
static inline void omp_set_nested(int n)      { (void)n; }
#ifdef __cplusplus
}
#endif
