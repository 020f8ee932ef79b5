#ifndef SHIM_CAML_THREADS_H
#define SHIM_CAML_THREADS_H
static inline void caml_release_runtime_system(void) {}
static inline void caml_acquire_runtime_system(void) {}
#endif
