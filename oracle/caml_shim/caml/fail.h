#ifndef SHIM_CAML_FAIL_H
#define SHIM_CAML_FAIL_H
/* raises through the driver's jump buffer (ref_driver.c) */
void caml_failwith(const char *msg) __attribute__((noreturn));
void caml_invalid_argument(const char *msg) __attribute__((noreturn));
void caml_raise_out_of_memory(void) __attribute__((noreturn));
#endif
