/* Minimal stand-in for <caml/mlvalues.h>: just enough of the OCaml runtime's C
 * interface to compile the reference's resample_stubs.c unmodified, outside
 * OCaml.  TEST INFRASTRUCTURE ONLY (see oracle/Makefile). */
#ifndef SHIM_CAML_MLVALUES_H
#define SHIM_CAML_MLVALUES_H
#include <stdint.h>
#include <stddef.h>
typedef intptr_t value;
typedef intptr_t intnat;
typedef uintptr_t uintnat;
#define Val_long(x) ((value)(((intnat)(x) << 1) + 1))
#define Long_val(v) ((intnat)(v) >> 1)
#define Val_bool(x) Val_long((x) != 0)
#define Bool_val(v) (Long_val(v) != 0)
#define Val_unit Val_long(0)
#define Int_val(v) ((int)Long_val(v))
#define Val_int(x) Val_long(x)
#define Double_val(v) (*(double *)(v))
#define CAMLprim
#endif
