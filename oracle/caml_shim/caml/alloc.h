/* Minimal stand-in for <caml/alloc.h> (see mlvalues.h in this directory): only what
 * ocaml/soundml_b200_stubs.c names, for its gcc -fsyntax-only check. */
#ifndef SHIM_CAML_ALLOC_H
#define SHIM_CAML_ALLOC_H
#include "mlvalues.h"
value caml_copy_double(double d);
value caml_alloc_tuple(uintnat n);
value caml_copy_nativeint(intnat i);
#endif
