/* Minimal stand-in for <caml/bigarray.h> (see mlvalues.h in this directory). */
#ifndef SHIM_CAML_BIGARRAY_H
#define SHIM_CAML_BIGARRAY_H
#include "mlvalues.h"
struct caml_ba_array {
  void *data;
  intnat num_dims;
  intnat flags;
  void *proxy;
  intnat dim[1];
};
enum caml_ba_kind {
  CAML_BA_FLOAT32 = 0, CAML_BA_FLOAT64 = 1, CAML_BA_COMPLEX32 = 10, CAML_BA_COMPLEX64 = 11,
  CAML_BA_KIND_MASK = 0xFF
};
/* the shim passes a pointer to the descriptor itself as the OCaml value */
#define Caml_ba_array_val(v) ((struct caml_ba_array *)(v))
#define Caml_ba_data_val(v) (Caml_ba_array_val(v)->data)
#endif
