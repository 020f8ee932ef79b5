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
enum caml_ba_layout { CAML_BA_C_LAYOUT = 0, CAML_BA_FORTRAN_LAYOUT = 0x100 };
enum caml_ba_managed { CAML_BA_EXTERNAL = 0, CAML_BA_MANAGED = 0x200, CAML_BA_MAPPED_FILE = 0x400 };
value caml_ba_alloc(int flags, int num_dims, void *data, intnat *dim);
/* the shim passes a pointer to the descriptor itself as the OCaml value */
#define Caml_ba_array_val(v) ((struct caml_ba_array *)(v))
#define Caml_ba_data_val(v) (Caml_ba_array_val(v)->data)
#endif
