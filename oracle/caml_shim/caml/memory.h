/* Minimal stand-in for <caml/memory.h> (see mlvalues.h in this directory): the root
 * registration macros expand to nothing but a use of their arguments. */
#ifndef SHIM_CAML_MEMORY_H
#define SHIM_CAML_MEMORY_H
#include "mlvalues.h"
#define CAMLparam0() do { } while (0)
#define CAMLparam1(a) (void)(a)
#define CAMLparam2(a, b) (void)(a), (void)(b)
#define CAMLparam3(a, b, c) (void)(a), (void)(b), (void)(c)
#define CAMLparam4(a, b, c, d) (void)(a), (void)(b), (void)(c), (void)(d)
#define CAMLparam5(a, b, c, d, e) (void)(a), (void)(b), (void)(c), (void)(d), (void)(e)
#define CAMLlocal1(a) value a = Val_unit
#define CAMLreturn(x) return (x)
#define Store_field(block, i, v) (((value *)(block))[i] = (v))
#endif
