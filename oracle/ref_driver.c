/* Driver for the reference's own resampler executors, compiled from
 * /root/reference/soundml/lib/resample_stubs.c (unmodified, where it lies) with
 * the OCaml headers replaced by oracle/caml_shim.  TEST INFRASTRUCTURE ONLY.
 *
 * Exposes plain-C entry points that build the Bigarray descriptors the stubs
 * expect and call the reference's CAMLprim functions:
 *   soundml_resample_step   resample_stubs.c:232-297  (polyphase dot executor)
 *   soundml_resample_shape  resample_stubs.c:376-408  (OLS spectrum shaping)
 */
#include <setjmp.h>
#include <stdint.h>
#include <string.h>

#include <caml/bigarray.h>
#include <caml/mlvalues.h>

value soundml_resample_step(value, value, value, value, value, value, value, value, value,
                            value, value, value, value, value, value, value, value);
value soundml_resample_shape(value, value, value, value, value, value, value);

static jmp_buf fail_jmp;
static char fail_msg[256];

void caml_failwith(const char *msg) {
  strncpy(fail_msg, msg, sizeof fail_msg - 1);
  longjmp(fail_jmp, 1);
}

const char *ref_last_error(void) { return fail_msg; }

static struct caml_ba_array ba(void *data, int64_t len, int kind) {
  struct caml_ba_array a;
  a.data = data;
  a.num_dims = 1;
  a.flags = kind;
  a.proxy = 0;
  a.dim[0] = len;
  return a;
}

/* kind: 0 = float32, 1 = float64.  Lengths are element counts. */
int ref_resample_step(int kind, void *bank, int64_t bank_len, void *hist, int64_t hist_len,
                      void *scratch, int64_t scratch_len, void *x, int64_t x_len, void *y,
                      int64_t y_len, int64_t n, int64_t n_out, int64_t channels, int64_t k,
                      int64_t l, int64_t m, int64_t row0, int64_t s0, int64_t y_off,
                      int64_t y_stride, int visit, int is_flush) {
  struct caml_ba_array b = ba(bank, bank_len, kind), h = ba(hist, hist_len, kind),
                       s = ba(scratch, scratch_len, kind), xi = ba(x, x_len, kind),
                       yo = ba(y, y_len, kind);
  if (setjmp(fail_jmp)) return 1;
  soundml_resample_step((value)&b, (value)&h, (value)&s, (value)&xi, (value)&yo, Val_long(n),
                        Val_long(n_out), Val_long(channels), Val_long(k), Val_long(l),
                        Val_long(m), Val_long(row0), Val_long(s0), Val_long(y_off),
                        Val_long(y_stride), Val_bool(visit), Val_bool(is_flush));
  return 0;
}

/* complex128 lines: x [lines, n/2+1], h plan spectrum, y [lines, w/2+1]. */
int ref_resample_shape(void *x, int64_t x_len, void *h, int64_t h_len, void *y, int64_t y_len,
                       int64_t lines, int64_t n, int64_t sl, int64_t sm) {
  struct caml_ba_array xa = ba(x, x_len, CAML_BA_COMPLEX64), ha = ba(h, h_len, CAML_BA_COMPLEX64),
                       ya = ba(y, y_len, CAML_BA_COMPLEX64);
  if (setjmp(fail_jmp)) return 1;
  soundml_resample_shape((value)&xa, (value)&ha, (value)&ya, Val_long(lines), Val_long(n),
                         Val_long(sl), Val_long(sm));
  return 0;
}
