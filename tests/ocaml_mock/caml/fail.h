/* MOCK of <caml/fail.h> (see mlvalues.h): the two raisers long-jump to the driver's handler. */
#ifndef MOCK_CAML_FAIL_H
#define MOCK_CAML_FAIL_H
#include "mlvalues.h"
CAMLextern void caml_failwith(const char *msg) __attribute__((noreturn));
CAMLextern void caml_invalid_argument(const char *msg) __attribute__((noreturn));
#endif
