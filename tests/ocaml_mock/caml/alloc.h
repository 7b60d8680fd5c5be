/* MOCK of <caml/alloc.h> (see mlvalues.h). */
#ifndef MOCK_CAML_ALLOC_H
#define MOCK_CAML_ALLOC_H
#include "mlvalues.h"
CAMLextern value caml_copy_double(double d);
CAMLextern value caml_copy_int64(int64_t i);
CAMLextern value caml_copy_string(const char *s);
CAMLextern value caml_alloc_tuple(uintnat n);
#endif
