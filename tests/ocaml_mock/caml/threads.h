/* MOCK of <caml/threads.h> (see mlvalues.h): the driver counts the calls and checks that they pair up. */
#ifndef MOCK_CAML_THREADS_H
#define MOCK_CAML_THREADS_H
#include "mlvalues.h"
CAMLextern void caml_release_runtime_system(void);
CAMLextern void caml_acquire_runtime_system(void);
#endif
