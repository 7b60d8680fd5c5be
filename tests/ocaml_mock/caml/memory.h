/* MOCK of <caml/memory.h> (see mlvalues.h): no GC, so registering roots is a no-op that keeps the syntax
 * and still makes the compiler check that every name given to CAMLparam / CAMLlocal is a `value`. */
#ifndef MOCK_CAML_MEMORY_H
#define MOCK_CAML_MEMORY_H
#include "mlvalues.h"

#define MOCK_ROOT(x) do { const value *mock_root_ = &(x); (void)mock_root_; } while (0)
#define CAMLparam0() int caml_mock_frame_ = 0; (void)caml_mock_frame_
#define CAMLparam1(a) CAMLparam0(); MOCK_ROOT(a)
#define CAMLparam2(a, b) CAMLparam1(a); MOCK_ROOT(b)
#define CAMLparam3(a, b, c) CAMLparam2(a, b); MOCK_ROOT(c)
#define CAMLparam4(a, b, c, d) CAMLparam3(a, b, c); MOCK_ROOT(d)
#define CAMLparam5(a, b, c, d, e) CAMLparam4(a, b, c, d); MOCK_ROOT(e)
#define CAMLxparam1(a) MOCK_ROOT(a)
#define CAMLxparam2(a, b) MOCK_ROOT(a); MOCK_ROOT(b)
#define CAMLxparam3(a, b, c) CAMLxparam2(a, b); MOCK_ROOT(c)
#define CAMLlocal1(a) value a = Val_unit
#define CAMLlocal2(a, b) value a = Val_unit, b = Val_unit
#define CAMLreturn(x) do { (void)caml_mock_frame_; return (x); } while (0)
#define CAMLreturn0 do { (void)caml_mock_frame_; return; } while (0)
#endif
