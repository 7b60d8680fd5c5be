/*
 * MOCK of the OCaml runtime's <caml/mlvalues.h> — test infrastructure only.
 *
 * The build image has no OCaml toolchain, so ocaml/phylo_b200_stubs.c cannot be compiled against the
 * real runtime here.  These headers declare just enough of the runtime's C interface (the macros and
 * functions the stubs use, with the real names, arities and value encoding of immediates) for the stubs
 * to be COMPILED with -Wall -Wextra -Werror and EXECUTED by tests/ocaml_mock/stubs_driver.c against
 * libphylo_b200.so.  Written from the published OCaml manual ("Interfacing C with OCaml"); nothing here
 * comes from the OCaml sources.  Blocks are plain malloc'd C objects and there is no garbage collector:
 * the CAMLparam/CAMLlocal/CAMLreturn macros only keep their syntax.
 */
#ifndef MOCK_CAML_MLVALUES_H
#define MOCK_CAML_MLVALUES_H

#include <stdint.h>

typedef intptr_t value;
typedef intptr_t intnat;
typedef uintptr_t uintnat;

#define CAMLprim
#define CAMLextern extern

/* immediates: 2n + 1, as in the real runtime */
#define Val_long(x) ((value)(((uintnat)(intnat)(x) << 1) + 1))
#define Long_val(v) ((intnat)(v) >> 1)
#define Val_int(x) Val_long(x)
#define Int_val(v) ((int)Long_val(v))
#define Val_unit Val_long(0)
#define Val_bool(b) Val_long((b) != 0)
#define Bool_val(v) Int_val(v)
#define Val_true Val_long(1)
#define Val_false Val_long(0)
#define Is_long(v) (((v) & 1) != 0)

/* boxed float / int64: here a pointer to a malloc'd scalar */
#define Double_val(v) (*(const double *)(v))
#define Int64_val(v) (*(const int64_t *)(v))
/* fields of a block (tuples): here a malloc'd array of values */
#define Field(v, i) (((value *)(v))[i])
#define Store_field(v, i, x) (((value *)(v))[i] = (x))
/* OCaml string: here a NUL-terminated C string */
#define String_val(v) ((const char *)(v))

#endif
