/* MOCK of <caml/bigarray.h> (see mlvalues.h): the descriptor's public fields, data / num_dims / flags / dim[]. */
#ifndef MOCK_CAML_BIGARRAY_H
#define MOCK_CAML_BIGARRAY_H
#include "mlvalues.h"

enum caml_ba_kind { CAML_BA_FLOAT32, CAML_BA_FLOAT64, CAML_BA_SINT8, CAML_BA_UINT8, CAML_BA_SINT16, CAML_BA_UINT16,
                    CAML_BA_INT32, CAML_BA_INT64 };
enum caml_ba_layout { CAML_BA_C_LAYOUT = 0, CAML_BA_FORTRAN_LAYOUT = 0x100 };

struct caml_ba_array {
    void *data;
    intnat num_dims;
    intnat flags;
    void *proxy;
    intnat dim[4];
};
CAMLextern uintnat caml_ba_byte_size(struct caml_ba_array *b);     /* payload size in bytes */
#define CAML_BA_KIND_MASK 0xff
#define Caml_ba_array_val(v) ((struct caml_ba_array *)(v))
#define Caml_ba_data_val(v) (Caml_ba_array_val(v)->data)
#endif
