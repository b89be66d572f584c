/* oracle/model_unit.c -- TEST INFRASTRUCTURE.  Wraps one ACE-emitted model translation unit
 * (a checked-in fhe-cmplr/rtlib/ant/dataset/<model>.onnx.inc, #included unmodified from where
 * it lies) so that it can be loaded next to oracle/_ref/libace_ref.so: the callbacks every
 * emitted unit defines are renamed, and model_register() hands them to the harness
 * (ref_set_callbacks in oracle/ref_harness.c).  Compiled against the REFERENCE's headers.
 *   cc -DMODEL_INC='"<path>.onnx.inc"' -shared model_unit.c -L_ref -lace_ref
 */
#define Get_context_params Emitted_get_context_params
#define Get_rt_data_info Emitted_get_rt_data_info
#define Get_input_count Emitted_get_input_count
#define Get_output_count Emitted_get_output_count
#define Get_encode_scheme Emitted_get_encode_scheme
#define Get_decode_scheme Emitted_get_decode_scheme
#define Main_graph Emitted_main_graph

#include "common/rtlib.h"
#include MODEL_INC

void model_register(void (*set)(void*, void*, void*, void*, void*)) {
  set((void*)Emitted_get_context_params, (void*)Emitted_get_rt_data_info, (void*)Emitted_main_graph,
      (void*)Emitted_get_encode_scheme, (void*)Emitted_get_decode_scheme);
}
