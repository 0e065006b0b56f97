#define DW_T __nv_bfloat16
#define DW_TILED_ENTRY dw_tiled_run_bf16
#include "dwconv_tiled_impl.cuh"
