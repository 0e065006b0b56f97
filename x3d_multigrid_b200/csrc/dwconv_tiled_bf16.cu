#define DW_T __nv_bfloat16
#define DW_TILED_ENTRY dw_tiled_run_bf16
#define DW_WGRAD_ENTRY dw_wgrad_tiled_bf16
#include "dwconv_tiled_impl.cuh"
