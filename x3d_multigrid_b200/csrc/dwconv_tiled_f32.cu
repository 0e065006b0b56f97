#define DW_T float
#define DW_TILED_ENTRY dw_tiled_run_f32
#define DW_WGRAD_ENTRY dw_wgrad_tiled_f32
#include "dwconv_tiled_impl.cuh"
