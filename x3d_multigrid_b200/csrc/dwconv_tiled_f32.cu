#define DW_T float
#define DW_TILED_ENTRY dw_tiled_run_f32
#include "dwconv_tiled_impl.cuh"
