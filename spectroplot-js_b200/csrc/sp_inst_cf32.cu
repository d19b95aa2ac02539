// generated: cf32 kernels
#define SP_INST_TAG cf32
#define SP_INST_FMT sp::CF32
#include "sp_inst.cuh"
