// generated: rt kernels
#define SP_INST_TAG rt
#define SP_INST_FMT sp::FMT_RUNTIME
#include "sp_inst.cuh"
