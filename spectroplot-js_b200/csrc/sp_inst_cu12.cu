// generated: cu12 kernels
#define SP_INST_TAG cu12
#define SP_INST_FMT sp::CU12
#include "sp_inst.cuh"
