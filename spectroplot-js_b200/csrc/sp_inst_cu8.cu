// generated: cu8 kernels
#define SP_INST_TAG cu8
#define SP_INST_FMT sp::CU8
#include "sp_inst.cuh"
