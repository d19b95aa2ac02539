// generated: cu16 kernels
#define SP_INST_TAG cu16
#define SP_INST_FMT sp::CU16
#include "sp_inst.cuh"
