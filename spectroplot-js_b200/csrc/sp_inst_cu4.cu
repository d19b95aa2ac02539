// generated: cu4 kernels
#define SP_INST_TAG cu4
#define SP_INST_FMT sp::CU4
#include "sp_inst.cuh"
