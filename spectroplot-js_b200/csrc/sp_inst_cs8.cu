// generated: cs8 kernels
#define SP_INST_TAG cs8
#define SP_INST_FMT sp::CS8
#include "sp_inst.cuh"
