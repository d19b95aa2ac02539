// generated: cs16 kernels
#define SP_INST_TAG cs16
#define SP_INST_FMT sp::CS16
#include "sp_inst.cuh"
