// generated: cs4 kernels
#define SP_INST_TAG cs4
#define SP_INST_FMT sp::CS4
#include "sp_inst.cuh"
