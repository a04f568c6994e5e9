// TEST-ONLY: builds the CUDA kernels as host code on top of cuda_emu.h (see that header).
#define CUDA_EMU_IMPLEMENTATION
#include "cuda_emu.h"
#include "../../cavif_rs_b200/csrc/rdo.cu"
#include "../../cavif_rs_b200/csrc/recon.cu"
#undef TXFM1D_GEN_H_
#include "../../cavif_rs_b200/csrc/filters.cu"
