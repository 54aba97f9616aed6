// Storage for the emulated thread/block indices (see crender_b200/csrc/platform.cuh, CRB_EMU).
#define CRB_EMU 1
#include "../../crender_b200/csrc/platform.cuh"
#undef threadIdx
#undef blockIdx
#undef blockDim
#undef gridDim
namespace crb_emu
{
    dim3 threadIdx_, blockIdx_, blockDim_, gridDim_;
}
