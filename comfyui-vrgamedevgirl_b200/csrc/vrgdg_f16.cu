// kernels and launchers for __half frames
#include "vrgdg_inst.cuh"
namespace vrgdg {
VRGDG_INSTANTIATE(__half)
}
