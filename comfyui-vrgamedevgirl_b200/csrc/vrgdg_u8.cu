// kernels and launchers for uint8 BGR frames (the reference's cv2 wire format), incl. the fixed-point Lanczos4 resize
#define VRGDG_LANCZOS_IMPL
#include "vrgdg_inst.cuh"
#include "vrgdg_lanczos.cuh"
namespace vrgdg {
VRGDG_INSTANTIATE(uint8_t)
}
