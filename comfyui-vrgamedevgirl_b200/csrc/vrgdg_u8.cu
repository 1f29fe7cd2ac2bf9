// kernels and launchers for uint8 BGR frames (the reference's cv2 wire format)
#include "vrgdg_inst.cuh"
namespace vrgdg {
VRGDG_INSTANTIATE(uint8_t)
}
