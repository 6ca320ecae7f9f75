// Translation unit: the PAIRING kernel of kzg_kernels.cuh and its launch wrapper (kzg_launch.h).
#define RK_TU_PAIRING
#include "kzg_launch.h"
namespace rk {
static_assert(PAIRING_LINES_BYTES == 2 * PAIRING_STEPS * sizeof(LineStep), "kzg_launch.h: PAIRING_LINES_BYTES out of step with pairing.cuh");
RK_KERNELS_PAIRING(RK_DEFINE_LAUNCH)
}  // namespace rk
