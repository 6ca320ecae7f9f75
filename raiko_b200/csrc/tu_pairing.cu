// Translation unit: the PAIRING kernel of kzg_kernels.cuh and its launch wrapper (kzg_launch.h).
#define RK_TU_PAIRING
#include "kzg_launch.h"
namespace rk {
RK_KERNELS_PAIRING(RK_DEFINE_LAUNCH)
}  // namespace rk
