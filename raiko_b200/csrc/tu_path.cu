// Translation unit: the PATH kernels of kzg_kernels.cuh and their launch wrappers (kzg_launch.h).
#define RK_TU_PATH
#include "kzg_launch.h"
namespace rk {
RK_KERNELS_PATH(RK_DEFINE_LAUNCH)
}  // namespace rk
