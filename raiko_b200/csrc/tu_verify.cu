// Translation unit: the VERIFY kernels of kzg_kernels.cuh and their launch wrappers (kzg_launch.h).
#define RK_TU_VERIFY
#include "kzg_launch.h"
namespace rk {
RK_KERNELS_VERIFY(RK_DEFINE_LAUNCH)
}  // namespace rk
