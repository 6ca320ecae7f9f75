// Translation unit: the TABLE kernels of kzg_kernels.cuh and their launch wrappers (kzg_launch.h).
#define RK_TU_TABLE
#include "kzg_launch.h"
namespace rk {
RK_KERNELS_TABLE(RK_DEFINE_LAUNCH)
}  // namespace rk
