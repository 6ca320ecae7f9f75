// Translation unit: the MSM kernels of kzg_kernels.cuh and their launch wrappers (kzg_launch.h).
#define RK_TU_MSM
#include "kzg_launch.h"
namespace rk {
RK_KERNELS_MSM(RK_DEFINE_LAUNCH)
cudaError_t configure_k_msm_affine() {
    return cudaFuncSetAttribute(k_msm_affine, cudaFuncAttributeMaxDynamicSharedMemorySize, MSM_AFF_MAX_K * MSM_AFF_THREADS * 4);
}
}  // namespace rk
