// Translation unit: the MSM kernels of kzg_kernels.cuh and their launch wrappers (kzg_launch.h).
#define RK_TU_MSM
#include "kzg_launch.h"
namespace rk {
RK_KERNELS_MSM(RK_DEFINE_LAUNCH)
cudaError_t configure_k_msm_affine() {
    cudaError_t e = cudaFuncSetAttribute(k_msm_affine, cudaFuncAttributeMaxDynamicSharedMemorySize, MSM_AFF_MAX_K * 256 * 4);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_msm_affine_w12, cudaFuncAttributeMaxDynamicSharedMemorySize, MSM_AFF_MAX_K * 384 * 4);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_msm_affine_w16, cudaFuncAttributeMaxDynamicSharedMemorySize, MSM_AFF_MAX_K * 512 * 4);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_msm_affine_w16n, cudaFuncAttributeMaxDynamicSharedMemorySize, MSM_AFF_MAX_K * 512 * 4);
    return e;
}
}  // namespace rk
