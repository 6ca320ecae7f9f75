// Typed launch wrappers: the kernels are compiled in five translation units (msm / path /
// table / verify / pairing) so the build parallelises; the host code in kzg_ctx.cu launches them through
// these functions.  One X-macro table per unit keeps declaration and definition in step.
#pragma once
#include "kzg_kernels.cuh"

namespace rk {

struct LineStep;                                   // pairing.cuh (only the verify / pairing units see its definition)
constexpr size_t PAIRING_LINES_BYTES = 2 * 68 * 4 * sizeof(Fp);   // 2 fixed G2 points x PAIRING_STEPS x LineStep

#define RK_UNPAREN(...) __VA_ARGS__

#define RK_KERNELS_MSM(X)                                                                      \
    X(k_msm, (MsmParams p), (p))                                                               \
    X(k_msm_affine, (MsmAffParams p), (p))
#define RK_KERNELS_PATH(X)                                                                     \
    X(k_finalize, (const G1Xyzz* partials, int splits, int nblobs, const uint32_t* bad, uint8_t* out_g1, uint8_t* out_vh, uint8_t* status, int stride), \
      (partials, splits, nblobs, bad, out_g1, out_vh, status, stride))                         \
    X(k_finalize_warp, (const G1Xyzz* partials, int splits, int nblobs, const uint32_t* bad, uint8_t* out_g1, uint8_t* out_vh, uint8_t* status, int stride), \
      (partials, splits, nblobs, bad, out_g1, out_vh, status, stride))                         \
    X(k_status_only, (const uint32_t* bad, int nblobs, uint8_t* rec, uint8_t* status, int stride, int zero_bytes), (bad, nblobs, rec, status, stride, zero_bytes)) \
    X(k_sha_blob, (const uint8_t* blobs, int nblobs, uint8_t* out_hash, int stride), (blobs, nblobs, out_hash, stride)) \
    X(k_sha_blob_duo, (const uint8_t* blobs, int nblobs, uint8_t* out_hash, int stride), (blobs, nblobs, out_hash, stride)) \
    X(k_fr_eval_quot, (FrParams p), (p))                                                       \
    X(k_imad_peak, (uint64_t* out, uint32_t seed, int iters), (out, seed, iters))               \
    X(k_synth_blobs, (uint64_t seed, uint32_t first_blob, uint32_t nblobs, uint8_t* out), (seed, first_blob, nblobs, out)) \
    X(k_decode_blob_data, (const uint8_t* blobs, int nblobs, uint8_t* out, uint32_t* out_len), (blobs, nblobs, out, out_len))

#define RK_KERNELS_TABLE(X)                                                                    \
    X(k_setup_decompress, (const uint8_t* in, int n, G1Affine* out, int* err), (in, n, out, err)) \
    X(k_setup_from_ref, (const uint8_t* in, int n, G1Affine* out, int* err), (in, n, out, err)) \
    X(k_setup_to_ref, (const G1Affine* in, int n, uint8_t* out), (in, n, out))                  \
    X(k_fp_be_to_ref, (const uint8_t* in, int n, uint8_t* out), (in, n, out))                   \
    X(k_fp_ref_to_be, (const uint8_t* in, int n, uint8_t* out), (in, n, out))                   \
    X(k_roots_brp, (Fr* out), (out))                                                           \
    X(k_roots_export, (int order, int n, uint8_t* out), (order, n, out))                        \
    X(k_table_bases, (const G1Affine* g1, TableGeom g, G1Xyzz* bases), (g1, g, bases))          \
    X(k_table_bases_affine, (const G1Xyzz* bases, int n, G1Affine* out), (bases, n, out))       \
    X(k_table_chain, (const G1Affine* bases_aff, TableGeom g, uint32_t d0, int D, G1Xyzz* state, G1Xyzz* tmp), (bases_aff, g, d0, D, state, tmp)) \
    X(k_table_normalize, (TableGeom g, uint32_t d0, int D, const G1Xyzz* tmp, TableEntry* table), (g, d0, D, tmp, table))

#define RK_KERNELS_VERIFY(X)                                                                   \
    X(k_sha_fs_challenge, (const uint8_t* blobs, const uint8_t* commitments, int nblobs, uint8_t* out_z), (blobs, commitments, nblobs, out_z)) \
    X(k_g1_decompress_validate, (const uint8_t* in, int n, G1Affine* out, int* out_inf, int* err), (in, n, out, out_inf, err)) \
    X(k_batch_challenge, (const uint8_t* c, const uint8_t* z, const uint8_t* y, const uint8_t* pr, int n, Fr* out_r), (c, z, y, pr, n, out_r)) \
    X(k_verify_pow128, (const G1Affine* pts, const int* inf, int n, G1Affine* hi), (pts, inf, n, hi)) \
    X(k_verify_terms, (const Fr* r_mont, const uint8_t* z, const uint8_t* y, const G1Affine* cs, const int* c_inf, const G1Affine* ps, const int* p_inf, int n, const G1Affine* cs_hi, const G1Affine* ps_hi, const G1Affine* g_hi, G1Xyzz* out_a, G1Xyzz* out_e, int* err), \
      (r_mont, z, y, cs, c_inf, ps, p_inf, n, cs_hi, ps_hi, g_hi, out_a, out_e, err))           \
    X(k_verify_reduce_partial, (const G1Xyzz* a, int na, const G1Xyzz* e, int ne, G1Xyzz* part_a, G1Xyzz* part_e), (a, na, e, ne, part_a, part_e)) \
    X(k_verify_reduce, (const G1Xyzz* part_a, const G1Xyzz* part_e, int nparts, G1Affine* out_pts, int* out_inf), (part_a, part_e, nparts, out_pts, out_inf))

#define RK_KERNELS_PAIRING(X)                                                                  \
    X(k_pairing_check, (const G1Affine* pts, const int* inf, const uint8_t* g2_s_be, const uint8_t* g2_gen_be, int* out_ok), (pts, inf, g2_s_be, g2_gen_be, out_ok)) \
    X(k_pairing_precompute, (const uint8_t* g2_s_be, const uint8_t* g2_gen_be, LineStep* out), (g2_s_be, g2_gen_be, out)) \
    X(k_pairing_check_lanes, (const G1Affine* pts, const int* inf, const LineStep* lines, int* out_ok), (pts, inf, lines, out_ok)) \
    X(k_pairing_check_cta, (const G1Affine* pts, const int* inf, const LineStep* lines, int* out_ok), (pts, inf, lines, out_ok))

#define RK_DECLARE_LAUNCH(name, params, args) \
    void launch_##name(dim3 grid, dim3 block, size_t smem, cudaStream_t st, RK_UNPAREN params);
#define RK_DEFINE_LAUNCH(name, params, args) \
    void launch_##name(dim3 grid, dim3 block, size_t smem, cudaStream_t st, RK_UNPAREN params) { name<<<grid, block, smem, st>>> args; }

RK_KERNELS_MSM(RK_DECLARE_LAUNCH)
RK_KERNELS_PATH(RK_DECLARE_LAUNCH)
RK_KERNELS_TABLE(RK_DECLARE_LAUNCH)
RK_KERNELS_VERIFY(RK_DECLARE_LAUNCH)
RK_KERNELS_PAIRING(RK_DECLARE_LAUNCH)
// opts k_msm_affine into its dynamic shared memory (digit codes: 4 B x 64 chains x block size)
cudaError_t configure_k_msm_affine();

}  // namespace rk
