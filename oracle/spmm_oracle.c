/*
 * spmm_oracle.c — TEST INFRASTRUCTURE ONLY (never on the product path).
 *
 * CPU restatement of the reference's propagation hop,
 *   SSRG/operators/csrc/matmul.c:23-40  FloatCSRMulDenseOMP
 * as the per-element arithmetic it performs when built the way the shipped libmatmul.so was
 * (gcc -O3 -mavx2 -mfma: `answer + coefficient * mat` is contracted into vfmadd):
 *
 *   for every row i, every feature k:   acc = answer[i,k]   (pre-zeroed by the caller)
 *     for j = indptr[i] .. indptr[i+1]-1, in CSR order:  acc = fma(data[j], mat[indices[j],k], acc)
 *
 * Differences from the reference source, none of which change a result bit:
 *   - offsets are 64-bit (the reference's `int` products overflow at N*F >= 2^31, matmul.c:29,33);
 *   - leading dimensions are explicit so padded layouts can be checked;
 *   - fmaf() is spelled out instead of relying on -ffp-contract.
 * Pinned against the real thing: tests/test_oracle.py compares it bit-for-bit with
 * oracle/_ref/libmatmul_ref.so (the reference's own matmul.c compiled in place) and with the
 * golden vectors in tests/golden/.
 */
#include <math.h>
#include <stdint.h>

void oracle_spmm_csr_f32(float *answer, int64_t ld_ans, const float *data, const int32_t *indices,
                         const int32_t *indptr, const float *mat, int64_t ld_mat, int64_t n_rows,
                         int32_t n_feat) {
#pragma omp parallel for schedule(dynamic, 256)
  for (int64_t i = 0; i < n_rows; i++) {
    float *out = answer + i * ld_ans;
    for (int32_t j = indptr[i]; j < indptr[i + 1]; j++) {
      const float *src = mat + (int64_t)indices[j] * ld_mat;
      const float w = data[j];
      for (int32_t k = 0; k < n_feat; k++) out[k] = fmaf(w, src[k], out[k]);
    }
  }
}

/* fp64 hop used by the Chebyshev oracle: separate multiply and add (no contraction), the
 * arithmetic of scipy's csr_matvecs/csc_matvecs axpy as shipped in the x86-64 wheels. */
void oracle_spmm_csr_f64(double *answer, int64_t ld_ans, const double *data, const int32_t *indices,
                         const int32_t *indptr, const double *mat, int64_t ld_mat, int64_t n_rows,
                         int32_t n_feat) {
#pragma omp parallel for schedule(dynamic, 256)
  for (int64_t i = 0; i < n_rows; i++) {
    double *out = answer + i * ld_ans;
    for (int32_t j = indptr[i]; j < indptr[i + 1]; j++) {
      const double *src = mat + (int64_t)indices[j] * ld_mat;
      const double w = data[j];
      for (int32_t k = 0; k < n_feat; k++) {
        volatile double prod = w * src[k];
        out[k] = out[k] + prod;
      }
    }
  }
}
