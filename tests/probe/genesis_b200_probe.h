/* TEST INFRASTRUCTURE -- tcgen05 probes (tests/probe/debug_umma.cu -> tests/probe/libgenesis_b200_probe.so, built by
 * genesis_b200/build.py:build_probe).  Not part of the product library: tests/test_umma_layouts_gpu.py pins the shared-memory
 * descriptor conventions the production kernels rely on, scripts/umma_rate.py measures the instruction cost table. */
#pragma once
#include <stdint.h>
typedef void* g2_stream_t;
#ifdef __cplusplus
extern "C" {
#endif

/* UMMA descriptor self-test (tests/probe/debug_umma.cu): runs `nk` tcgen05.mma.kind::tf32 (M=128) on caller-provided
 * shared-memory images of A and B with caller-provided descriptor templates and dumps D[128][N]. */
int g2_debug_umma_probe(const float* a_img, const float* b_img, float* D, int a_bytes, int b_bytes, long adesc_t,
                        long bdesc_t, int idesc, int N, int nk, int a_kstep, int b_kstep, int a_off, int b_off,
                        int base_off_auto, g2_stream_t stream);
/* UMMA issue-rate probe (tests/probe/debug_umma.cu): out[148 * ctas_per_sm] = clocks for n_mma back-to-back tcgen05.mma.kind::tf32
 * (M = 128, N, K = 8; 4 K-steps per A start row, then the row advances by a_shift_rows) per CTA, operands in shared memory. */
int g2_debug_umma_rate(int64_t* out, int N, int n_mma, int a_shift_rows, int n_acc, int a_rows, int ctas_per_sm,
                       int b_tiles, g2_stream_t stream);


#ifdef __cplusplus
}
#endif
