/* Plain C99 consumer of include/sqe_b200.h, linked against libsqe_b200.so: proves the boundary is a
 * C ABI (no C++ types, no torch) and exercises the no-device error paths.  Built (-std=c99 -pedantic
 * -Werror) and run by tests/test_abi_and_host.py. */
#include <stdio.h>
#include <string.h>

#include "sqe_b200.h"

int main(void) {
    int sms = 0, major = 0, minor = 0, rc;
    if (sqe_abi_version() != SQE_ABI_VERSION) return 5;
    /* wrong dim is an argument error on any machine */
    if (sqe_normalize_cast((const float *)16, (void *)16, 4, SQE_DIM / 2, SQE_BF16, NULL) != SQE_E_ARG) return 6;
    if (strstr(sqe_last_error(), "dim") == NULL) return 7;
    if (sqe_topk_gemv_workspace_bytes(1, 10) <= 0) return 8;
    if (sqe_exchange_buffer_bytes(8, 10240) != 256 + 2 * 8 * 10240 * 16) return 10;
    /* the prefiltered scan: k out of range, too many queries, unaligned coarse rows */
    if (sqe_topk_gemv_prefiltered_workspace_bytes(1000, 1, 10) < 1000 * 4) return 11;
    if (sqe_quantize_rows((const void *)16, SQE_BF16, 4, SQE_DIM / 2, (void *)16, (void *)16, NULL) != SQE_E_ARG) return 12;
    if (sqe_topk_gemv_prefiltered((const void *)16, SQE_BF16, 10, SQE_DIM, (const void *)16, (const void *)16,
                                  (const void *)16, 1, 0, (float *)16, (int64_t *)16, 0, NULL, (void *)16,
                                  1 << 20, NULL) != SQE_E_ARG) return 13;
    if (sqe_topk_gemv_prefiltered((const void *)16, SQE_BF16, 10, SQE_DIM, (const void *)16, (const void *)16,
                                  (const void *)16, SQE_MAX_NQ_PREFILTER + 1, 5, (float *)16, (int64_t *)16, 0,
                                  NULL, (void *)16, 1 << 20, NULL) != SQE_E_ARG) return 14;
    if (sqe_topk_gemv_prefiltered((const void *)16, SQE_BF16, 10, SQE_DIM, (const void *)8, (const void *)16,
                                  (const void *)16, 1, 5, (float *)16, (int64_t *)16, 0, NULL, (void *)16,
                                  1 << 20, NULL) != SQE_E_ARG) return 15;
    rc = sqe_device_info(&sms, &major, &minor);
    if (rc == 1) {
        printf("device sm_%d%d with %d SMs\n", major, minor, sms);
    } else {
        /* no sm_100 device: compute entry points must refuse, not fall back */
        if (sqe_normalize_cast((const float *)16, (void *)16, 4, SQE_DIM, SQE_BF16, NULL) != SQE_E_CUDA) return 9;
        printf("no sm_100 device: compute calls return SQE_E_CUDA (%s)\n", sqe_last_error());
    }
    printf("abi %d ok\n", sqe_abi_version());
    return 0;
}
