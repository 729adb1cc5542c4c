/* Plain C99 client of the group entry points of include/satmc.h: the multi-GPU split of the path, the collective
 * (NCCL) inside the library.  One process drives every GPU of the box (satmc_group_create); both shard modes must return
 * the counts of a single-device call.  Compiled (not run) on CPU by tests/test_abi.py, run on the GPU box by
 * tests/test_gpu_group.py. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "satmc.h"

#define N_PAIRS 1001

int main(int argc, char** argv)
{
    int n_dev = argc > 1 ? atoi(argv[1]) : 0;
    satmc_ctx* ctx = NULL;
    int rc = satmc_create(0, NULL, &ctx);
    if (rc != SATMC_OK) {
        fprintf(stderr, "satmc_create: %d %s\n", rc, satmc_last_error(NULL));
        return rc == SATMC_ERR_NO_DEVICE ? 77 : 1;
    }
    if (n_dev <= 0) {                                          /* every device the library accepts, up to 8 */
        satmc_ctx* probe = NULL;
        n_dev = 1;
        while (n_dev < 8 && satmc_create(n_dev, NULL, &probe) == SATMC_OK) { satmc_destroy(probe); n_dev++; }
    }
    static satmc_pair pairs[N_PAIRS];
    static uint64_t want[N_PAIRS], got[N_PAIRS];
    for (int i = 0; i < N_PAIRS; i++) {
        satmc_pair* p = &pairs[i];
        const float a = 6.2831853f * (float)i / N_PAIRS;
        p->rx = 5.5f * cosf(a); p->ry = 4.5f * sinf(a); p->rtheta = 3.0f * a; p->rw = 4.07f; p->rh = 1.74f;
        p->ow = 1.0f + (float)(i % 7) * 0.5f; p->oh = 0.5f + (float)(i % 5) * 0.7f;
        p->sd_x = 0.1f + 0.4f * (float)(i % 3); p->sd_y = 0.3f; p->sd_theta = 0.05f * (float)(i % 11); p->sd_w = 0.0f; p->sd_h = 0.0f;
    }
    const uint64_t n = 20001, seed = 99, offset = 7000000000ull;
    rc = satmc_count_fused_host(ctx, pairs, N_PAIRS, n, seed, offset, 3, want, 0);
    if (rc != SATMC_OK) { fprintf(stderr, "single: %s\n", satmc_last_error(ctx)); return 1; }
    satmc_group* g = NULL;
    rc = satmc_group_create(NULL, n_dev, &g);
    if (rc != SATMC_OK) { fprintf(stderr, "satmc_group_create(%d): %d %s\n", n_dev, rc, satmc_group_last_error(NULL)); return 1; }
    const int modes[2] = {SATMC_SHARD_BY_PAIR, SATMC_SHARD_BY_SAMPLE_RANGE};
    for (int m = 0; m < 2; m++) {
        memset(got, 0xff, sizeof(got));
        rc = satmc_group_count_fused_host(g, pairs, N_PAIRS, n, seed, offset, 3, modes[m], got, 0);
        if (rc != SATMC_OK) { fprintf(stderr, "group mode %d: %s\n", modes[m], satmc_group_last_error(g)); return 1; }
        if (memcmp(got, want, sizeof(got)) != 0) { fprintf(stderr, "mode %d: counts differ from the single-device call\n", modes[m]); return 1; }
    }
    uint64_t lo = 0, hi = 0, total = 0;
    for (int r = 0; r < n_dev; r++) { satmc_shard_range(SATMC_SHARD_BY_SAMPLE_RANGE, n, n_dev, r, &lo, &hi); total += hi - lo; }
    printf("group of %d device(s), NCCL %d: by pair and by sample range identical to one device (%llu hits of pair 0, ranges cover %llu)\n",
           satmc_group_world(g), satmc_group_nccl_version(), (unsigned long long)want[0], (unsigned long long)total);
    satmc_group_destroy(g);
    satmc_destroy(ctx);
    return total == n ? 0 : 1;
}
