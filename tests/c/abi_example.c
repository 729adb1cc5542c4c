/* Plain C99 client of include/satmc.h: proves the header is valid C and the library links from C.
 * Compiled and linked (not run) by tests/test_abi.py on CPU; run on the GPU box by tests/test_gpu_programs.py. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "satmc.h"

int main(void)
{
    satmc_ctx* ctx = NULL;
    int rc = satmc_create(0, NULL, &ctx);
    if (rc != SATMC_OK) {
        fprintf(stderr, "satmc_create: %d %s\n", rc, satmc_last_error(NULL));
        return rc == SATMC_ERR_NO_DEVICE ? 77 : 1;          /* 77 = no GPU: the library never computes on the CPU */
    }
    satmc_pair p;
    p.rx = 3.1f; p.ry = 1.9f; p.rtheta = 0.7f; p.rw = 4.07f; p.rh = 1.74f; p.ow = 2.3f; p.oh = 1.1f;
    p.sd_x = sqrtf(0.2f); p.sd_y = sqrtf(0.1f); p.sd_theta = sqrtf(0.15f); p.sd_w = 0.0f; p.sd_h = 0.0f;
    uint64_t hits = 0, again = 0;
    float cp = 0.0f;
    rc = satmc_count_fused_host(ctx, &p, 1, 1000000, 42, 0, 0, &hits, 0);
    if (rc == SATMC_OK) rc = satmc_count_fused_host(ctx, &p, 1, 1000000, 42, 0, 0, &again, 0);
    if (rc == SATMC_OK) rc = satmc_collision_probability_host(ctx, &p, 1, 1000000, 42, &cp);
    if (rc != SATMC_OK) { fprintf(stderr, "error %d: %s\n", rc, satmc_last_error(ctx)); return 1; }
    printf("%s hits %llu of 1000000 (p = %.4f) launches %llu\n", satmc_version(), (unsigned long long)hits, cp,
           (unsigned long long)satmc_launch_count(ctx));
    satmc_destroy(ctx);
    return (hits == again && hits > 150000 && hits < 185000) ? 0 : 1;
}
