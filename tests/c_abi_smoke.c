/* Plain-C consumer of include/p2p.h: proves the boundary is a C ABI (no C++ / Python / torch types).
 *   gcc -std=c99 -Wall -Iinclude tests/c_abi_smoke.c -o c_abi_smoke -L<dir> -lp2p_b200 -Wl,-rpath,<dir>
 *   ./c_abi_smoke Wp Hp W H fov out.bin      (writes (n_yaw + 1) * n_pitch * H * W * 3 bytes: three integer-roll yaws
 *                                             through p2p_project_views, then yaw 33.3 - not an integer column roll -
 *                                             through p2p_project_views_table)
 * The panorama is a deterministic LCG pattern the Python test regenerates. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "p2p.h"

#define CHECK(call)                                                                 \
    do {                                                                            \
        int rc_ = (call);                                                           \
        if (rc_ != P2P_OK) {                                                        \
            fprintf(stderr, "%s -> %d (%s): %s\n", #call, rc_, p2p_status_string(rc_), \
                    ctx ? p2p_last_error(ctx) : "");                                \
            return 2;                                                               \
        }                                                                           \
    } while (0)

int main(int argc, char **argv) {
    p2p_ctx *ctx = NULL;
    if (argc != 7) {
        fprintf(stderr, "usage: %s Wp Hp W H fov out.bin\n", argv[0]);
        return 1;
    }
    const int Wp = atoi(argv[1]), Hp = atoi(argv[2]), W = atoi(argv[3]), H = atoi(argv[4]), fov = atoi(argv[5]);
    const double yaws[3] = {0.0, 90.0, 270.0};
    const double pitches[2] = {60.0, 120.0};
    const int n_yaw = 3, n_pitch = 2;
    size_t i, n = (size_t)Wp * Hp * 3;
    unsigned char *pano = (unsigned char *)malloc(n);
    unsigned char *out = (unsigned char *)malloc((size_t)n_yaw * n_pitch * W * H * 3);
    int32_t *ix = (int32_t *)malloc(sizeof(int32_t) * (size_t)Wp), *fx = (int32_t *)malloc(sizeof(int32_t) * (size_t)Wp);
    int32_t shifts[3];
    p2p_pitch_consts pc[2];
    unsigned int s = 12345u;
    int k;
    FILE *f;
    if (!pano || !out || !ix || !fx) return 1;
    for (i = 0; i < n; ++i) {
        s = s * 1664525u + 1013904223u;
        pano[i] = (unsigned char)(s >> 24);
    }
    if (p2p_abi_version() != P2P_ABI_VERSION) return 3;
    CHECK(p2p_create(0, 1, &ctx));
    for (k = 0; k < n_yaw; ++k) {
        CHECK(p2p_yaw_table(Wp, yaws[k], ix, fx, &shifts[k]));
        if (shifts[k] < 0) return 4; /* these yaws are integer rolls for Wp % 4 == 0 */
    }
    for (k = 0; k < n_pitch; ++k) CHECK(p2p_pitch_constants((double)fov, pitches[k], W, &pc[k]));
    CHECK(p2p_upload_pano(ctx, 0, pano, Wp, Hp, (size_t)Wp * 3));
    CHECK(p2p_project_views(ctx, 0, n_yaw, shifts, n_pitch, pc, W, H, out, 0));
    CHECK(p2p_sync(ctx, 0));
    f = fopen(argv[6], "wb");
    if (!f) return 5;
    fwrite(out, 1, (size_t)n_yaw * n_pitch * W * H * 3, f);
    {   /* a fractional yaw: both remap passes of the reference in one call, from the same slot */
        int32_t frac_shift;
        const int32_t *ixs[1], *fxs[1];
        CHECK(p2p_yaw_table(Wp, 33.3, ix, fx, &frac_shift));
        if (frac_shift >= 0) return 6;
        ixs[0] = ix;
        fxs[0] = fx;
        CHECK(p2p_project_views_table(ctx, 0, 1, ixs, fxs, n_pitch, pc, W, H, out, 0));
        CHECK(p2p_sync(ctx, 0));
        fwrite(out, 1, (size_t)n_pitch * W * H * 3, f);
    }
    fclose(f);
    p2p_destroy(ctx);
    free(pano); free(out); free(ix); free(fx);
    return 0;
}
