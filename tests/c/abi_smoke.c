/* Plain-C consumer of include/b200_sixdof.h: proves the boundary is a C ABI (no C++/torch types)
 * and doubles as the cgo/FFI-style usage example.  Build: gcc -std=c99 tests/c/abi_smoke.c -Iinclude
 * -Lelodin_b200 -lb200_sixdof.  Without a GPU it checks the loud-failure contract; with one it
 * integrates a free body and checks x = v*t. */
#include <math.h>
#include <stdio.h>
#include <string.h>

#include "b200_sixdof.h"

int main(void)
{
    b200_sixdof_desc d;
    b200_sixdof *h = 0;
    memset(&d, 0, sizeof d);
    d.abi_version = B200_SIXDOF_ABI_VERSION;
    d.integrator = B200_INTEGRATOR_RK4;
    d.math_mode = B200_MATH_EXACT;
    d.n_entities = 1;
    d.n_worlds = 1;
    d.sim_time_step = 1.0 / 64.0;
    d.time_step = NAN;
    d.device = -1;
    if (b200_component_id("world_pos") != B200_ID_WORLD_POS) { printf("component id mismatch\n"); return 1; }
    {   /* host-only entry: the EGM08 term stream of a degree-1 field (3 terms x 8 f64; a_bar[0][0] = 1 leads it) */
        const double c[4] = {1.0, 0.0, 0.25, 0.5}, s[4] = {0.0, 0.0, 0.0, 0.125};
        double st[24];
        if (b200_egm08_stream_len(1) != 24 || b200_egm08_stream(1, c, s, st, 24) != B200_OK || st[0] != 1.0 || st[4] != 1.0 ||
            st[8 + 4] != 0.25 || st[16 + 4] != 0.5 || st[16 + 5] != 0.125 ||
            b200_egm08_stream(1, c, s, st, 23) != B200_ERR_VALUE_SIZE_MISMATCH) { printf("egm08 stream: %s\n", b200_last_error()); return 1; }
    }
    if (b200_device_count() <= 0) {
        int rc = b200_sixdof_create(&d, &h);
        if (rc != B200_ERR_NO_DEVICE || h != 0) { printf("expected B200_ERR_NO_DEVICE, got %d\n", rc); return 1; }
        printf("no GPU: create failed loudly as designed: %s\n", b200_last_error());
        return 0;
    }
    if (b200_sixdof_create(&d, &h) != B200_OK) { printf("create: %s\n", b200_last_error()); return 1; }
    {
        double pos[7] = {0, 0, 0, 1, 0, 0, 0}, vel[6] = {0, 0, 0, 2.0, 0, 0}, ine[7] = {1, 1, 1, 0, 0, 0, 1}, out[7];
        if (b200_sixdof_upload(h, B200_ID_WORLD_POS, pos, sizeof pos) || b200_sixdof_upload(h, B200_ID_WORLD_VEL, vel, sizeof vel) ||
            b200_sixdof_upload(h, B200_ID_INERTIA, ine, sizeof ine) || b200_sixdof_step(h, 64) ||
            b200_sixdof_download(h, B200_ID_WORLD_POS, out, sizeof out)) { printf("error: %s\n", b200_last_error()); return 1; }
        if (out[4] != 2.0 || out[3] != 1.0 || b200_sixdof_tick_count(h) != 64) { printf("wrong result %g\n", out[4]); return 1; }
        if (b200_sixdof_upload(h, B200_ID_WORLD_POS, pos, 8) != B200_ERR_VALUE_SIZE_MISMATCH) return 1;
        printf("C ABI ok: 64 ticks, x = %g\n", out[4]);
    }
    b200_sixdof_destroy(h);
    return 0;
}
