/* A plain C99 host of the C ABI (what any FFI -- Julia ccall, ctypes, cgo -- sees): prints the struct layouts
 * so the language bindings can be checked against them, and exercises the argument validation that needs no GPU. */
#include <stddef.h>
#include <stdio.h>
#include <string.h>

#include "rrtmgp_b200.h"

#define OFF(T, f) printf("%s.%s %zu\n", #T, #f, offsetof(rrtmgp_b200_##T##_t, f))

int main(void) {
    printf("abi_version %d\n", rrtmgp_b200_abi_version());
    printf("sizeof.config %zu\n", sizeof(rrtmgp_b200_config_t));
    printf("sizeof.buffers %zu\n", sizeof(rrtmgp_b200_buffers_t));
    printf("sizeof.lut_info %zu\n", sizeof(rrtmgp_b200_lut_info_t));
    OFF(config, abi_version); OFF(config, nlay); OFF(config, isothermal_boundary_layer); OFF(config, col_offset);
    OFF(config, grav); OFF(config, avogad);
    OFF(buffers, layerdata); OFF(buffers, vmr); OFF(buffers, cld_frac); OFF(buffers, aero_mass); OFF(buffers, sfc_emis);
    OFF(buffers, metric_scaling); OFF(buffers, lw_flux_up); OFF(buffers, net_flux); OFF(buffers, clear_net_flux);
    OFF(buffers, sw_band_flux_net);
    OFF(lut_info, iband_550nm); OFF(lut_info, p_ref_min); OFF(lut_info, solar_src_tot);

    rrtmgp_b200_config_t c;
    memset(&c, 0, sizeof c);
    rrtmgp_b200_handle_t* h = (rrtmgp_b200_handle_t*)0x1;
    c.abi_version = RRTMGP_B200_ABI_VERSION + 1;           /* wrong ABI version: refused before anything else */
    printf("create.bad_abi %d\n", rrtmgp_b200_create(&c, &h));
    printf("create.bad_abi.handle_null %d\n", h == NULL);
    c.abi_version = RRTMGP_B200_ABI_VERSION;
    c.ncol = 8; c.nlay = 8; c.ngas = 8; c.ice_rgh = 2; c.n_gauss_angles = 3;   /* 3 angles with the two-stream solver */
    printf("create.angles_without_noscat %d\n", rrtmgp_b200_create(&c, &h));
    printf("create.null %d\n", rrtmgp_b200_create(NULL, NULL));
    printf("update.null %d\n", rrtmgp_b200_update_fluxes(NULL, 0, 0, NULL));
    printf("strerror.ok %s\n", rrtmgp_b200_strerror(RRTMGP_B200_OK));
    printf("strerror.bad_pack %s\n", rrtmgp_b200_strerror(RRTMGP_B200_ERR_BAD_LUT_PACK));
    return 0;
}
