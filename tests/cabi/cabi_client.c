/* A plain C99 client of include/sdvpcm.h: what a maintainer's translation unit sees.  Built and run by tests/test_capi.py:
 * checks that the header is valid C, that the structure layouts are the documented ones and that the library refuses to
 * work without a GPU (no CPU fallback); with a GPU it goes on to the argument checks.  The program that DECODES through the ABI
 * on a GPU box is cabi_decode_client.c (tests/test_capi.py::test_c_client_decodes_golden_frames). */
#include <stdio.h>
#include <string.h>
#include <stddef.h>
#include "sdvpcm.h"

int main(void)
{
    sdv_handle *h = NULL;
    int rc;
    if(sizeof(sdv_line_rec)!=32 || sizeof(sdv_line_aux)!=16 || sizeof(sdv_block_rec)!=32 || sizeof(sdv_bin_config)!=16
       || sizeof(sdv_deint_config)!=16 || sizeof(sdv_stc007_geometry)!=16 || sizeof(sdv_pcm1_subline)!=8
       || sizeof(sdv_pcm16x0_subline)!=8 || sizeof(sdv_pcm1_frame_info)!=16 || sizeof(sdv_pcm1_stitch_config)!=8 || sizeof(sdv_pcm16x0_geometry)!=8
       || sizeof(sdv_stc007_stitch_config)!=16 || sizeof(sdv_stc007_frame_info)!=44 || sizeof(sdv_countdown)!=8 || offsetof(sdv_deint_config, countdown_in)!=7
       || offsetof(sdv_line_rec, flags)!=18 || offsetof(sdv_line_rec, data_start)!=24 || offsetof(sdv_line_rec, mark_stages)!=30)
    { printf("layout mismatch\n"); return 2; }
    if(sdv_version()!=100) { printf("version\n"); return 2; }
    rc = sdv_create(&h, 0);
    if(rc==SDV_ERR_CUDA) { printf("no-gpu: sdv_create refused (%s)\n", sdv_last_error(NULL)); return (h==NULL) ? 0 : 2; }
    if(rc!=SDV_OK) { printf("sdv_create: %d\n", rc); return 2; }
    /* argument checking needs no device buffers */
    {
        sdv_bin_config cfg; memset(&cfg, 0, sizeof(cfg));
        cfg.pcm_type = SDV_TYPE_STC007; cfg.mode = SDV_MODE_NORMAL; cfg.check_line_dup = 1;
        rc = sdv_bin_decode_frames(h, &cfg, NULL, 0, 576, 720, 720, NULL, NULL, NULL);
        if(rc!=SDV_OK) { printf("empty decode: %d %s\n", rc, sdv_last_error(h)); return 2; }
        rc = sdv_bin_decode_frames(h, &cfg, NULL, 1, 576, 100, 100, NULL, NULL, NULL);
        if(rc!=SDV_ERR_ARG) { printf("short line accepted: %d\n", rc); return 2; }
        cfg.pcm_type = 9;
        rc = sdv_bin_decode_frames(h, &cfg, NULL, 1, 576, 720, 720, NULL, NULL, NULL);
        if(rc!=SDV_ERR_ARG && rc!=SDV_ERR_UNSUPPORTED) { printf("bad type accepted: %d\n", rc); return 2; }
    }
    sdv_destroy(h);
    printf("gpu: ok\n");
    return 0;
}
