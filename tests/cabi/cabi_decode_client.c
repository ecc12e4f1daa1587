/* A plain C99 program that DECODES through include/sdvpcm.h on a GPU box: what a maintainer's plugin does.  Run by
 * tests/test_capi.py::test_c_client_decodes_golden_frames (gpu): the test writes the luma of a tape and the expected records /
 * samples / flags (tests/golden fixtures of the reference, or the reference run live) as raw files; this program reads
 * the luma, calls the host-buffer entry point of the format and compares byte for byte.
 *
 *   cabi_decode_client <format: stc007|pcm1|pcm16x0> <n_frames> <H> <W> <luma.bin> <recs.bin|-> <samples.bin> <flags.bin>
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "sdvpcm.h"

static void *slurp(const char *path, size_t *n)
{
    FILE *f = fopen(path, "rb");
    void *p; long sz;
    if(!f) { printf("cannot open %s\n", path); exit(2); }
    fseek(f, 0, SEEK_END); sz = ftell(f); fseek(f, 0, SEEK_SET);
    p = malloc((size_t)sz+16);
    if(fread(p, 1, (size_t)sz, f)!=(size_t)sz) { printf("short read %s\n", path); exit(2); }
    fclose(f);
    *n = (size_t)sz;
    return p;
}

static int differs(const char *what, const void *a, const void *b, size_t n, size_t unit)
{
    const unsigned char *x = (const unsigned char *)a, *y = (const unsigned char *)b;
    size_t i;
    for(i=0;i<n;i++) if(x[i]!=y[i]) { printf("%s: first difference in element %lu\n", what, (unsigned long)(i/unit)); return 1; }
    return 0;
}

int main(int argc, char **argv)
{
    sdv_handle *h = NULL;
    sdv_bin_config bcfg;
    int rc, n_frames, H, W, bad = 0;
    size_t n_luma, n_recs = 0, n_smp, n_fl, rec_mult = 1;
    unsigned char *luma, *exp_recs = NULL, *exp_smp, *exp_fl;
    short *smp; unsigned char *fl; sdv_line_rec *recs;
    long n_blocks;
    if(argc!=9) { printf("usage\n"); return 2; }
    n_frames = atoi(argv[2]); H = atoi(argv[3]); W = atoi(argv[4]);
    luma = (unsigned char *)slurp(argv[5], &n_luma);
    if(strcmp(argv[6], "-")) exp_recs = (unsigned char *)slurp(argv[6], &n_recs);
    exp_smp = (unsigned char *)slurp(argv[7], &n_smp);
    exp_fl = (unsigned char *)slurp(argv[8], &n_fl);
    if(n_luma!=(size_t)n_frames*H*W) { printf("luma size\n"); return 2; }
    rc = sdv_create(&h, 0);
    if(rc!=SDV_OK) { printf("sdv_create: %d (this program needs a GPU: the library has no CPU path)\n", rc); return 3; }
    memset(&bcfg, 0, sizeof(bcfg));
    bcfg.mode = SDV_MODE_NORMAL; bcfg.check_line_dup = 1;
    if(!strcmp(argv[1], "stc007"))
    {
        sdv_deint_config dcfg; sdv_stc007_geometry geo;
        memset(&dcfg, 0, sizeof(dcfg)); memset(&geo, 0, sizeof(geo));
        bcfg.pcm_type = SDV_TYPE_STC007;
        dcfg.res_mode = SDV_RES_MODE_14BIT; dcfg.force_check = 1; dcfg.p_corr = 1; dcfg.q_corr = 1; dcfg.broken_mask_dur = 128;
        geo.lines_per_field = (H>500) ? 294 : 245; geo.lead_in = 80;
        n_blocks = sdv_stc007_block_count(&geo, n_frames);
        smp = (short *)malloc((size_t)n_blocks*12+16); fl = (unsigned char *)malloc((size_t)n_blocks*6+16);
        recs = (sdv_line_rec *)malloc((size_t)n_frames*H*sizeof(sdv_line_rec)+16);
        rc = sdv_stc007_decode_tape_host(h, &bcfg, &dcfg, &geo, luma, n_frames, H, W, smp, fl, recs);
        if(rc!=SDV_OK) { printf("decode: %d %s\n", rc, sdv_last_error(h)); return 2; }
        if(n_smp>(size_t)n_blocks*12) { printf("expected stream longer than the decode\n"); return 2; }
    }
    else if(!strcmp(argv[1], "pcm1"))
    {
        sdv_pcm1_stitch_config scfg;
        memset(&scfg, 0, sizeof(scfg));
        bcfg.pcm_type = SDV_TYPE_PCM1; scfg.file_start = 1;
        n_blocks = (long)n_frames*2*1470;
        smp = (short *)malloc((size_t)n_blocks*2+16); fl = (unsigned char *)malloc((size_t)n_blocks+16);
        recs = (sdv_line_rec *)malloc((size_t)n_frames*H*sizeof(sdv_line_rec)+16);
        rc = sdv_pcm1_decode_tape_host(h, &bcfg, &scfg, luma, n_frames, H, W, smp, fl, recs);
        if(rc!=SDV_OK) { printf("decode: %d %s\n", rc, sdv_last_error(h)); return 2; }
    }
    else
    {
        sdv_pcm16x0_config xcfg; sdv_pcm16x0_geometry geo;
        memset(&xcfg, 0, sizeof(xcfg)); memset(&geo, 0, sizeof(geo));
        bcfg.pcm_type = SDV_TYPE_PCM16X0; rec_mult = 3;
        xcfg.force_check = 1; xcfg.p_corr = 1;
        geo.top_padding_odd = 5; geo.top_padding_even = 5; geo.broken_mask_dur = 81;
        n_blocks = (long)n_frames*2*1470;
        smp = (short *)malloc((size_t)n_blocks*2+16); fl = (unsigned char *)malloc((size_t)n_blocks+16);
        recs = (sdv_line_rec *)malloc((size_t)n_frames*H*3*sizeof(sdv_line_rec)+16);
        rc = sdv_pcm16x0_decode_tape_host(h, &bcfg, &xcfg, &geo, luma, n_frames, H, W, smp, fl, recs);
        if(rc!=SDV_OK) { printf("decode: %d %s\n", rc, sdv_last_error(h)); return 2; }
    }
    if(exp_recs)
    {
        if(n_recs!=(size_t)n_frames*H*rec_mult*sizeof(sdv_line_rec)) { printf("record file size\n"); return 2; }
        bad |= differs("line records", recs, exp_recs, n_recs, sizeof(sdv_line_rec));
    }
    bad |= differs("samples", smp, exp_smp, n_smp, 2);
    bad |= differs("sample flags", fl, exp_fl, n_fl, 1);
    sdv_destroy(h);
    if(bad) return 1;
    printf("decoded %d frames of %s through the C ABI: records, samples and flags equal the expected files\n", n_frames, argv[1]);
    return 0;
}
