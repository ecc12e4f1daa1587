// pcm1_chain.cuh -- the inter-line chain of VideoToDigital::doBinarize (videotodigital.cpp:698-1815) for PCM-1.
//
// Differences from STC-007 (stc007_chain.cuh): every frame starts with prescanCoordinates() (videotodigital.cpp:148-345),
// four full coordinate searches on lines spread over the frame, run with a reset Binarizer -- so the per-frame presets
// (reference level, data coordinates) are a pure function of the frame's pixels and are computed for all frames at once
// by pcm1_prescan_kernel; lines with data are those with black/white levels found; the header line is the service line.
#pragma once
#include "pcm1_line.cuh"
#include "stc007_chain.cuh"

namespace sdv {

enum { P1_COORD_CHECK_LINES = 4, P1_COORD_CHECK_PARTS = P1_COORD_CHECK_LINES+2 };      // videotodigital.h:101-102

// Result of one prescan line / of the whole prescan of a frame (8 bytes).
struct P1Preset { Coord coords; u8 valid, ref, pad[2]; };

struct P1ChainCtx
{
    BinState bin;
    u8 field_state, line_dup, prescan_ref, pad0;
    u16 last_words[6];                      // words of the previous line with PCM in this field (last_pcm1_line)
    Coord last_valid[COORD_HISTORY_DEPTH];
    Coord long_valid[COORD_LONG_HISTORY];
    int n_last, n_long;
    Coord frame_avg;
    int n_fv, n_fi;
    int good_in_field, pcm_in_field;        // good_coords_in_field, pcm_lines_in_field
    unsigned long long lines_chain, lines_search;
    Coord frame_valid[SDV_MAX_H];
    Coord frame_invalid[SDV_MAX_H];
};

// PCM1Line::getSample / isNearSilence / isAlmostSilent (pcm1line.cpp:196-233,335-366).
SDV_HD i16 p1_sample(u16 w)
{
    if((w&P1_BIT_RANGE)==0) return (i16)(u16)(w<<4);
    bool pos = (w&0x0800)==0;
    w = (u16)(w&~P1_BIT_RANGE);
    w = (u16)(w<<2);
    if(!pos) w |= 0xC000;
    return (i16)w;
}
SDV_HD bool p1_words_almost_silent(const u16 *w)
{
    int cnt = 0;
    for(int i=0;i<6;i++) { i16 s = p1_sample(w[i]); if(!(s>=8)&&!(s<-8)) cnt++; }
    return cnt>=2;
}
// PCM1Line::getWordsDiffBitCount: the XOR is truncated to 8 bits (pcm1line.cpp:236-264).
SDV_HD int p1_words_diff8(const u16 *a, const u16 *b)
{
    int cnt = 0;
    for(int i=0;i<6;i++)
    {
        u32 d = (u32)((a[i]^b[i])&0xFF);
#if defined(__CUDA_ARCH__)
        cnt += __popc(d);
#else
        cnt += __builtin_popcount(d);
#endif
    }
    return cnt;
}
SDV_HD void p1_bin_set_good(BinState *b, const P1Line *l)
{
    if(p1_crc_ok_ign(l)) { b->def_ref = l->ref; bin_set_coords(b, l->coords); bin_set_bw(b, l->black, l->white); }
}

SDV_HD void p1_chain_reset(P1ChainCtx *x, int mode, int line_dup)
{
    bin_set_mode(&x->bin, mode);
    x->bin.def_coord = coord_none();
    bin_reset_good(&x->bin);
    x->field_state = FIELD_NEW; x->line_dup = (u8)(line_dup ? 1 : 0); x->prescan_ref = 128;
    for(int i=0;i<6;i++) x->last_words[i] = P1_BIT_RANGE;
    x->n_last = x->n_long = 0; x->n_fv = x->n_fi = 0;
    x->frame_avg = coord_none();
    x->good_in_field = x->pcm_in_field = 0;
    x->lines_chain = x->lines_search = 0;
}

// Which video row prescanCoordinates() reads as its [idx]-th line: frame_buf = [NEW_FILE] odd field, END_FIELD, even
// field, END_FIELD, END_FRAME (videotodigital.cpp:196-205).  Returns the frame row or -1 for a service line.
SDV_HD int p1_prescan_row(int H, bool first, int idx)
{
    const int hf = H/2;
    const int lines_cnt = H+3+(first ? 1 : 0);
    const int gap = lines_cnt/(P1_COORD_CHECK_PARTS-1);
    int e = (idx+1)*gap-(first ? 1 : 0);
    if(e<0) return -1;
    if(e<hf) return 2*e;
    if(e==hf) return -1;
    if(e<(2*hf+1)) return 2*(e-hf-1)+1;
    return -1;
}
SDV_HD bool p1_prescan_runs(int H, bool first, int mode) { return (mode!=SDV_MODE_DRAFT)&&((H+3+(first ? 1 : 0))>P1_COORD_CHECK_PARTS); }

// Median of the prescan results (videotodigital.cpp:303-330): element n/2 of the sorted coordinates and reference levels.
SDV_HD P1Preset p1_prescan_reduce(const P1Preset *r /*[4]*/)
{
    P1Preset out; out.valid = 0; out.ref = 0; out.coords = coord_none(); out.pad[0] = out.pad[1] = 0;
    Coord cl[P1_COORD_CHECK_LINES]; u8 rl[P1_COORD_CHECK_LINES]; int n = 0;
    for(int i=0;i<P1_COORD_CHECK_LINES;i++) if(r[i].valid) { cl[n] = r[i].coords; rl[n] = r[i].ref; n++; }
    if(n==0) return out;
    for(int i=1;i<n;i++) { u8 v = rl[i]; int k = i; while((k>0)&&(rl[k-1]>v)) { rl[k] = rl[k-1]; k--; } rl[k] = v; }
    for(int i=1;i<n;i++) { Coord v = cl[i]; int k = i; while((k>0)&&coord_less(v, 0, cl[k-1], 0)) { cl[k] = cl[k-1]; k--; } cl[k] = v; }
    out.valid = 1; out.coords = cl[n/2]; out.ref = rl[n/2];
    return out;
}

// Frame start (videotodigital.cpp:774-823).
SDV_HD void p1_chain_frame_start(P1ChainCtx *x, bool prescan_ran, P1Preset ps)
{
    x->frame_avg = coord_none();
    if(prescan_ran)
    {
        bin_reset_good(&x->bin);
        if(ps.valid) { x->frame_avg = ps.coords; x->prescan_ref = ps.ref; }
    }
    if(!coord_valid(x->frame_avg)) x->frame_avg = median_small(x->long_valid, x->n_long);
    else x->bin.def_ref = x->prescan_ref;
    if(coord_valid(x->frame_avg)) bin_set_coords2(&x->bin, x->frame_avg.start, x->frame_avg.stop);
    x->field_state = FIELD_NEW;
    x->good_in_field = x->pcm_in_field = 0;
}
SDV_HD void p1_chain_field_end(P1ChainCtx *x)
{
    x->field_state = FIELD_NEW;
    x->good_in_field = x->pcm_in_field = 0;
    for(int i=0;i<6;i++) x->last_words[i] = P1_BIT_RANGE;
}
SDV_HD void p1_chain_frame_end(P1ChainCtx *x, Coord med_valid, Coord med_invalid)
{
    x->frame_avg = med_valid;
    if(coord_valid(x->frame_avg))
    {
        if(x->n_long==COORD_LONG_HISTORY) { for(int i=1;i<COORD_LONG_HISTORY;i++) x->long_valid[i-1] = x->long_valid[i]; x->n_long--; }
        x->long_valid[x->n_long++] = x->frame_avg;
    }
    else
    {
        x->frame_avg = med_invalid;
        if(!coord_valid(x->frame_avg)) x->frame_avg = median_small(x->long_valid, x->n_long);
    }
    x->n_fv = x->n_fi = 0;
}
// setCoordinatesSearch() before each line (videotodigital.cpp:871-893).
SDV_HD bool p1_chain_coord_search(const P1ChainCtx *x)
{
    if((x->bin.mode==SDV_MODE_DRAFT)||(x->bin.mode==SDV_MODE_FAST)) return !((x->good_in_field>2)||(x->pcm_in_field>2));
    return true;
}

// What doBinarize does with one decoded line (videotodigital.cpp:1006-1657), thread 0 only.
SDV_HD void p1_chain_line(P1ChainCtx *x, P1Line *line)
{
    if(line->service!=0)
    {
        if((line->service==SDV_SRV_HEADER_LINE)&&(x->field_state==FIELD_NEW)) x->field_state = FIELD_SAFE;
        return;
    }
    bool has_data = line->bw_set!=0;
    bool has_pcm = p1_crc_ok(line)||has_data;
    if(has_pcm&&(x->field_state==FIELD_NEW)) x->field_state = FIELD_UNSAFE;
    if(p1_crc_ok(line))
    {
        x->good_in_field++;
        if(x->line_dup)
        {
            if(x->field_state==FIELD_UNSAFE)
            {
                p1_bin_set_good(&x->bin, line);
                if(FINE_FIRST_LINE_DUP) line->forced_bad = 1;
            }
            else
            {
                bool same = p1_words_diff8(line->words, x->last_words)<=(P1_BITS/32);
                if((!p1_words_almost_silent(line->words))&&same) line->forced_bad = 1;
            }
        }
        if(p1_crc_ok_ign(line))
        {
            if(x->n_last==COORD_HISTORY_DEPTH) { for(int i=1;i<COORD_HISTORY_DEPTH;i++) x->last_valid[i-1] = x->last_valid[i]; x->n_last--; }
            x->last_valid[x->n_last++] = line->coords;
            if(x->n_fv<SDV_MAX_H) x->frame_valid[x->n_fv++] = line->coords;
            if(x->n_last>(COORD_HISTORY_DEPTH/2))
            {
                Coord target = median_small(x->last_valid, x->n_last);
                if(!coord_valid(target)) target = x->frame_avg;
                if(coord_valid(target))
                {
                    i16 ds = (i16)(line->coords.start-target.start), de = (i16)(line->coords.stop-target.stop);
                    if(delta_warning(ds, de, (int)(u8)(p1_get_ppb(line)*3))) line->forced_bad = 1;
                }
            }
        }
        if(p1_crc_ok(line)) p1_bin_set_good(&x->bin, line);
        x->field_state = FIELD_INIT;
    }
    else
    {
        if(coord_valid(line->coords)) { if(x->n_fi<SDV_MAX_H) x->frame_invalid[x->n_fi++] = line->coords; }
        if(has_data)
        {
            Coord preset = median_small(x->last_valid, x->n_last);
            if(!coord_valid(preset)) preset = x->frame_avg;
            x->field_state = FIELD_INIT;
            bin_set_coords(&x->bin, preset);
            bin_set_bw(&x->bin, 0, 0);
        }
        else bin_set_bw(&x->bin, 0, 0);
    }
    if(has_pcm) { x->pcm_in_field++; for(int i=0;i<6;i++) x->last_words[i] = line->words[i]; }
}

SDV_HD void p1_export_line(const P1Line *l, sdv_line_rec *r, sdv_line_aux *a)
{
    sdv_line_rec t;
    for(int i=0;i<P1L_WORDS;i++) t.words[i] = l->words[i];
    t.words[7] = t.words[8] = 0;
    u16 f = 0;
    if(p1_crc_ok(l)) f |= SDV_LF_CRC_OK;
    if(p1_crc_ok_ign(l)) f |= SDV_LF_CRC_OK_IGN;
    if(l->forced_bad) f |= SDV_LF_FORCED_BAD;
    if(l->bw_set) f |= SDV_LF_BW_SET;
    if(l->coords_set) f |= SDV_LF_COORDS_SET;
    if(l->sweeped) f |= SDV_LF_REF_SWEEP;
    if(l->by_ext) f |= SDV_LF_BY_EXT;
    if(l->coord_sweeped) f |= SDV_LF_COORD_SWEEP;
    if(p1_words_almost_silent(l->words)) f |= SDV_LF_ALMOST_SILENT;
    t.flags = f;
    t.ref = l->ref; t.black = l->black; t.white = l->white; t.hyst = l->hyst;
    t.data_start = l->coords.start; t.data_stop = l->coords.stop;
    t.shift = l->shift; t.service_type = l->service;
    t.mark_stages = (u8)(l->picked_left|(l->picked_right<<4));
    t.reserved = 0;
#if defined(__CUDA_ARCH__)
    {
        uint4 v[2];
        memcpy(v, &t, sizeof(t));
        ((uint4 *)r)[0] = v[0]; ((uint4 *)r)[1] = v[1];
    }
#else
    *r = t;
#endif
    if(a)
    {
        sdv_line_aux u;
        u.ref_low = l->ref_low; u.ref_high = l->ref_high;
        u.marker_start_bg = u.marker_start_ed = u.marker_stop_ed = 0;
        u.word_crc_mask = u.word_valid_mask = 0;
        u.pad[0] = u.pad[1] = u.pad[2] = u.pad[3] = 0;
        *a = u;
    }
}

}   // namespace sdv
