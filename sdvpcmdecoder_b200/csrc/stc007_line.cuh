// stc007_line.cuh -- STC-007 line decode (the Binarizer operator) as cooperative integer code.
//
// B200-native restructuring of Binarizer::processLine (binarizer.cpp:443-1724): one thread block decodes one video
// line held in shared memory.  The reference-level sweep (binarizer.cpp:3551-4120) runs one reference level per
// thread, the 24 marker trials of findSTC007Coordinates (binarizer.cpp:6047-6113) one trial per thread, the
// hysteresis x pixel-shift candidates of readPCMdata (binarizer.cpp:7695-8055) one candidate per thread; the
// reference's selection rules are then replayed by thread 0 over the per-thread results so that every field of the
// output line matches the sequential reference.  With Cta{0,1} the same code is a sequential program (tests/hostemu).
#pragma once
#include "sdv_common.cuh"

namespace sdv {

// ------------------------------------------------------------------------------------------------ line object
// STC007Line + PCMLine payload (pcmline.h:132-160, stc007line.h:154-166) without the pixel coordinate table
// (bit positions are recomputed from [ppb] where needed).
struct Line
{
    u16 words[9];
    u16 calc_crc;
    Coord coords;
    u8 black, white, ref_low, ref, ref_high, hyst, shift, service;
    u8 mst, med;                        // mark_st_stage, mark_ed_stage
    u16 m_bg, m_ed, m_stop;             // marker_start_bg_coord, marker_start_ed_coord, marker_stop_ed_coord
    u8 sweeped, by_ext, bw_set, coords_set, forced_bad, wflags;   // wflags: word_crc[]/word_valid[] (always set together)
    u8 m2;                              // STC007Line::m2_format (set by VideoToDigital after the decode, M2 tapes only)
    Ppb ppb;
};

SDV_HD void line_set_invalid_crc(Line *l) { l->words[8] = (u16)~l->calc_crc; }
SDV_HD bool line_crc_ok_ign(const Line *l) { return l->calc_crc==l->words[8]; }
SDV_HD bool line_crc_ok(const Line *l) { return (!l->forced_bad)&&line_crc_ok_ign(l); }
SDV_HD bool line_has_start(const Line *l) { return l->mst==MARK_ST_BOT_2; }
SDV_HD bool line_has_stop(const Line *l) { return l->med==MARK_ED_LEN_OK; }
SDV_HD bool line_has_markers(const Line *l) { return line_has_start(l)&&line_has_stop(l); }
SDV_HD int line_get_ppb(const Line *l) { return (int)((l->ppb.psm/INT_CALC_MULT)&0xFF); }

SDV_HD void line_clear(Line *l)
{   // STC007Line::clear (stc007line.cpp:69-98)
    for(int i=0;i<9;i++) l->words[i] = 0;
    l->coords = coord_none();
    l->black = l->white = l->ref_low = l->ref = l->ref_high = l->hyst = l->shift = l->service = 0;
    l->mst = l->med = 0; l->m_bg = l->m_ed = l->m_stop = 0;
    l->sweeped = l->by_ext = l->bw_set = l->coords_set = l->forced_bad = l->wflags = 0;
    l->m2 = 0;
    l->ppb.psm = INT_CALC_MULT; l->ppb.half = INT_CALC_MULT/2; l->ppb.ofs = 0;
    l->calc_crc = 0xA96A;
    line_set_invalid_crc(l);
}
// PCMLine::clear through the base pointer (pcmline.cpp:94-112, used at binarizer.cpp:3615): words, markers survive.
SDV_HD void line_base_clear(Line *l)
{
    l->black = l->white = l->ref_low = l->ref = l->ref_high = 0;
    l->coords = coord_none();
    l->hyst = l->shift = 0;
    l->sweeped = l->by_ext = 0;
    l->calc_crc = 0;
    l->bw_set = l->coords_set = l->forced_bad = 0;
    l->service = 0;
    l->ppb.psm = INT_CALC_MULT; l->ppb.half = INT_CALC_MULT/2; l->ppb.ofs = 0;
}
SDV_HD bool line_coord_set(Line *l, i32 s, i32 e) { if(e>s) { l->coords.start = (i16)s; l->coords.stop = (i16)e; return true; } return false; }

SDV_HD void line_set_serv_ctrl_blk(Line *l)
{   // STC007Line::setServCtrlBlk (stc007line.cpp:101-131)
    u16 w4 = l->words[4], w5 = l->words[5], w6 = l->words[6], w7 = l->words[7];
    line_clear(l);
    l->words[4] = w4; l->words[5] = w5; l->words[6] = w6; l->words[7] = w7;
    l->calc_crc = crc_stc007(l->words);
    l->words[8] = l->calc_crc;
    l->service = SDV_SRV_CTRL_BLOCK;
}

// ------------------------------------------------------------------------------------------------ binarizer presets
// The part of the Binarizer object that survives between lines (binarizer.h:277-292).
struct BinState
{
    u8 def_black, def_white, def_ref, mode;
    Coord def_coord;
    u8 max_hyst, max_shift;         // in_max_hysteresis_depth / in_max_shift_stages (set by the mode)
};
SDV_HD void bin_set_mode(BinState *b, int mode)
{   // binarizer.cpp:120-152
    if(mode==SDV_MODE_DRAFT) { b->mode = SDV_MODE_DRAFT; b->max_hyst = HYST_DEPTH_SAFE; b->max_shift = SHIFT_MIN; }
    else if(mode==SDV_MODE_FAST) { b->mode = SDV_MODE_FAST; b->max_hyst = 7; b->max_shift = SHIFT_SAFE; }
    else if(mode==SDV_MODE_INSANE) { b->mode = SDV_MODE_INSANE; b->max_hyst = HYST_DEPTH_MAX; b->max_shift = SHIFT_MAX; }
    else { b->mode = SDV_MODE_NORMAL; b->max_hyst = HYST_DEPTH_SAFE; b->max_shift = SHIFT_SAFE; }
}
SDV_HD void bin_set_bw(BinState *b, u8 bl, u8 wh)
{
    if((bl<wh)&&(bl<MAX_BLACK_LVL)&&(wh>MIN_WHITE_LVL)&&(wh!=0)) { b->def_black = bl; b->def_white = wh; }
    else b->def_black = b->def_white = 0;
}
SDV_HD void bin_set_coords(BinState *b, Coord c) { if(coord_valid(c)) b->def_coord = c; else b->def_coord = coord_none(); }
SDV_HD void bin_set_coords2(BinState *b, i16 s, i16 e)
{
    Coord t = coord_none();
    if((s<e)&&(e!=0)&&(s!=NO_COORD_LEFT)&&(e!=NO_COORD_RIGHT)) { t.start = s; t.stop = e; }
    bin_set_coords(b, t);
}
SDV_HD void bin_reset_good(BinState *b) { b->def_ref = 0; bin_set_coords2(b, 0, 0); bin_set_bw(b, 0, 0); }
SDV_HD void bin_set_good(BinState *b, const Line *l)
{
    if(line_crc_ok_ign(l)) { b->def_ref = l->ref; bin_set_coords(b, l->coords); bin_set_bw(b, l->black, l->white); }
}
SDV_HD bool bin_ref_preset(const BinState *b) { return b->def_ref>=MIN_REF_LVL; }
SDV_HD bool bin_bw_preset(const BinState *b)
{
    if((b->def_white>MIN_WHITE_LVL)&&(b->def_black<MAX_BLACK_LVL))
    {
        if(bin_ref_preset(b)) { if((b->def_ref<=b->def_black)||(b->def_ref>=b->def_white)) return false; }
        return true;
    }
    return false;
}
// True when the next line will take the preset-only path (STG_INPUT_ALL with black/white preset, binarizer.cpp:774-931).
SDV_HD bool bin_fast_ready(const BinState *b) { return bin_ref_preset(b)&&coord_valid(b->def_coord)&&bin_bw_preset(b); }

// ------------------------------------------------------------------------------------------------ marker search
// One pass of Binarizer::searchSTC007Markers (binarizer.cpp:5275-5595) as a pure function of (pixels, ref, hyst).
struct MarkRes
{
    u8 mst, med;            // stages; [med] is only meaningful when the START marker was found
    u16 st1s, st1e, st3e;   // START bit 1 begin/end, START bit 3 end
    u16 ed_s, ed_e;         // STOP marker begin/end
};
SDV_HD bool mark_has_both(const MarkRes &r) { return (r.mst==MARK_ST_BOT_2)&&(r.med==MARK_ED_LEN_OK); }

SDV_HDN MarkRes search_markers(const u8 *px, const Geom &g, u8 ref, u8 hyst_lvl)
{
    MarkRes r; r.st1s = r.st1e = r.st3e = r.ed_s = r.ed_e = 0; r.med = MARK_ED_START;
    u8 stage = MARK_ST_START, pv, bin_low, bin_high;
    u32 pixel, pixel_limit, st3s = 0;
    u32 ppb = g.est_ppb;
    bin_low = get_low_level(ref, hyst_lvl);
    if(bin_low<MIN_REF_LVL) bin_low = MIN_REF_LVL;
    bin_high = ref;
    pixel_limit = (u16)(g.mark_start_max+ppb*5);
    if(pixel_limit>(u32)g.W) pixel_limit = (u32)g.W;
    pixel = 0;
    while(pixel<pixel_limit)
    {
        pv = px[pixel];
        if(stage==MARK_ST_START)
        {
            if(pixel>g.mark_start_max) break;
            if(pv>=bin_low) { r.st1s = (u16)pixel; stage = MARK_ST_TOP_1; }
        }
        else if(stage==MARK_ST_TOP_1)
        {
            if(pv<bin_low) { r.st1e = (u16)pixel; stage = MARK_ST_BOT_1; }
        }
        else if(stage==MARK_ST_BOT_1)
        {
            if(pv>=bin_high)
            {
                st3s = pixel;
                i32 len = (i32)st3s-(i32)r.st1e;
                if((len>(i32)(ppb*2))||(len<(i32)(ppb/2))) stage = MARK_ST_START;
                else stage = MARK_ST_TOP_2;
            }
        }
        else if(stage==MARK_ST_TOP_2)
        {
            if(pv<bin_high)
            {
                r.st3e = (u16)pixel;
                i32 len = (i32)r.st3e-(i32)st3s;
                if((len>(i32)(ppb*2))||(len<(i32)(ppb/2))) stage = MARK_ST_START;
                else { stage = MARK_ST_BOT_2; break; }
            }
        }
        pixel++;
    }
    r.mst = stage;
    if(stage==MARK_ST_BOT_2)
    {
        stage = MARK_ED_START;
        bin_low = ref;
        if(g.mark_end_min>(ppb*6)) pixel_limit = (u16)(g.mark_end_min-ppb*6);
        else pixel_limit = 0;
        pixel = g.scan_end;
        while(pixel>pixel_limit)
        {
            pv = px[pixel];
            if(stage==MARK_ED_START)
            {
                if(pixel<g.mark_end_min) break;
                if(pv>=bin_low) { r.ed_e = (u16)(pixel+1); stage = MARK_ED_TOP; }
            }
            else if(stage==MARK_ED_TOP)
            {
                if(pv<bin_high)
                {
                    r.ed_s = (u16)(pixel+1);
                    i32 len = (i32)r.ed_e-(i32)r.ed_s;
                    if((len>=(i32)(ppb*2))&&(len<=(i32)(ppb*5))) { stage = MARK_ED_LEN_OK; break; }
                    else stage = MARK_ED_START;
                }
            }
            pixel--;
        }
        r.med = stage;
    }
    return r;
}

// Effects of a searchSTC007Markers() call on the line object.
SDV_HD void apply_markers(Line *l, const MarkRes &r)
{
    l->mst = r.mst;
    l->m_bg = r.st1s;
    l->m_ed = r.st3e;
    if(r.mst==MARK_ST_BOT_2) l->med = r.med;
    line_coord_set(l, r.st1e, r.ed_s);
    l->m_stop = r.ed_e;
    l->coords_set = line_has_markers(l);
}

// Winner of the 24 hysteresis trials: minimum under CoordinatePair::operator< (what std::sort puts first), else 0.
SDV_HD int pick_marker_trial(const MarkRes *trials)
{
    int best = -1; Coord bc = coord_none();
    for(int h=0;h<MARK_TRIALS;h++)
    {
        if(mark_has_both(trials[h]))
        {
            Coord c; c.start = (i16)trials[h].st1e; c.stop = (i16)trials[h].ed_s;
            if((best<0)||coord_less(c, h, bc, best)) { bc = c; best = h; }
        }
    }
    return (best<0) ? 0 : best;
}

// findSTC007Coordinates, one thread doing all trials (used per reference level inside the sweep).
SDV_HDN void find_coordinates_seq(const u8 *px, const Geom &g, Line *l)
{
    int best = -1; Coord bc = coord_none(); MarkRes br;
    MarkRes first = search_markers(px, g, l->ref, 0);
    br = first;
    if(mark_has_both(first)) { best = 0; bc.start = (i16)first.st1e; bc.stop = (i16)first.ed_s; }
    for(int h=1;h<MARK_TRIALS;h++)
    {
        MarkRes r = search_markers(px, g, l->ref, (u8)h);
        if(mark_has_both(r))
        {
            Coord c; c.start = (i16)r.st1e; c.stop = (i16)r.ed_s;
            if((best<0)||coord_less(c, h, bc, best)) { bc = c; best = h; br = r; }
        }
    }
    // The reference repeats the search with the winning hysteresis on the output line (binarizer.cpp:6107-6112):
    // same pixels, same thresholds -> same result as the stored trial.
    apply_markers(l, br);
}

// A marker trial reduced to what the sweep needs: both markers found, data start, data stop.
enum { SWEEP_MAX_LEVELS = 254 };         // reference levels a sweep can visit: between black + 1 and white - 1
SDV_HD u32 pack_trial(const MarkRes &r) { return (mark_has_both(r) ? 0x80000000u : 0u)|((u32)(r.st1e&0x7FFF)<<15)|(u32)(r.ed_s&0x7FFF); }
// findSTC007Coordinates from the 24 packed trials of one reference level; true when markers were found.
SDV_HD bool pick_packed_trial(const u32 *tr, Coord *out)
{
    int best = -1; Coord bc = coord_none();
    for(int h=0;h<MARK_TRIALS;h++)
    {
        const u32 t = tr[h];
        if(t&0x80000000u)
        {
            Coord c; c.start = (i16)((t>>15)&0x7FFF); c.stop = (i16)(t&0x7FFF);
            if((best<0)||coord_less(c, h, bc, best)) { bc = c; best = h; }
        }
    }
    if(best<0) return false;
    *out = bc;
    return true;
}

// findSTC007Coordinates with one trial per thread.  [trials] is shared scratch of MARK_TRIALS entries; [l] is shared.
SDV_HD void find_coordinates_cta(const Cta &c, const u8 *px, const Geom &g, Line *l, MarkRes *trials)
{
    c.sync();
    u8 ref = l->ref;
    for(int h=c.tid;h<MARK_TRIALS;h+=c.n) trials[h] = search_markers(px, g, ref, (u8)h);
    c.sync();
    if(c.tid==0) apply_markers(l, trials[pick_marker_trial(trials)]);
    c.sync();
}

// ------------------------------------------------------------------------------------------------ bit extraction
// Binarizer::fillSTC007 (binarizer.cpp:7322-7445): 128 bit cells sampled with the reference's level hysteresis.
SDV_HDN void fill_stc007(const u8 *px, int pixel_stop, Ppb ppb, int shift_stage, u8 low_ref, u8 high_ref, u16 *words /*[9]*/)
{
    bool prev_high = false;
    int sh = pix_shift(shift_stage);
    u32 acc = 0;
    int bit = 0;
    for(int w=0;w<9;w++)
    {
        int nb = (w<8) ? 14 : 16;
        acc = 0;
        for(int k=0;k<nb;k++, bit++)
        {
            u8 pv = px[pixel_of_bit(ppb, bit, sh, pixel_stop)];
            bool one;
            if(!prev_high) { one = pv>low_ref; if(one) prev_high = true; }
            else { one = pv>=high_ref; if(!one) prev_high = false; }
            acc = (acc<<1)|(one ? 1u : 0u);
        }
        words[w] = (u16)acc;
    }
}

// Result of one (hysteresis, shift) candidate of Binarizer::fillDataWords (binarizer.cpp:7560-7691).
struct Cand { u16 words[9]; u16 calc_crc; u8 ok; u8 low, high; u8 pad; };

SDV_HD void eval_cand(const u8 *px, int pixel_stop, const Line *l, int hyst, int shift, Cand *cd)
{
    cd->low = get_low_level(l->ref, (u8)hyst);
    cd->high = get_high_level(l->ref, (u8)hyst);
    cd->ok = 0;
    if(cd->low<=l->black) return;
    if(cd->high>=l->white) return;
    fill_stc007(px, pixel_stop, l->ppb, shift, cd->low, cd->high, cd->words);
    cd->calc_crc = crc_stc007(cd->words);
    cd->ok = 1;
}
// Effects of fillDataWords() on the line, given the evaluated candidate.
SDV_HD bool apply_cand(Line *l, int hyst, int shift, const Cand *cd)
{
    l->ref_low = cd->low; l->ref_high = cd->high;
    if(!cd->ok) { line_set_invalid_crc(l); return false; }
    l->hyst = (u8)hyst; l->shift = (u8)shift;
    for(int i=0;i<9;i++) l->words[i] = cd->words[i];
    l->wflags = 0;
    l->calc_crc = cd->calc_crc;
    return true;
}

// Binarizer::readPCMdata (binarizer.cpp:7695-8055): lexicographic (hysteresis, shift) search, first valid CRC wins,
// then the final fill with the winner (or with (0,0)).  CandFn(h, s) returns the evaluated candidate.
template<class CandFn>
SDV_HD void read_pcm_core(Line *l, int hlim, int slim, CandFn cand)
{
    int win_h = 0, win_s = 0;
    if(!l->sweeped)
    {
        bool found = false;
        for(int h=0;(h<=hlim)&&(!found);h++)
        {
            bool invalid_hyst = false;
            for(int s=0;s<=slim;s++)
            {
                const Cand *cd = cand(h, s);
                if(!apply_cand(l, h, s, cd)) { invalid_hyst = true; break; }
                if(line_crc_ok(l)) { found = true; win_h = h; win_s = s; break; }
            }
            if(invalid_hyst) break;
        }
    }
    else { win_h = hlim; win_s = slim; }
    apply_cand(l, win_h, win_s, cand(win_h, win_s));
}

struct SeqCandFn
{
    const u8 *px; int pixel_stop; const Line *l; Cand *tmp; int *last;      // *last = h*8+s of the candidate held in *tmp
    SDV_HD const Cand *operator()(int h, int s) const
    {
        if(*last!=(h*8+s)) { eval_cand(px, pixel_stop, l, h, s, tmp); *last = h*8+s; }     // same inputs -> same result: do not sample the line twice
        return tmp;
    }
};
struct TabCandFn
{
    const Cand *tab; int slim;
    SDV_HD const Cand *operator()(int h, int s) const { return &tab[h*(slim+1)+s]; }
};

SDV_HDN void read_pcm_seq(const u8 *px, const Geom &g, Line *l, int hlim, int slim)
{
    Cand tmp; int last = -1;
    if(hlim>HYST_DEPTH_MAX) hlim = HYST_DEPTH_MAX;
    if(slim>SHIFT_MAX) slim = SHIFT_MAX;
    l->ppb = make_ppb(l->coords);
    SeqCandFn f; f.px = px; f.pixel_stop = g.W-1; f.l = l; f.tmp = &tmp; f.last = &last;
    read_pcm_core(l, hlim, slim, f);
}

// readPCMdata with one candidate per thread ([tab] is shared scratch of MAX_CAND entries, [l] is shared).
SDV_HD void read_pcm_cta(const Cta &c, const u8 *px, const Geom &g, Line *l, int hlim, int slim, Cand *tab)
{
    if(hlim>HYST_DEPTH_MAX) hlim = HYST_DEPTH_MAX;
    if(slim>SHIFT_MAX) slim = SHIFT_MAX;
    c.sync();
    if(c.tid==0) l->ppb = make_ppb(l->coords);
    c.sync();
    bool sweeped = l->sweeped!=0;
    int ncand = (hlim+1)*(slim+1);
    if(sweeped)
    {
        if(c.tid==0) eval_cand(px, g.W-1, l, hlim, slim, &tab[hlim*(slim+1)+slim]);
    }
    else
    {
        for(int i=c.tid;i<ncand;i+=c.n) eval_cand(px, g.W-1, l, i/(slim+1), i%(slim+1), &tab[i]);
    }
    c.sync();
    if(c.tid==0) { TabCandFn f; f.tab = tab; f.slim = slim; read_pcm_core(l, hlim, slim, f); }
    c.sync();
}

// ------------------------------------------------------------------------------------------------ CRC statistics
struct CrcH { u8 result; u8 hyst, shift; u8 pad; u16 crc; i16 start, stop; };

SDV_HD void reset_crc_stats(CrcH *a, int n) { for(int i=0;i<n;i++) { a[i].result = 0; a[i].start = a[i].stop = 0; a[i].crc = 0; a[i].hyst = a[i].shift = 0x0f; } }
SDV_HD void update_crc_stats(CrcH *a, const CrcH &in, u8 *cnt)
{   // binarizer.cpp:1771-1830
    bool found = false;
    if(*cnt>=MAX_COLL_CRCS) *cnt = MAX_COLL_CRCS-1;
    for(u8 i=1;i<=*cnt;i++) if(a[i].crc==in.crc) { a[i].result++; found = true; break; }
    if(!found)
    {
        (*cnt)++;
        if(*cnt<MAX_COLL_CRCS) { a[*cnt].crc = in.crc; a[*cnt].hyst = in.hyst; a[*cnt].shift = in.shift; a[*cnt].result++; }
    }
}
// k calls of update_crc_stats with the same entry: the first inserts or finds it, the others find it (or, with the table full,
// all of them leave it untouched); the count wraps like k increments.
SDV_HD void update_crc_stats_n(CrcH *a, const CrcH &in, u8 *cnt, u8 k)
{
    if(k==0) return;
    bool found = false;
    if(*cnt>=MAX_COLL_CRCS) *cnt = MAX_COLL_CRCS-1;
    for(u8 i=1;i<=*cnt;i++) if(a[i].crc==in.crc) { a[i].result = (u8)(a[i].result+k); found = true; break; }
    if(!found)
    {
        (*cnt)++;
        if(*cnt<MAX_COLL_CRCS) { a[*cnt].crc = in.crc; a[*cnt].hyst = in.hyst; a[*cnt].shift = in.shift; a[*cnt].result = (u8)(a[*cnt].result+k); }
    }
}
SDV_HD void find_most_frequent_crc(CrcH *a, u8 *cnt, bool skip_equal)
{   // binarizer.cpp:1833-1900
    a[0].result = 0; a[0].start = 0; a[0].stop = 0; a[0].hyst = 0; a[0].shift = 0;
    if(*cnt>=MAX_COLL_CRCS) *cnt = MAX_COLL_CRCS-1;
    for(u8 i=1;i<=*cnt;i++)
        if(a[i].result>a[0].result) { a[0].result = a[i].result; a[0].crc = a[i].crc; a[0].hyst = a[i].hyst; a[0].shift = a[i].shift; a[0].start = i; }
    if(skip_equal)
        for(u8 i=1;i<=*cnt;i++)
            if(a[0].start!=i) if(a[0].result<=(2*a[i].result)) { a[0].result = 0; a[0].hyst = 0; a[0].shift = 0; break; }
    if(a[0].result==0) *cnt = 0;
}
SDV_HD void invalidate_non_frequent(CrcH *a, u8 lo, u8 hi, u8 cnt, u16 target)
{   // binarizer.cpp:1903-1950
    u8 idx = hi;
    while(idx>=lo)
    {
        if(a[idx].result==REF_CRC_OK) if((cnt==0)||(a[idx].crc!=target)) a[idx].result = REF_CRC_COLL;
        if(idx==lo) break;
        idx--;
    }
}
// Binarizer::pickLevelByCRCStats (binarizer.cpp:1953-2131)
SDV_HDN u8 pick_level_by_stats(const CrcH *c, u8 *res, u8 low_lvl, u8 high_lvl, u8 target, u8 max_hyst, u8 max_shift)
{
    bool good = false, range_lock = false, second_lock = false;
    u8 idx, low_depth = 0xFF, low_shift = 0xFF, low_ref = 0, high_ref = 0, tst_low = 0, tst_high = 0, picked;
    idx = high_lvl;
    while(idx>=low_lvl)
    {
        if((c[idx].result==target)&&(c[idx].hyst<=max_hyst)&&(c[idx].shift<=max_shift))
        {
            good = true;
            if(c[idx].hyst<low_depth) { low_depth = c[idx].hyst; low_shift = c[idx].shift; high_ref = idx; }
            else if(c[idx].hyst==low_depth) { if(c[idx].shift<low_shift) { low_shift = c[idx].shift; high_ref = idx; } }
        }
        if(idx==low_lvl) break;
        idx--;
    }
    if(!good) return SPAN_NOT_FOUND;
    idx = high_ref;
    while(idx>=low_lvl)
    {
        if((c[idx].result==target)&&(c[idx].hyst==low_depth)&&(c[idx].shift==low_shift))
        {
            if(!range_lock) low_ref = idx;
            else { if(!second_lock) { tst_high = idx; second_lock = true; } tst_low = idx; }
        }
        else
        {
            range_lock = true;
            if(second_lock)
            {
                second_lock = false;
                if((tst_high-tst_low)>=(high_ref-low_ref)) { low_ref = tst_low; high_ref = tst_high; }
            }
        }
        if(idx==low_lvl) break;
        idx--;
    }
    picked = (u8)(high_ref-low_ref); picked = picked/2; picked = (u8)(low_ref+picked);
    *res = picked;
    return SPAN_OK;
}
// Binarizer::pickLevelByCRCStatsOpt (binarizer.cpp:2134-2300)
SDV_HDN u8 pick_level_by_stats_opt(const CrcH *c, u8 *res, u8 low_lvl, u8 high_lvl, u8 target, u8 max_hyst, u8 max_shift)
{
    bool range_lock = false, good = false;
    u8 idx, hold_cnt, same_cnt, low_depth, low_shift = 0, high_shift = 0, low_ref = 0, high_ref = 0, picked;
    idx = high_lvl;
    while(idx>=low_lvl)
    {
        if((c[idx].result==target)&&(c[idx].hyst<=max_hyst)&&(c[idx].shift<=max_shift))
        {
            if(!good) { good = true; low_ref = high_ref = idx; }
            else { low_ref = idx; if(low_ref==low_lvl) { low_shift = low_ref; high_shift = high_ref; range_lock = true; } }
        }
        else if(good)
        {
            if((high_ref-low_ref+1)>=(high_shift-low_shift+1)) { low_shift = low_ref; high_shift = high_ref; range_lock = true; }
            good = false;
        }
        if(idx==low_lvl) break;
        idx--;
    }
    if(range_lock) { high_lvl = high_shift; low_lvl = low_shift; }
    good = false;
    hold_cnt = 0;
    low_depth = low_shift = 255;
    same_cnt = MIN_VALID_CRCS;
    low_ref = high_ref = picked = MAX_REF_LVL;
    idx = high_lvl;
    while(idx>=low_lvl)
    {
        if((c[idx].result==target)&&(c[idx].hyst<=max_hyst)&&(c[idx].shift<=max_shift))
        {
            good = true;
            if(low_depth>c[idx].hyst) { low_depth = c[idx].hyst; low_shift = c[idx].shift; low_ref = high_ref = idx; hold_cnt = MIN_VALID_CRCS; }
            else if(low_depth==c[idx].hyst)
            {
                if(low_shift>c[idx].shift) { low_shift = c[idx].shift; low_ref = high_ref = idx; same_cnt = MIN_VALID_CRCS; hold_cnt = MIN_VALID_CRCS; }
                else if(low_shift==c[idx].shift) { low_ref = idx; same_cnt--; if(same_cnt==0) { hold_cnt = 0; break; } }
                else { hold_cnt--; if(hold_cnt==0) break; }
            }
            else { hold_cnt--; if(hold_cnt==0) break; }
        }
        if(idx==low_lvl) break;
        idx--;
    }
    if(good) { picked = (u8)(high_ref-low_ref); picked = picked/2; picked = (u8)(low_ref+picked); *res = picked; return SPAN_OK; }
    return SPAN_NOT_FOUND;
}

// ------------------------------------------------------------------------------------------------ shared work area of one line decode
// Per reference level result of the sweep plus what the sequential carry pass needs (see sweep_fixup()).
struct SweepAux { u16 out_w8; u8 did_read; u8 quirk_ok; i16 q_start, q_stop; };
struct CandLite { u16 calc_crc, w8; u8 ok; u8 pad; };
struct SweepPlan { Coord coords; u8 markers, do_read, coords_set, pad; };

struct Work
{
    Line o;                         // the output line being built
    u32 sprd[256];                  // brightness histogram
    CrcH sw[256];                   // sweep result per reference level
    SweepAux swa[256];
    CrcH stats[MAX_COLL_CRCS+1];
    MarkRes trials[MARK_TRIALS];
    Cand cand[MAX_CAND];
    u32 sweep_trials[SWEEP_MAX_LEVELS*MARK_TRIALS];    // packed marker trial of every (reference level, hysteresis) pair; then, once the levels have
                                                       // picked their coordinates, the (level, hysteresis, shift) candidate table (CandLite)
    SweepPlan plan[256];
    // scalars of the processLine state machine
    u8 proc_state, was_bw_scanned, do_sweep, hlim, slim, stage_count;
    u8 sweep_low, sweep_high;
    // scalars of the black/white search
    u8 bw_white_detected, bw_useful_low, bw_mark_white, bw_has_stop;
    u16 bw_ed_start;
};

// ------------------------------------------------------------------------------------------------ AGC
// Binarizer::findBlackWhite + findSTC007BW (binarizer.cpp:2385-2500, 2684-3070, 3116-3473).
SDV_HD u32 most_frequent_count(const u32 *s) { u32 m = 0; for(int i=255;i>=0;i--) if(s[i]>m) m = s[i]; return m; }
SDV_HD u8 usefull_low(const u32 *s)
{
    u8 lev = 0, lowest = 0; bool found = false;
    u32 mf = (u16)most_frequent_count(s)/64;
    while(lev<MAX_BLACK_LVL) { if(s[lev]>mf) { lowest = lev; found = true; break; } lev++; }
    if(!found) while(lev<MAX_BLACK_LVL) { if(s[lev]>0) { lowest = lev; break; } lev++; }
    return lowest;
}
SDV_HD u8 usefull_high(const u32 *s)
{
    u8 lev = 255, highest = 255;
    u32 mf = (u16)most_frequent_count(s)/64;
    while(lev>=MIN_WHITE_LVL) { if(s[lev]>mf) { highest = lev; break; } lev--; }
    while(lev>=MIN_WHITE_LVL) { if(s[lev]>0) { highest = lev; break; } lev--; }
    return highest;
}
SDV_HD void hist_clear(const Cta &c, u32 *sprd) { c.sync(); for(int i=c.tid;i<256;i+=c.n) sprd[i] = 0; c.sync(); }
SDV_HD void hist_add(const Cta &c, u32 *sprd, const u8 *px, int from, int to /*exclusive*/)
{
    c.sync();
    for(int i=from+c.tid;i<to;i+=c.n) hist_inc(&sprd[px[i]]);
    c.sync();
}

// Common tail of Binarizer::findBlackWhite (binarizer.cpp:3196-3473): black/white peaks of the gathered histogram.
SDV_HD void bw_pick_levels(const u32 *sprd, bool do_ref_lvl_sweep, u8 *black, u8 *white, u8 *set)
{
    u8 brt_lev, br_black, br_white, useful_low, useful_high, low_scan_limit, high_scan_limit, range_limit, bin_low, bin_high;
    u32 black_cnt, white_cnt, search_lim, t;
    bool black_det, white_det;
    useful_low = low_scan_limit = br_black = usefull_low(sprd);
    useful_high = high_scan_limit = br_white = usefull_high(sprd);
    range_limit = (u8)(high_scan_limit-low_scan_limit);
    low_scan_limit = (u8)(low_scan_limit+(range_limit/3));
    high_scan_limit = (u8)(high_scan_limit-(range_limit/3));
    t = range_limit; t = t*10/100; bin_low = (u8)t;
    t = range_limit; t = t*12/100; bin_high = (u8)t;
    search_lim = (u16)most_frequent_count(sprd)/64;
    brt_lev = useful_low; black_cnt = 0; black_det = false;
    while(brt_lev<=low_scan_limit)
    {
        if(sprd[brt_lev]>black_cnt) { black_cnt = sprd[brt_lev]; if(black_cnt>search_lim) { br_black = brt_lev; black_det = true; } }
        if(black_det) if((brt_lev-br_black)>=bin_low) break;
        brt_lev++;      // u8 wrap mirrors the reference
    }
    brt_lev = useful_high; white_cnt = 0; white_det = false;
    if(black_det)
    {
        while(brt_lev>=high_scan_limit)
        {
            if(brt_lev<(br_black+MIN_CONTRAST)) break;
            if(sprd[brt_lev]>white_cnt) { white_cnt = sprd[brt_lev]; if(white_cnt>search_lim) { br_white = brt_lev; white_det = true; } }
            if(white_det) if((br_white-brt_lev)>=bin_high) break;
            brt_lev--;
        }
    }
    if(black_det&&white_det)
    {
        bool inv = false;
        if(br_white<br_black) inv = true;
        else if((br_white-br_black)<MIN_CONTRAST) inv = true;
        else if(do_ref_lvl_sweep&&((br_white-br_black)<FINE_MIN_VALID_CRCS)) inv = true;
        else if(br_black>MAX_BLACK_LVL) inv = true;
        else if(br_white<MIN_WHITE_LVL) inv = true;
        if(inv) { black_det = white_det = false; br_black = useful_low; br_white = useful_high; }
    }
    *black = br_black;
    *white = br_white;
    *set = (black_det&&white_det) ? 1 : 0;
}

SDV_HD void find_black_white_cta(const Cta &c, Work *w, const u8 *px, const Geom &g, bool do_ref_lvl_sweep)
{
    Line *out = &w->o;
    u32 *sprd = w->sprd;
    int ppb = g.est_ppb;
    hist_clear(c, sprd);
    // Rough white level from the marker areas (left 10 bits, right 20 bits).
    hist_add(c, sprd, px, 0, (u16)(ppb*10));
    hist_add(c, sprd, px, (u16)(g.scan_end-ppb*20), g.scan_end+1);
    if(c.tid==0)
    {
        u8 useful_low = usefull_low(sprd);
        u8 useful_high = usefull_high(sprd), high_scan_limit = useful_high, br_mark_white = useful_high;
        u8 range_limit = (u8)(high_scan_limit-useful_low);
        high_scan_limit = (u8)(high_scan_limit-(range_limit/4));
        u8 bin_high = range_limit/8;
        u8 brt_lev = useful_high;
        u32 white_lvl_count = 0;
        bool white_detected = false;
        while(brt_lev>=high_scan_limit)
        {
            if(sprd[brt_lev]>white_lvl_count) { white_lvl_count = sprd[brt_lev]; br_mark_white = brt_lev; white_detected = true; }
            if(white_detected) if((br_mark_white-brt_lev)>=bin_high) break;
            brt_lev--;      // u8 wrap mirrors the reference
        }
        w->bw_white_detected = white_detected; w->bw_useful_low = useful_low; w->bw_mark_white = br_mark_white;
    }
    // Histogram of the centre 3/4 of the line.
    {
        u16 pixel_limit = (u16)(g.scan_end-0);
        u32 t = pixel_limit/8;
        hist_clear(c, sprd);
        hist_add(c, sprd, px, (u16)t, (u16)(g.scan_end-(u16)t));
    }
    if(w->bw_white_detected)
    {
        if(c.tid==0)
        {   // Rough STOP marker search with the provisional reference level.
            u8 bin_level = pick_center_ref(w->bw_useful_low, w->bw_mark_white);
            u8 stage = MARK_ED_START, pv;
            u32 pixel, pixel_limit;
            u16 ed_s = 0, ed_e = 0;
            if(g.mark_end_min>(ppb*6)) pixel_limit = (u16)(g.mark_end_min-ppb*6);
            else pixel_limit = 0;
            pixel = g.scan_end;
            while(pixel>pixel_limit)
            {
                pv = px[pixel];
                if(stage==MARK_ED_START)
                {
                    if(pixel<g.mark_end_min) break;
                    if(pv>=bin_level) { ed_e = (u16)(pixel+1); stage = MARK_ED_TOP; }
                }
                else if(stage==MARK_ED_TOP)
                {
                    if(pv<bin_level)
                    {
                        ed_s = (u16)(pixel+1);
                        stage = MARK_ED_BOT;
                        if((ed_e-ed_s)>=(ppb*2)) { stage = MARK_ED_LEN_OK; break; }
                        else stage = MARK_ED_START;
                    }
                }
                pixel--;
            }
            out->med = stage;
            out->coords.stop = (i16)ed_s;
            out->m_stop = ed_e;
            w->bw_has_stop = line_has_stop(out);
            w->bw_ed_start = ed_s;
        }
        c.sync();
        if(w->bw_has_stop)
        {   // Histogram of the 64 bit cells before the STOP marker.
            u16 ed_s = w->bw_ed_start;
            u16 search_lim = (u16)(ppb*64);
            if(search_lim>ed_s) search_lim = g.mark_start_max;
            else search_lim = (u16)(ed_s-search_lim);
            hist_clear(c, sprd);
            int from = (int)search_lim+1, to = (int)ed_s;      // pixels ed_s-1 down to search_lim+1
            int cnt = (to>from) ? (to-from) : 0;
            hist_add(c, sprd, px, from, to);
            if(cnt<32)
            {
                u16 pl = (u16)(g.scan_end/8);
                hist_add(c, sprd, px, pl, (u16)(g.scan_end-pl));
            }
        }
    }
    c.sync();
    if(c.tid==0)
    {
        u8 bl, wh, st;
        bw_pick_levels(sprd, do_ref_lvl_sweep, &bl, &wh, &st);
        w->was_bw_scanned = 1;
        out->black = bl; out->white = wh; out->bw_set = st;
    }
    c.sync();
}

// ------------------------------------------------------------------------------------------------ reference level sweep
// One reference level of Binarizer::sweepRefLevel (binarizer.cpp:3551-3817), evaluated as if the CRC word carried over
// from the previous (higher) level were non-zero; sweep_fixup() replays the carry afterwards.
// The level is done in three steps so that the bit-cell sampling -- nearly all of the work -- can be spread over every
// thread of the block instead of one thread per level:
//   sweep_level_plan   : coordinates the level reads its data with (markers of this level, else the preset coordinates)
//   eval_cand_lite     : one (level, hysteresis, shift) candidate of readPCMdata: CRC as read, CRC as computed
//   sweep_level_finish : readPCMdata's first-valid search replayed over the level's candidates, result of the level

SDV_HD void sweep_level_plan(Coord def_coord, const u32 *trials /*[MARK_TRIALS] packed marker trials of this level*/, SweepPlan *pl, SweepAux *aux)
{
    aux->quirk_ok = 0; aux->q_start = aux->q_stop = 0; aux->did_read = 0; aux->out_w8 = 0;
    Coord mc = coord_none();
    const bool markers = pick_packed_trial(trials, &mc);
    Coord lc; lc.start = 0; lc.stop = 0;                // coordinates of the cleared dummy line (line_base_clear)
    { Line t; line_clear(&t); line_base_clear(&t); lc = t.coords; }
    if(markers&&(mc.stop>mc.start)) lc = mc;
    pl->markers = markers ? 1 : 0; pl->coords_set = markers ? 1 : 0; pl->do_read = 0;
    if(coord_valid(def_coord))
    {
        if(!markers) { lc = def_coord; pl->do_read = 1; }
        else if(coord_valid(lc)) { aux->quirk_ok = 1; aux->q_start = lc.start; aux->q_stop = lc.stop; }   // markers found, nothing read yet: with a carried CRC word of 0x0000 the reference takes this level as valid
    }
    if(markers) pl->do_read = 1;
    pl->coords = lc;
}
SDV_HD void eval_cand_lite(const u8 *px, int pixel_stop, Ppb ppb, u8 ref, u8 black, u8 white, int hyst, int shift, CandLite *cd)
{
    const u8 low = get_low_level(ref, (u8)hyst), high = get_high_level(ref, (u8)hyst);
    cd->ok = 0; cd->calc_crc = 0; cd->w8 = 0;
    if((low<=black)||(high>=white)) return;
    u16 words[9];
    fill_stc007(px, pixel_stop, ppb, shift, low, high, words);
    cd->calc_crc = crc_stc007(words); cd->w8 = words[8]; cd->ok = 1;
}
struct LiteCandFn
{
    const CandLite *tab; int slim; u8 ref; Cand *tmp;
    SDV_HD const Cand *operator()(int h, int s) const
    {
        const CandLite &c = tab[h*(slim+1)+s];
        tmp->low = get_low_level(ref, (u8)h); tmp->high = get_high_level(ref, (u8)h);
        tmp->ok = c.ok; tmp->calc_crc = c.calc_crc; tmp->words[8] = c.w8;
        return tmp;
    }
};
SDV_HD void sweep_level_finish(const SweepPlan *pl, u8 black_lvl, u8 white_lvl, int ref_index, int hlim, int slim, const CandLite *cands, CrcH *res, SweepAux *aux)
{
    Line d;
    line_clear(&d);
    line_base_clear(&d);
    d.black = black_lvl; d.white = white_lvl;
    d.ref = (u8)ref_index;
    d.coords = pl->coords; d.coords_set = pl->coords_set;
    const bool did_read = pl->do_read!=0;
    if(did_read)
    {
        if(hlim>HYST_DEPTH_MAX) hlim = HYST_DEPTH_MAX;
        if(slim>SHIFT_MAX) slim = SHIFT_MAX;
        Cand tmp;
        for(int i=0;i<9;i++) tmp.words[i] = 0;
        LiteCandFn f; f.tab = cands; f.slim = slim; f.ref = d.ref; f.tmp = &tmp;
        d.ppb = make_ppb(d.coords);
        read_pcm_core(&d, hlim, slim, f);
    }
    if(d.hyst>0x0F) d.hyst = 0x0F;
    if(did_read&&line_crc_ok(&d)&&coord_valid(d.coords))
    {
        res->result = REF_CRC_OK; res->start = d.coords.start; res->stop = d.coords.stop;
        res->hyst = d.hyst; res->shift = d.shift; res->crc = d.calc_crc;
    }
    else if(d.coords_set)
    {
        res->result = REF_BAD_CRC; res->start = d.coords.start; res->stop = d.coords.stop;
        res->hyst = d.hyst; res->shift = d.shift; res->crc = d.calc_crc;
    }
    aux->did_read = did_read ? 1 : 0;
    aux->out_w8 = d.words[8];
}

// Sequential carry of the previous level's CRC word (dummy_line->clear() through the base pointer keeps words[] while
// zeroing calc_crc, binarizer.cpp:3615): while the carried word is 0x0000 a level whose markers were found is taken as
// "already valid" (CRC 0, hysteresis 0, shift 0) and a level without preset coordinates is skipped.
SDV_HD void sweep_fixup(CrcH *sw, const SweepAux *swa, int low_lvl, int high_lvl, bool def_coord_valid)
{
    u16 cw = (u16)~0xA96A;      // STC007Line constructor state
    for(int ref=high_lvl;ref>=low_lvl;ref--)
    {
        if(cw==0)
        {
            if(def_coord_valid)
            {
                if(swa[ref].quirk_ok)
                {
                    sw[ref].result = REF_CRC_OK; sw[ref].start = swa[ref].q_start; sw[ref].stop = swa[ref].q_stop;
                    sw[ref].hyst = 0; sw[ref].shift = 0; sw[ref].crc = 0;
                    continue;       // nothing read: the carried word stays 0
                }
                // markers not found: the level read its data with the preset coordinates as usual
            }
            else
            {   // no preset coordinates: the level is skipped entirely, its entry stays reset
                sw[ref].result = 0; sw[ref].start = sw[ref].stop = 0; sw[ref].crc = 0; sw[ref].hyst = sw[ref].shift = 0x0f;
                continue;
            }
        }
        if(swa[ref].did_read) cw = swa[ref].out_w8;
    }
}

// Binarizer::calcRefLevelBySweep after the sweep itself (binarizer.cpp:3821-4120); thread 0 only.
// Returns true when the marker search has to be repeated on the output line.
SDV_HD bool sweep_select(Work *w, const BinState *b, const Geom &g)
{
    Line *l = &w->o;
    CrcH *sw = w->sw;
    CrcH *stats = w->stats;
    u8 valid_cnt = 0, span_res = SPAN_NOT_FOUND;
    u8 fast_ref = pick_center_ref(l->black, l->white);
    bool refind = false;
    reset_crc_stats(stats, MAX_COLL_CRCS+1);
    stats[0].hyst = 0; stats[0].shift = 0;
    for(u8 lvl=(u8)(l->white-1);lvl>l->black;lvl--)
        if(sw[lvl].result==REF_CRC_OK) update_crc_stats(stats, sw[lvl], &valid_cnt);
    if(valid_cnt>0)
    {
        find_most_frequent_crc(stats, &valid_cnt, true);
        invalidate_non_frequent(sw, (u8)(l->black+1), (u8)(l->white-1), valid_cnt, stats[0].crc);
        if(valid_cnt>0)
        {
            if(stats[0].result<FINE_MIN_VALID_CRCS) span_res = SPAN_TOO_NARROW;
            else span_res = pick_level_by_stats(sw, &l->ref, (u8)(l->black+1), (u8)(l->white-1), REF_CRC_OK, 0x0F, SHIFT_MAX);
        }
    }
    if(span_res==SPAN_OK)
    {
        CrcH t = sw[l->ref];
        l->sweeped = 1;
        line_coord_set(l, t.start, t.stop);
        l->coords_set = 1;
        refind = true;
        w->hlim = t.hyst;
        if(w->hlim>HYST_DEPTH_MAX) w->hlim = HYST_DEPTH_MAX;
        w->slim = t.shift;
    }
    else
    {
        if(span_res==SPAN_TOO_NARROW)
        {
            span_res = pick_level_by_stats_opt(sw, &l->ref, (u8)(l->black+1), (u8)(l->white-1), REF_CRC_OK, w->hlim, w->slim);
            l->forced_bad = 1;
        }
        else
            span_res = pick_level_by_stats(sw, &l->ref, (u8)(l->black+1), (u8)(l->white-1), REF_BAD_CRC, 0xFF, 0xFF);
        if(span_res==SPAN_OK)
        {
            CrcH t = sw[l->ref];
            line_coord_set(l, t.start, t.stop);
            l->coords_set = 1;
            refind = true;
        }
        else if(bin_ref_preset(b))
        {
            l->ref = b->def_ref;
            if(coord_valid(b->def_coord)) l->coords = b->def_coord;
        }
        else
        {
            l->ref = fast_ref;
            if(!coord_valid(b->def_coord)) line_coord_set(l, g.est_ppb, g.scan_end-(4*g.est_ppb));
            else l->coords = b->def_coord;
        }
        w->hlim = HYST_DEPTH_MIN;
        w->slim = SHIFT_MIN;
    }
    return refind;
}

// ------------------------------------------------------------------------------------------------ processLine
// Binarizer::processLine for a non-empty, non-service STC-007 line (binarizer.cpp:443-1724).  All threads of the
// block call this with identical arguments; [w] and [b] live in shared memory; the result is w->o.
SDV_HD void process_line_cta(const Cta &c, Work *w, const BinState *b, const u8 *px, const Geom &g)
{
    Line *o = &w->o;
    c.sync();
    if(c.tid==0)
    {
        line_clear(o);
        line_coord_set(o, 0, g.scan_end);
        w->proc_state = STG_REF_FIND;
        w->was_bw_scanned = 0;
        w->do_sweep = 0;
        if(bin_bw_preset(b)) { o->black = b->def_black; o->white = b->def_white; o->bw_set = 1; }
        if(bin_ref_preset(b)) { if(coord_valid(b->def_coord)) w->proc_state = STG_INPUT_ALL; else w->proc_state = STG_INPUT_LEVEL; }
        w->hlim = b->max_hyst;
        w->slim = b->max_shift;
        w->stage_count = 0;
    }
    c.sync();
    for(;;)
    {
        c.sync();
        int st = w->proc_state;
        c.sync();
        if(c.tid==0) w->stage_count++;
        if(st==STG_INPUT_ALL)
        {
            if(!o->bw_set) find_black_white_cta(c, w, px, g, w->do_sweep!=0);
            c.sync();
            bool go_read = false;
            if(c.tid==0)
            {
                o->coords = b->def_coord;
                o->ref = b->def_ref;
                o->sweeped = 0;
                if(!o->bw_set) w->proc_state = STG_NO_GOOD;
                else if((b->def_ref>=o->white)||(b->def_ref<=o->black)) w->proc_state = STG_REF_FIND;
                else w->proc_state = 0xFF;
            }
            c.sync();
            go_read = (w->proc_state==0xFF);
            c.sync();
            if(go_read)
            {
                read_pcm_cta(c, px, g, o, w->hlim, w->slim, w->cand);
                if(c.tid==0)
                {
                    if(line_crc_ok(o)) { o->by_ext = 1; w->proc_state = STG_DATA_OK; }
                    else w->proc_state = STG_INPUT_LEVEL;
                }
            }
        }
        else if(st==STG_INPUT_LEVEL)
        {
            if(!w->was_bw_scanned) find_black_white_cta(c, w, px, g, w->do_sweep!=0);
            c.sync();
            if(c.tid==0)
            {
                line_coord_set(o, 0, g.scan_end);
                o->ref = b->def_ref;
                o->sweeped = 0;
                if(!o->bw_set) w->proc_state = STG_NO_GOOD;
                else
                {
                    w->proc_state = STG_REF_FIND;
                    if((b->def_ref<o->white)&&(b->def_ref>o->black)) w->proc_state = 0xFF;
                }
            }
            c.sync();
            bool go_search = (w->proc_state==0xFF);
            c.sync();
            if(go_search)
            {
                // do_coord_search is always on for STC-007 in VideoToDigital (videotodigital.cpp:966-972).
                find_coordinates_cta(c, px, g, o, w->trials);
                if(c.tid==0)
                {
                    w->proc_state = STG_REF_FIND;
                    if(line_has_markers(o)&&((!coord_valid(b->def_coord))||(!coord_eq(o->coords, b->def_coord)))) w->proc_state = 0xFE;
                }
                c.sync();
                bool go_read = (w->proc_state==0xFE);
                c.sync();
                if(go_read)
                {
                    read_pcm_cta(c, px, g, o, w->hlim, w->slim, w->cand);
                    if(c.tid==0)
                    {
                        if(line_crc_ok(o)) { o->by_ext = 1; w->proc_state = STG_DATA_OK; }
                        else w->proc_state = STG_REF_FIND;
                    }
                }
            }
        }
        else if(st==STG_REF_FIND)
        {
            if(!w->was_bw_scanned) find_black_white_cta(c, w, px, g, w->do_sweep!=0);
            c.sync();
            if(c.tid==0)
            {
                if(!o->bw_set) w->proc_state = STG_NO_GOOD;
                else
                {
                    w->do_sweep = ((b->mode==SDV_MODE_NORMAL)||(b->mode==SDV_MODE_INSANE)) ? 1 : 0;
                    if(w->do_sweep) w->proc_state = STG_REF_SWEEP_RUN;
                    else
                    {
                        w->hlim = HYST_DEPTH_SAFE; w->slim = SHIFT_MIN;
                        o->ref = pick_center_ref(o->black, o->white);
                        w->proc_state = 0xFF;
                    }
                }
            }
            c.sync();
            bool go_search = (w->proc_state==0xFF);
            c.sync();
            if(go_search)
            {
                find_coordinates_cta(c, px, g, o, w->trials);
                if(c.tid==0)
                {
                    if(!line_has_markers(o)) { w->hlim = HYST_DEPTH_SAFE; w->slim = SHIFT_MIN; }
                    else { w->hlim = b->max_hyst; w->slim = b->max_shift; }
                    w->proc_state = STG_READ_PCM;
                }
            }
        }
        else if(st==STG_REF_SWEEP_RUN)
        {
            // calcRefLevelBySweep: one reference level per thread.
            if(c.tid==0)
            {
                w->hlim = b->max_hyst; w->slim = b->max_shift;
                u8 lo = (u8)(o->black+1), hi = (u8)(o->white-1);
                if(MIN_REF_LVL>lo) lo = MIN_REF_LVL;
                if(MAX_REF_LVL<hi) hi = MAX_REF_LVL;
                w->sweep_low = lo; w->sweep_high = hi;
            }
            c.sync();
            for(int i=c.tid;i<256;i+=c.n)
            {
                reset_crc_stats(&w->sw[i], 1);
                w->swa[i].did_read = 0; w->swa[i].quirk_ok = 0; w->swa[i].out_w8 = 0; w->swa[i].q_start = w->swa[i].q_stop = 0;
            }
            c.sync();
            {
                int lo = w->sweep_low, hi = w->sweep_high, hl = w->hlim, sl = w->slim;
                u8 bl = (u8)lo, wh = (u8)hi;      // dummy line black/white = sweep limits (binarizer.cpp:3622-3623)
                // every (reference level, marker hysteresis) trial on its own thread ...
                const int ntr = (hi>=lo) ? ((hi-lo+1)*MARK_TRIALS) : 0;
                for(int t=c.tid;t<ntr;t+=c.n)
                    w->sweep_trials[t] = pack_trial(search_markers(px, g, (u8)(hi-t/MARK_TRIALS), (u8)(t%MARK_TRIALS)));
                c.sync();
                // ... one reference level per thread: which coordinates the level reads its data with ...
                for(int ref=hi-c.tid;ref>=lo;ref-=c.n)
                    sweep_level_plan(b->def_coord, &w->sweep_trials[(hi-ref)*MARK_TRIALS], &w->plan[ref], &w->swa[ref]);
                c.sync();
                // ... every (level, hysteresis, shift) candidate on its own thread (the table reuses the trial array, a chunk
                // of levels at a time when it does not hold them all), then the first-valid search per level
                const int hl2 = (hl>HYST_DEPTH_MAX) ? HYST_DEPTH_MAX : hl, sl2 = (sl>SHIFT_MAX) ? SHIFT_MAX : sl;
                const int ncand = (hl2+1)*(sl2+1);
                CandLite *tab = (CandLite *)w->sweep_trials;
                const int cap = (int)(sizeof(w->sweep_trials)/sizeof(CandLite));
                const int per = (cap/ncand>0) ? (cap/ncand) : 1;
                for(int top=hi;top>=lo;top-=per)
                {
                    const int nlev = (top-lo+1<per) ? (top-lo+1) : per;
                    for(int t=c.tid;t<nlev*ncand;t+=c.n)
                    {
                        const int ref = top-t/ncand, k = t%ncand;
                        const SweepPlan *pl = &w->plan[ref];
                        if(pl->do_read) eval_cand_lite(px, g.W-1, make_ppb(pl->coords), (u8)ref, bl, wh, k/(sl2+1), k%(sl2+1), &tab[t]);
                    }
                    c.sync();
                    for(int i=c.tid;i<nlev;i+=c.n)
                        sweep_level_finish(&w->plan[top-i], bl, wh, top-i, hl, sl, &tab[i*ncand], &w->sw[top-i], &w->swa[top-i]);
                    c.sync();
                }
            }
            c.sync();
            bool refind = false;
            if(c.tid==0)
            {
                sweep_fixup(w->sw, w->swa, w->sweep_low, w->sweep_high, coord_valid(b->def_coord));
                refind = sweep_select(w, b, g);
                w->proc_state = refind ? 0xFF : STG_READ_PCM;
            }
            c.sync();
            refind = (w->proc_state==0xFF);
            c.sync();
            if(refind)
            {
                find_coordinates_cta(c, px, g, o, w->trials);
                if(c.tid==0) w->proc_state = STG_READ_PCM;
            }
        }
        else if(st==STG_READ_PCM)
        {
            bool do_read = o->coords_set!=0;
            c.sync();
            if(do_read) read_pcm_cta(c, px, g, o, w->hlim, w->slim, w->cand);
            c.sync();
            if(c.tid==0)
            {
                if(line_crc_ok(o)) w->proc_state = STG_DATA_OK;
                else
                {
                    w->proc_state = STG_NO_GOOD;
                    if(coord_valid(b->def_coord)&&(!w->do_sweep)&&(!o->forced_bad)&&(!o->coords_set))
                        if(!coord_eq(o->coords, b->def_coord))
                        {
                            o->coords = b->def_coord;
                            o->m_bg = 0; o->m_ed = 0; o->m_stop = 0;
                            w->proc_state = 0xFF;
                        }
                }
            }
            c.sync();
            bool retry = (w->proc_state==0xFF);
            c.sync();
            if(retry)
            {
                read_pcm_cta(c, px, g, o, w->hlim, w->slim, w->cand);
                if(c.tid==0) w->proc_state = line_crc_ok(o) ? STG_DATA_OK : STG_NO_GOOD;
            }
        }
        else if(st==STG_DATA_OK)
        {
            bool done = false;
            if(c.tid==0)
            {
                if(o->forced_bad) w->proc_state = STG_NO_GOOD;
                else
                {
                    o->wflags = line_crc_ok(o) ? 1 : 0;
                    if(words_control_block(o->words)) line_set_serv_ctrl_blk(o);
                    w->proc_state = 0xFD;
                }
            }
            c.sync();
            done = (w->proc_state==0xFD);
            c.sync();
            if(done) break;
        }
        else
        {   // STG_NO_GOOD
            if(c.tid==0)
            {
                if(line_crc_ok(o)) line_set_invalid_crc(o);
                o->wflags = line_crc_ok(o) ? 1 : 0;
            }
            break;
        }
        c.sync();
        bool overrun = w->stage_count>STG_MAX;
        c.sync();
        if(overrun) break;
    }
    c.sync();
}

}   // namespace sdv
