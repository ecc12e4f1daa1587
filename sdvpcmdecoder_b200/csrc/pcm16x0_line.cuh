// pcm16x0_line.cuh -- PCM-16x0 line decode (the Binarizer operator for PCM16X0SubLine) as cooperative integer code.
//
// A PCM-16x0 video line is three sub-lines of 64 bit cells (3 x 16-bit words + CRCC) with one control bit between the
// middle and the right one; the Binarizer is run once per part.  No markers: the data coordinates come from a 21 x 21
// grid of (start, stop) offsets, every grid point read as all three parts (Binarizer::searchPCM16X0Data,
// binarizer.cpp:4514-5271), once per video line.  Here all 441 x 3 reads run on their own threads; the reference's
// early-exit scan, per-part CRC votes and the row/column selection are then replayed over the stored results, one grid
// row per thread, and the sequential pieces (the break-out rules, the line state the last visited read leaves behind,
// a line forced bad by a bit-picker CRC collision) are restored from them.  With Cta{0,1} it is a sequential program.
#pragma once
#include "pcm1_line.cuh"

namespace sdv {

enum { X0L_BITS = 193, X0L_PART_BITS = 64, X0L_WORDS = 4, X0L_CRC_SILENT = 0x0E10, X0L_CTRL_BIT = 128 };       // pcm16x0subline.h:77-110
enum { X0L_SEARCH_STEP_DIV = 2, X0L_SEARCH_MAX_OFS = 10, X0L_SEARCH_STEP_CNT = (X0L_SEARCH_MAX_OFS+1)*2,       // binarizer.h:262-264
       X0L_GRID = 2*X0L_SEARCH_MAX_OFS+1 };
enum { X0L_LEFT = 0, X0L_MIDDLE = 1, X0L_RIGHT = 2 };                                                           // PCM16X0SubLine::PART_*

struct X0Line
{
    u16 words[X0L_WORDS];               // R1P1L1 L2P2R2 R3P3L3 CRCC
    u16 calc_crc;
    Coord coords;
    u8 black, white, ref_low, ref, ref_high, hyst, shift, service;
    u8 picked_left, picked_right, control_bit, line_part;
    u8 sweeped, coord_sweeped, by_ext, bw_set, coords_set, forced_bad;
    u16 queue_order;
    Ppb ppb;
};

SDV_HD u16 x0_calc_crc(const u16 *w)
{
    u16 c = 0xFFFF;
    for(int i=0;i<3;i++) { c = crc16_byte(c, (u32)(w[i]>>8)); c = crc16_byte(c, (u32)(w[i]&0xFF)); }
    return c;
}
SDV_HD bool x0_crc_ok_ign(const X0Line *l) { return l->calc_crc==l->words[3]; }
SDV_HD bool x0_crc_ok(const X0Line *l) { return (!l->forced_bad)&&x0_crc_ok_ign(l); }
SDV_HD void x0_set_invalid_crc(X0Line *l) { l->words[3] = (u16)~l->calc_crc; }
SDV_HD int x0_get_ppb(const X0Line *l) { return (int)((l->ppb.psm/INT_CALC_MULT)&0xFF); }

SDV_HD void x0_clear(X0Line *l)
{   // PCM16X0SubLine::clear (pcm16x0subline.cpp:62-85)
    l->black = l->white = l->ref_low = l->ref = l->ref_high = 0;
    l->coords = coord_none();
    l->hyst = l->shift = 0;
    l->sweeped = l->coord_sweeped = l->by_ext = 0;
    l->bw_set = l->coords_set = l->forced_bad = 0;
    l->service = 0;
    l->ppb.psm = INT_CALC_MULT; l->ppb.half = INT_CALC_MULT/2; l->ppb.ofs = 0;
    l->control_bit = 1; l->line_part = X0L_LEFT; l->picked_left = l->picked_right = 0; l->queue_order = 0;
    l->words[0] = l->words[1] = l->words[2] = 0;
    l->calc_crc = X0L_CRC_SILENT;
    x0_set_invalid_crc(l);
}

// PCMLine::setPPB / getVideoPixeBylCalc for the 193 bit cells between the PCM-16x0 data coordinates.
SDV_HD Ppb x0_make_ppb(Coord c)
{
    Ppb p;
    p.psm = (u32)(c.stop-c.start);
    p.psm = (p.psm*INT_CALC_MULT+X0L_BITS/2)/X0L_BITS;
    p.ofs = c.start;
    p.half = (p.psm+1)/2;
    return p;
}
SDV_HD int x0_part_start_bit(int part) { return (part==X0L_LEFT) ? 0 : ((part==X0L_MIDDLE) ? X0L_PART_BITS : (2*X0L_PART_BITS+1)); }

// Binarizer::fillPCM16X0 (binarizer.cpp:7134-7320) without the control bit.
SDV_HDN void x0_fill(const u8 *px, int pixel_stop, Ppb ppb, int part, int shift_stage, u8 low_ref, u8 high_ref, u16 *words /*[4]*/)
{
    bool prev_high = false;
    int sh = pix_shift(shift_stage);
    int bit = x0_part_start_bit(part);
    for(int w=0;w<X0L_WORDS;w++)
    {
        u32 acc = 0;
        for(int k=0;k<16;k++, bit++)
        {
            u8 pv = px[p1_pixel_of_bit(ppb, bit, sh, pixel_stop)];
            bool one;
            if(!prev_high) { one = pv>low_ref; if(one) prev_high = true; }
            else { one = pv>=high_ref; if(!one) prev_high = false; }
            acc = (acc<<1)|(one ? 1u : 0u);
        }
        words[w] = (u16)acc;
    }
}

// Binarizer::pickCutBitsUpPCM16X0 (binarizer.cpp:6599-7011): left part -> leading bits of the first word, right part ->
// trailing bits of the CRCC, middle part -> nothing.
SDV_HDN void x0_pick_cut_bits(X0Line *l, int mode, int part, int pixel_stop, int scan_end)
{
    l->picked_left = l->picked_right = 0;
    if(part==X0L_MIDDLE) return;
    const bool left = (part==X0L_LEFT);
    const int half = (x0_get_ppb(l)+1)/2;
    int max_cut = left ? P1_LEFT_BIT_PICK : P1_RIGHT_BIT_PICK; if(mode==SDV_MODE_DRAFT) max_cut = max_cut/2;
    int cnt = 0, first = left ? 0 : scan_end;
    for(int idx=0;idx<max_cut;idx++)
    {
        int cur = p1_pixel_of_bit(l->ppb, left ? idx : (X0L_BITS-1-idx), 0, pixel_stop);
        if((left ? (cur-first) : (first-cur))>=half) break;
        if(idx==0) first = cur;
        cnt = idx+1;
    }
    if(x0_crc_ok(l)) { if(left) l->picked_left = (u8)cnt; else l->picked_right = (u8)cnt; return; }
    if(cnt==0) return;
    const int wi = left ? 0 : 3;
    const u16 orig = l->words[wi];
    const int rep = 1<<cnt;
    const u16 clean = left ? (u16)(orig&(u16)~((rep-1)<<(16-cnt))) : (u16)(orig&(u16)~(rep-1));
    bool found = false, coll = false;
    u16 fix = 0;
    // (no CRC per candidate: a left patch changes the computed CRC linearly -- message bit t contributes X0_CRC_BIT[t] --, a right
    // patch replaces the low bits of the CRC that was read; see p1_pick_cut_bits)
    if(!l->forced_bad)
    {
        const u16 X0_CRC_BIT[4] = { 0xD420, 0x6A10, 0x3508, 0x1A84 };
        if(left)
        {
            l->words[0] = clean;
            const u16 base = x0_calc_crc(l->words);
            for(int i=0;i<rep;i++)
            {
                u16 target = base;
                for(int t=0;t<cnt;t++) if((i>>(cnt-1-t))&1) target ^= X0_CRC_BIT[t];
                if(target==l->words[3])
                {
                    if(found) { coll = true; break; }
                    found = true; fix = (u16)(i<<(16-cnt));
                }
            }
        }
        else
        {   // the computed CRC is what it is: the one patch that equals its low bits fits, if the rest of the word does
            const u16 calc = x0_calc_crc(l->words);
            if((u16)(calc&(u16)~(rep-1))==clean) { found = true; fix = (u16)(calc&(u16)(rep-1)); }
        }
    }
    if(coll||(!found))
    {
        l->words[wi] = orig;
        l->calc_crc = x0_calc_crc(l->words);
        if(coll) l->forced_bad = 1;
        return;
    }
    l->words[wi] = (u16)(clean|fix);
    l->calc_crc = x0_calc_crc(l->words);
    if(left) l->picked_left = (u8)cnt; else l->picked_right = (u8)cnt;
}

// Binarizer::fillDataWords for PCM-16x0 (binarizer.cpp:7560-7670).
SDV_HD bool x0_fill_data_words(const u8 *px, const Geom &g, int mode, int part, X0Line *l, int hyst, int shift)
{
    u8 low = get_low_level(l->ref, (u8)hyst), high = get_high_level(l->ref, (u8)hyst);
    l->ref_low = low; l->ref_high = high;
    if((low<=l->black)||(high>=l->white)) { x0_set_invalid_crc(l); return false; }
    l->hyst = (u8)hyst; l->shift = (u8)shift;
    x0_fill(px, g.W-1, l->ppb, part, shift, low, high, l->words);
    l->calc_crc = x0_calc_crc(l->words);
    l->control_bit = 1;
    if(x0_crc_ok(l)) { if(px[p1_pixel_of_bit(l->ppb, X0L_CTRL_BIT, pix_shift(shift), g.W-1)]<l->ref) l->control_bit = 0; }
    x0_pick_cut_bits(l, mode, part, g.W-1, g.scan_end);
    return true;
}

// Binarizer::readPCMdata for one part of a PCM-16x0 line (see p1_read_pcm).
SDV_HDN void x0_read_pcm(const u8 *px, const Geom &g, int mode, int part, X0Line *l, int hlim, int slim)
{
    if(hlim>HYST_DEPTH_MAX) hlim = HYST_DEPTH_MAX;
    if(slim>SHIFT_MAX) slim = SHIFT_MAX;
    l->ppb = x0_make_ppb(l->coords);
    int win_h = 0, win_s = 0;
    if(!l->sweeped)
    {
        bool found = false;
        X0Line first;                   // the line as the (0,0) fill left it (see p1_read_pcm)
        for(int h=0;(h<=hlim)&&(!found);h++)
        {
            bool invalid_hyst = false;
            for(int s=0;s<=slim;s++)
            {
                const bool filled = x0_fill_data_words(px, g, mode, part, l, h, s);
                if((h==0)&&(s==0)) first = *l;
                if(!filled) { invalid_hyst = true; break; }
                if(x0_crc_ok(l)) { found = true; win_h = h; win_s = s; break; }
            }
            if(invalid_hyst) break;
        }
        if(found&&(win_h==l->hyst)&&(win_s==l->shift)&&(!l->forced_bad)) return;
        if(!found)
        {
            const u8 fb = l->forced_bad;
            *l = first;
            l->forced_bad = (u8)(fb|first.forced_bad);
            return;
        }
    }
    else { win_h = hlim; win_s = slim; }
    x0_fill_data_words(px, g, mode, part, l, win_h, win_s);
}

// ------------------------------------------------------------------------------------------------ coordinate search
enum { X0F_DIAGS = 2*(X0L_GRID-1)+1, X0F_LANES = X0L_GRID+2 };
// Outcome of one grid row (one left offset) of searchPCM16X0Data.
struct X0Row
{
    CrcH best;              // scan_right_res[right_ofs] when valid
    u8 valid;               // valid_right_crcs > 0 after the selection
    u8 parts_ok;            // number of parts with a valid CRC at right_ofs (after the per-part votes)
    u8 lock_left;           // a column of this row had all three parts valid
    u8 last_j;              // last column the scan visited
    u8 saw_coll;            // a visited read ended in a bit-picker CRC collision (line forced bad from there on)
    u8 pad[3];
};

struct X0Work
{
    X0Line o;
    X0Line last;
    u32 sprd[256];
    CrcH grid[X0L_GRID][X0L_GRID][3];           // every (left offset, right offset, part) read, as the reference records it
    u8 coll[X0L_GRID][X0L_GRID][3];
    X0Row rows[X0L_GRID];
    CrcH left_res[X0L_SEARCH_STEP_CNT];
    u8 proc_state, was_bw_scanned, hlim, slim, stage_count, do_coord_search, search_ok, scan_done;
    i16 s_left_start, s_right_stop, s_step;
    Coord s_data_loc;
    int s_any_coll;
    int s_next;                                 // next read to hand out (dynamic distribution)
    // MODE_INSANE: reference level sweep (Binarizer::sweepRefLevel for PCM16X0SubLine)
    CrcH sw[256];
    X0Line sweep_d, sweep_save;
    u8 do_sweep, sweep_low, sweep_high, pad1;
    // bit-sliced grid search (x0_search_fills_cta, see p1_search_fills_cta): per anti-diagonal one 32-lane word per bit cell
    u32 f_gbits[P1F_WORDS], f_ebits[P1F_WORDS];
    u32 f_top[X0F_DIAGS][4];                    // the four leading bits of the left part (what the bit picker may replace)
    u32 f_read[X0F_DIAGS][3][16];               // CRCC as read, per part
    u32 f_calc[X0F_DIAGS][2][16];               // CRC computed, left and right part (what their bit pickers start from)
    u32 f_valid[X0F_DIAGS][3];
    u8 rp_flags[MAX_CAND+1]; u8 rp_go, rp_win, rp_forced;          // x0_read_pcm_cta: per (hysteresis, shift) candidate: bit 0 filled, 1 CRC valid, 2 collision
};

// x0_read_pcm by the whole group: the (hysteresis, shift) candidates are independent fills -- a fill rewrites everything of the
// sub-line but the forced-bad state, which only a bit-picker collision sets -- so every candidate is tried on its own thread from
// the entry state and the reference's loop is replayed over three flag bits per candidate: stop at the first hysteresis depth whose
// levels touch black / white, the first candidate with a valid CRC wins unless a collision came before it.  The line is then
// filled once more with the winner (or with (0, 0), keeping a collision's forced-bad mark): one fill instead of up to 56 in a row.
SDV_HD void x0_read_pcm_cta(const Cta &c, X0Work *w, const u8 *px, const Geom &g, int mode, int part, int hlim, int slim)
{
    X0Line *o = &w->o;
    c.sync();
    if(hlim>HYST_DEPTH_MAX) hlim = HYST_DEPTH_MAX;
    if(slim>SHIFT_MAX) slim = SHIFT_MAX;
    const int n = (hlim+1)*(slim+1);
    if(o->sweeped||(n<=2))
    {
        if(c.tid==0) x0_read_pcm(px, g, mode, part, o, hlim, slim);
        c.sync();
        return;
    }
    const X0Line entry = *o;
    X0Line mine = entry; int mine_q = -1;            // the last candidate this thread filled: the winner is not filled a second time
    for(int q=c.tid;q<n;q+=c.n)
    {
        X0Line t = entry;
        t.ppb = x0_make_ppb(t.coords);
        const bool filled = x0_fill_data_words(px, g, mode, part, &t, q/(slim+1), q%(slim+1));
        w->rp_flags[q] = (u8)((filled ? 1 : 0)|((filled&&x0_crc_ok(&t)) ? 2 : 0)|((t.forced_bad&&!entry.forced_bad) ? 4 : 0));
        mine = t; mine_q = q;
    }
    c.sync();
    if(c.tid==0)
    {
        bool found = false, forced = entry.forced_bad!=0;
        int win = 0;
        for(int h=0;(h<=hlim)&&(!found);h++)
        {
            bool invalid_hyst = false;
            for(int sidx=0;sidx<=slim;sidx++)
            {
                const u8 f = w->rp_flags[h*(slim+1)+sidx];
                if(!(f&1)) { invalid_hyst = true; break; }
                if(forced) continue;                    // nothing counts on a forced-bad line (and its bit picker does not run)
                if(f&4) { forced = true; continue; }
                if(f&2) { found = true; win = h*(slim+1)+sidx; break; }
            }
            if(invalid_hyst) break;
        }
        w->rp_win = (u8)(found ? win : 0); w->rp_forced = (u8)(((!found)&&forced) ? 1 : 0);
    }
    c.sync();
    const int target = w->rp_win;
    if(c.tid==(target%c.n))
    {
        if(mine_q!=target)
        {
            mine = entry;
            mine.ppb = x0_make_ppb(mine.coords);
            x0_fill_data_words(px, g, mode, part, &mine, target/(slim+1), target%(slim+1));
        }
        if(w->rp_forced) mine.forced_bad = 1;
        *o = mine;
    }
    c.sync();
}

// The inner (right offset) loop of searchPCM16X0Data for left offset [i], over the stored reads
// (binarizer.cpp:4640-5140).  *forced: the line is already forced bad on entry / becomes so on the way.
SDV_HDN void x0_row_vote(const X0Work *w, int i, bool *forced, X0Row *out)
{
    CrcH p0[X0L_SEARCH_STEP_CNT], p1[X0L_SEARCH_STEP_CNT], p2[X0L_SEARCH_STEP_CNT], rr[X0L_SEARCH_STEP_CNT];
    CrcH c0[MAX_COLL_CRCS+1], c1[MAX_COLL_CRCS+1], c2[MAX_COLL_CRCS+1], cr[MAX_COLL_CRCS+1];
    u8 n0 = 0, n1 = 0, n2 = 0, nr = 0;
    reset_crc_stats(p0, X0L_SEARCH_STEP_CNT); reset_crc_stats(p1, X0L_SEARCH_STEP_CNT); reset_crc_stats(p2, X0L_SEARCH_STEP_CNT);
    reset_crc_stats(rr, X0L_SEARCH_STEP_CNT);
    reset_crc_stats(c0, MAX_COLL_CRCS); reset_crc_stats(c1, MAX_COLL_CRCS); reset_crc_stats(c2, MAX_COLL_CRCS); reset_crc_stats(cr, MAX_COLL_CRCS);
    bool lock_right = false, lock_min = false, lock_left = false;
    u8 step_min = 0, step_max = X0L_SEARCH_STEP_CNT, right_ofs = 0xFF;
    int last_j = 0;
    out->saw_coll = 0;
    for(int j=0;j<X0L_GRID;j++)
    {
        last_j = j;
        CrcH *pp[3] = { &p0[j], &p1[j], &p2[j] };
        CrcH *cc[3] = { c0, c1, c2 };
        u8 *nn[3] = { &n0, &n1, &n2 };
        for(int part=0;part<3;part++)
        {
            CrcH r = w->grid[i][j][part];
            if(*forced) r.result = REF_BAD_CRC;
            else if(w->coll[i][j][part]) { *forced = true; out->saw_coll = 1; }
            *pp[part] = r;
            if(r.result==REF_CRC_OK)
            {
                update_crc_stats(cc[part], r, nn[part]);
                if(!lock_min) { step_min = (u8)j; lock_min = true; }
                step_max = (u8)j;
            }
        }
        const bool a0 = p0[j].result==REF_CRC_OK, a1 = p1[j].result==REF_CRC_OK, a2 = p2[j].result==REF_CRC_OK;
        if(lock_right&&(!a0)&&(!a1)&&(!a2)) break;
        if((!lock_right)&&a0&&a1&&a2) lock_right = true;
    }
    if(n0>0) { find_most_frequent_crc(c0, &n0, true); invalidate_non_frequent(p0, 0, X0L_SEARCH_STEP_CNT-1, n0, c0[0].crc); }
    if(n1>0) { find_most_frequent_crc(c1, &n1, true); invalidate_non_frequent(p1, 0, X0L_SEARCH_STEP_CNT-1, n1, c1[0].crc); }
    if(n2>0) { find_most_frequent_crc(c2, &n2, true); invalidate_non_frequent(p2, 0, X0L_SEARCH_STEP_CNT-1, n2, c2[0].crc); }
    if(step_max>=X0L_SEARCH_STEP_CNT) step_max = X0L_SEARCH_STEP_CNT-1;
    for(int j=step_min;j<=step_max;j++)
    {
        u8 valid = 0;
        CrcH &d = rr[j];
        if(p1[j].result==REF_CRC_OK)
        {
            valid++;
            d.result = REF_CRC_OK; d.crc = X0L_CRC_SILENT; d.hyst = p1[j].hyst; d.shift = p1[j].shift; d.start = p1[j].start; d.stop = p1[j].stop;
            if(p2[j].result==REF_CRC_OK) { valid++; d.hyst = (u8)(d.hyst+p2[j].hyst); if(p2[j].shift>d.shift) d.shift = p2[j].shift; }
            else d.hyst = (u8)(d.hyst+HYST_DEPTH_SAFE);
            if(p0[j].result==REF_CRC_OK) { valid++; d.hyst = (u8)(d.hyst+p0[j].hyst); if(p0[j].shift>d.shift) d.shift = p0[j].shift; }
            else d.hyst = (u8)(d.hyst+HYST_DEPTH_SAFE);
            if(d.hyst>0x0F) d.hyst = 0x0F;
            update_crc_stats(cr, d, &nr);
        }
        else if((p0[j].result==REF_CRC_OK)&&(p2[j].result==REF_CRC_OK))
        {
            valid = 2;
            d.result = REF_CRC_OK; d.crc = X0L_CRC_SILENT; d.hyst = p2[j].hyst; d.shift = p2[j].shift; d.start = p2[j].start; d.stop = p2[j].stop;
            if(p0[j].hyst>d.hyst) { d.hyst = p0[j].hyst; d.shift = p0[j].shift; }
            else if(p0[j].hyst==d.hyst) { if(p0[j].shift>d.shift) d.shift = p0[j].shift; }
            d.hyst = (u8)(d.hyst+HYST_DEPTH_SAFE);
            if(d.hyst>0x0F) d.hyst = 0x0F;
            update_crc_stats(cr, d, &nr);
        }
        else d.result = REF_BAD_CRC;
        if(valid==3) lock_left = true;
    }
    if(nr>0) if(pick_level_by_stats(rr, &right_ofs, step_min, step_max, REF_CRC_OK, 0x0F, SHIFT_MAX)!=SPAN_OK) nr = 0;
    if(nr==0)
    {   // second chance: a single valid part
        reset_crc_stats(rr, X0L_SEARCH_STEP_CNT);
        reset_crc_stats(cr, MAX_COLL_CRCS); nr = 0;
        for(int j=step_min;j<=step_max;j++)
        {
            CrcH &d = rr[j];
            if(p2[j].result==REF_CRC_OK)
            {
                d.result = REF_CRC_OK; d.crc = X0L_CRC_SILENT; d.hyst = p2[j].hyst; d.shift = p2[j].shift; d.start = p2[j].start; d.stop = p2[j].stop;
                d.hyst = (u8)(d.hyst+HYST_DEPTH_MAX); if(d.hyst>0x0F) d.hyst = 0x0F;
                update_crc_stats(cr, d, &nr);
            }
            else if(p0[j].result==REF_CRC_OK)
            {
                d.result = REF_CRC_OK; d.crc = X0L_CRC_SILENT; d.hyst = p0[j].hyst; d.shift = p0[j].shift; d.start = p0[j].start; d.stop = p0[j].stop;
                d.hyst = (u8)(d.hyst+2*HYST_DEPTH_SAFE); if(d.hyst>0x0F) d.hyst = 0x0F;
                update_crc_stats(cr, d, &nr);
            }
            else d.result = REF_BAD_CRC;
        }
        if(nr>0) if(pick_level_by_stats(rr, &right_ofs, step_min, step_max, REF_CRC_OK, 0x0F, SHIFT_MAX)!=SPAN_OK) nr = 0;
    }
    out->valid = (nr>0) ? 1 : 0;
    out->lock_left = lock_left ? 1 : 0;
    out->last_j = (u8)last_j;
    out->parts_ok = 0;
    reset_crc_stats(&out->best, 1);
    if(nr>0)
    {
        out->best = rr[right_ofs];
        out->parts_ok = (u8)((p0[right_ofs].result==REF_CRC_OK)+(p1[right_ofs].result==REF_CRC_OK)+(p2[right_ofs].result==REF_CRC_OK));
    }
    out->pad[0] = out->pad[1] = out->pad[2] = 0;
}

// ------------------------------------------------------------------------------------------------ the grid search, bit-sliced
// As p1_search_fills_cta: the 21 x 21 grid one pixel apart, three pixel shifts, three parts per point -- 3969 fills of 64 bit
// cells in the reference -- are the 23 lanes of 41 anti-diagonals: one pass over the 193 bit cells of the line per diagonal, the
// same-level rule and three CRC-16s (one per part, state reset at the part's first cell) as lane-parallel logic.
SDV_HD void x0_search_fills_cta(const Cta &c, X0Work *w, const u8 *px, const Geom &g, int ls, int re)
{
    const int level = w->o.ref;
    const int last_px = g.W-2;
    for(int wd=c.tid;wd<P1F_WORDS;wd+=c.n)
    {
        u32 gb = 0, eb = 0;
        for(int k=0;k<32;k++)
        {
            int pidx = wd*32+k-P1F_PAD;
            if(pidx<0) pidx = 0; else if(pidx>last_px) pidx = last_px;
            const int v = px[pidx];
            if(v>level) gb |= 1u<<k;
            if(v==level) eb |= 1u<<k;
        }
        w->f_gbits[wd] = gb; w->f_ebits[wd] = eb;
    }
    c.sync();
    const u32 lanes = (1u<<X0F_LANES)-1u;
    for(int k=c.tid;k<X0F_DIAGS;k+=c.n)
    {
        Coord cc; cc.start = (i16)ls; cc.stop = (i16)(re-k);
        const Ppb pp = x0_make_ppb(cc);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for(int part=0;part<3;part++)
        {
            const int bit0 = (part==0) ? 0 : ((part==1) ? X0L_PART_BITS : (2*X0L_PART_BITS+1));
            u32 s[16];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
            for(int j=0;j<16;j++) s[j] = 0xFFFFFFFFu;
            u32 prev = 0, mismatch = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
            for(int b=0;b<X0L_PART_BITS;b++)
            {
                const int e0 = (int)((((u32)(bit0+b)*pp.psm)+pp.half)/INT_CALC_MULT)+ls-1+P1F_PAD;
                const u32 one = p1f_window(w->f_gbits, e0)|(p1f_window(w->f_ebits, e0)&prev);
                prev = one;
                if(b<48)
                {   // logical CRC bit j lives in s[(j - b) & 15]
                    const int hi = (15-b)&15;
                    const u32 fb = s[hi]^one;
                    s[hi] = fb; s[(4-b)&15] ^= fb; s[(11-b)&15] ^= fb;
                    if((part==0)&&(b<4)) w->f_top[k][b] = one;
                }
                else
                {
                    const int q = b-48, j = 15-q;
                    mismatch |= s[(j-48)&15]^one;
                    w->f_read[k][part][q] = one;
                }
            }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
            for(int j=0;j<16;j++) if(part!=1) w->f_calc[k][part>>1][j] = s[(j-48)&15];
            w->f_valid[k][part] = (~mismatch)&lanes;
        }
    }
    c.sync();
}
// One (grid point, part) from the lane bits: what x0_read_pcm(hysteresis limit 0, shift limit slim) leaves of the sub-line.
SDV_HD CrcH x0_search_point(const X0Work *w, const X0Line *o, int mode, int slim, int i, int j, int part, int ls, int re, int pixel_stop, int scan_end, bool entry_forced, u8 *coll)
{
    const int k = i+j;
    X0Line t = *o;
    t.coords.start = (i16)(ls+i); t.coords.stop = (i16)(re-j);
    t.ppb = x0_make_ppb(t.coords);
    int cnt = 0;
    if(part!=X0L_MIDDLE)
    {   // cells cut off at the line edge (x0_pick_cut_bits, first part)
        const bool left = (part==X0L_LEFT);
        const int half = (x0_get_ppb(&t)+1)/2;
        int max_cut = left ? P1_LEFT_BIT_PICK : P1_RIGHT_BIT_PICK; if(mode==SDV_MODE_DRAFT) max_cut = max_cut/2;
        int first = left ? 0 : scan_end;
        for(int idx=0;idx<max_cut;idx++)
        {
            const int cur = p1_pixel_of_bit(t.ppb, left ? idx : (X0L_BITS-1-idx), 0, pixel_stop);
            if((left ? (cur-first) : (first-cur))>=half) break;
            if(idx==0) first = cur;
            cnt = idx+1;
        }
    }
    bool forced = entry_forced, found = false, picked = false;
    u16 first_crc = 0, win_crc = 0; int win_s = 0;
    for(int sidx=0;(sidx<=slim)&&(!found);sidx++)
    {
        const int m = i+pix_shift(sidx)+1;
        if(forced) continue;
        if((w->f_valid[k][part]>>m)&1u) { found = true; win_s = sidx; win_crc = p1f_lane_crc(w->f_read[k][part], m); picked = cnt>0; break; }
        if(cnt==0) continue;
        const u16 rd = p1f_lane_crc(w->f_read[k][part], m);
        u16 calc = 0;
        for(int q=15;q>=0;q--) calc = (u16)((calc<<1)|((w->f_calc[k][part>>1][q]>>m)&1u));
        const int rep = 1<<cnt;
        if(part==X0L_LEFT)
        {
            const u16 X0_CRC_BIT[4] = { 0xD420, 0x6A10, 0x3508, 0x1A84 };
            u16 base = calc;
            for(int tb=0;tb<cnt;tb++) if((w->f_top[k][tb]>>m)&1u) base ^= X0_CRC_BIT[tb];       // the computed CRC with the cut cells cleared
            bool pf = false, pc = false;
            for(int ii=0;ii<rep;ii++)
            {
                u16 target = base;
                for(int tb=0;tb<cnt;tb++) if((ii>>(cnt-1-tb))&1) target ^= X0_CRC_BIT[tb];
                if(target==rd) { if(pf) { pc = true; break; } pf = true; }
            }
            if(pc) { forced = true; continue; }
            if(pf) { found = true; win_s = sidx; win_crc = rd; picked = true; }
        }
        else
        {
            const u16 clean = (u16)(rd&(u16)~(rep-1));
            if((u16)(calc&(u16)~(rep-1))==clean) { found = true; win_s = sidx; win_crc = (u16)(clean|(calc&(u16)(rep-1))); picked = true; }
        }
    }
    if(!found) first_crc = p1f_lane_crc(w->f_read[k][part], i+1);
    CrcH r;
    r.crc = found ? win_crc : first_crc; r.hyst = 0; r.shift = (u8)(found ? win_s : 0); r.start = t.coords.start; r.stop = t.coords.stop; r.pad = 0;
    if(found&&picked) r.hyst = (u8)((part==X0L_LEFT) ? 2 : 3);
    r.result = found ? REF_CRC_OK : REF_BAD_CRC;
    *coll = (forced&&(!entry_forced)) ? 1 : 0;
    return r;
}

// Binarizer::searchPCM16X0Data (binarizer.cpp:4514-5271).  Result in w->o / w->search_ok.
SDV_HD void x0_search_data_cta(const Cta &c, X0Work *w, const u8 *px, const Geom &g, int mode, Coord data_loc_in)
{
    X0Line *o = &w->o;
    c.sync();
    if(c.tid==0)
    {
        Coord data_loc = data_loc_in;
        i16 step = 1, ls = 0, le = 0, rs = 0, re = 0;
        int guard = 2;
        while(guard>0)
        {
            o->ppb = x0_make_ppb(data_loc);
            u16 scan_step = (u16)x0_get_ppb(o);
            if(scan_step>=X0L_SEARCH_STEP_DIV) scan_step = scan_step/X0L_SEARCH_STEP_DIV; else scan_step = 1;
            u16 span = (u16)(scan_step*X0L_SEARCH_MAX_OFS);
            step = (i16)scan_step;
            ls = (i16)(data_loc.start-span); le = (i16)(data_loc.start+span);
            rs = (i16)(data_loc.stop-span); re = (i16)(data_loc.stop+span);
            const int s0 = 0, s1 = g.scan_end;
            if(((ls<s0)&&(le<s0))||((ls>s0)&&(le>s0))||((rs<s1)&&(re<s1))||((rs>s1)&&(re>s1))) { data_loc.start = 0; data_loc.stop = (i16)g.scan_end; }
            else break;
            guard--;
        }
        w->s_left_start = ls; w->s_right_stop = re; w->s_step = step; w->s_data_loc = data_loc;
        w->s_any_coll = 0;
        w->s_next = 0;
    }
    c.sync();
    const int slim = ((mode==SDV_MODE_NORMAL)||(mode==SDV_MODE_INSANE)) ? SHIFT_SAFE : 0;
    const int ls = w->s_left_start, re = w->s_right_stop, step = w->s_step;
    const bool entry_forced = o->forced_bad!=0;
    // ---- every (left offset, right offset, part) read: from the bit-sliced fills where the grid allows it (see p1_search_data_cta),
    // else on its own thread
    const u8 lev_lo = get_low_level(o->ref, 0), lev_hi = get_high_level(o->ref, 0);
    const bool fast = (step==1)&&(lev_lo==lev_hi)&&(lev_lo==o->ref)&&(lev_lo>o->black)&&(lev_hi<o->white)&&(!o->sweeped)&&(ls>=-(P1F_PAD-2))&&(re<=g.W+P1F_PAD);
    if(fast)
    {
#if defined(SDV_EMU_COUNTERS)
        g_emu_counters[2]++;
#endif
        x0_search_fills_cta(c, w, px, g, ls, re);
        for(int q=c.tid;q<X0L_GRID*X0L_GRID*3;q+=c.n)
        {
            const int pt = q/3, part = q-3*pt;
            const int i = pt/X0L_GRID, j = pt-i*X0L_GRID;
            u8 cl = 0;
            w->grid[i][j][part] = x0_search_point(w, o, mode, slim, i, j, part, ls, re, g.W-1, g.scan_end, entry_forced, &cl);
            w->coll[i][j][part] = cl;
            if(cl) w->s_any_coll = 1;
        }
    }
    else for(;;)
    {
#if defined(SDV_EMU_COUNTERS)
        if(c.tid==0) g_emu_counters[3]++;
#endif
        const int q = grab_next(&w->s_next);
        if(q>=X0L_GRID*X0L_GRID*3) break;
        const int pt = q/3, part = q-3*pt;
        const int i = pt/X0L_GRID, j = pt-i*X0L_GRID;
        X0Line t = *o;
        t.coords.start = (i16)(ls+i*step); t.coords.stop = (i16)(re-j*step);
        x0_read_pcm(px, g, mode, part, &t, 0, slim);
        CrcH r;
        r.crc = t.words[3]; r.hyst = t.hyst; r.shift = t.shift; r.start = t.coords.start; r.stop = t.coords.stop; r.pad = 0;
        if((part==X0L_LEFT)&&t.picked_left) { r.hyst = (u8)(r.hyst+2); if(r.hyst>0x0F) r.hyst = 0x0F; }
        if((part==X0L_RIGHT)&&t.picked_right) { r.hyst = (u8)(r.hyst+3); if(r.hyst>0x0F) r.hyst = 0x0F; }
        r.result = x0_crc_ok(&t) ? REF_CRC_OK : REF_BAD_CRC;
        w->grid[i][j][part] = r;
        const u8 cl = (t.forced_bad&&(!entry_forced)) ? 1 : 0;
        w->coll[i][j][part] = cl;
        if(cl) w->s_any_coll = 1;
    }
    c.sync();
    // ---- one grid row per thread (the line is not forced bad on the way unless a collision shows up)
    const bool any_coll = (w->s_any_coll!=0)||entry_forced;
    if(!any_coll)
    {
        for(int i=c.tid;i<X0L_GRID;i+=c.n) { bool forced = false; x0_row_vote(w, i, &forced, &w->rows[i]); }
    }
    c.sync();
    // ---- the outer (left offset) loop
    if(c.tid==0)
    {
        CrcH stats[MAX_COLL_CRCS+1];
        u8 cnt = 0, ofs = 0xFF;
        reset_crc_stats(stats, MAX_COLL_CRCS);
        reset_crc_stats(w->left_res, X0L_SEARCH_STEP_CNT);
        bool lock_left = false, forced = entry_forced, forced_before_last = entry_forced;
        int last_i = 0;
        for(int i=0;i<X0L_GRID;i++)
        {
            last_i = i;
            forced_before_last = forced;
            if(any_coll) x0_row_vote(w, i, &forced, &w->rows[i]);       // sequential: the forced-bad state carries from row to row
            const X0Row &r = w->rows[i];
            if(r.lock_left) lock_left = true;
            if(r.valid)
            {
                w->left_res[i] = r.best;
                w->left_res[i].result = REF_CRC_OK;
                update_crc_stats(stats, r.best, &cnt);
                if(lock_left&&(r.parts_ok<2)) break;
            }
        }
        if(cnt>0)
        {
            find_most_frequent_crc(stats, &cnt, false);
            invalidate_non_frequent(w->left_res, 0, X0L_SEARCH_STEP_CNT-1, cnt, stats[0].crc);
        }
        if(cnt>0) if(pick_level_by_stats(w->left_res, &ofs, 0, X0L_SEARCH_STEP_CNT-1, REF_CRC_OK, 0x0F, SHIFT_MAX)!=SPAN_OK) cnt = 0;
        // the line object as the last visited grid point left it (its three reads, in order, on the running line)
        {
            X0Line t = *o;
            t.forced_bad = forced_before_last ? 1 : t.forced_bad;
            if(any_coll&&(!forced_before_last))
            {   // the collision may sit earlier in this very row: replay the row's reads up to the last column
                for(int j=0;j<w->rows[last_i].last_j;j++)
                    for(int part=0;part<3;part++) if(w->coll[last_i][j][part]) t.forced_bad = 1;
            }
            t.coords.start = (i16)(ls+last_i*step); t.coords.stop = (i16)(re-w->rows[last_i].last_j*step);
            if(fast&&(cnt>0))
            {   // The search found coordinates: the caller reads the sub-line again with them (coords_set), which rewrites every field
                // the three reads would leave -- except the forced-bad state, and that is known from the collision flags.
                for(int part=0;part<3;part++) if((!t.forced_bad)&&w->coll[last_i][w->rows[last_i].last_j][part]) t.forced_bad = 1;
            }
            else for(int part=0;part<3;part++) x0_read_pcm(px, g, mode, part, &t, 0, slim);
            w->last = t;
        }
        *o = w->last;
        if(cnt>0)
        {
            o->coords.start = w->left_res[ofs].start; o->coords.stop = w->left_res[ofs].stop;
            o->coords_set = 1; o->coord_sweeped = 1;
            w->search_ok = 1;
        }
        else
        {
            o->coords = w->s_data_loc;
            o->coord_sweeped = 0;
            w->search_ok = 0;
        }
    }
    c.sync();
}

// Binarizer::findPCM16X0Coordinates (binarizer.cpp:5819-6042); runs once per video line (VideoLine::scan_done).
SDV_HD void x0_find_coordinates_cta(const Cta &c, X0Work *w, const u8 *px, const Geom &g, int mode, Coord history)
{
    c.sync();
    const bool done = w->scan_done!=0;
    c.sync();
    if(done) return;
    Coord dc = history;
    if(!coord_valid(history))
    {
        const X0Line *o = &w->o;
        const int margin = (int)((u16)g.scan_end/40);
        const u8 ref = o->ref;
        dc.start = 0;
        bool state = px[0]>ref;
        for(int pixel=0;pixel<margin;pixel++)
        {
            if(!state) { if(px[pixel]>ref) { dc.start = (i16)(pixel-1); break; } }
            else { if(px[pixel]<ref) { dc.start = (i16)(pixel-1); break; } }
        }
        dc.stop = (i16)g.scan_end;
        state = px[g.scan_end]>ref;
        for(int pixel=g.scan_end;pixel>((int)g.scan_end-margin);pixel--)
        {
            if(!state) { if(px[pixel]>ref) { dc.stop = (i16)(pixel+1); break; } }
            else { if(px[pixel]<ref) { dc.stop = (i16)(pixel+1); break; } }
        }
    }
    x0_search_data_cta(c, w, px, g, mode, dc);
    if(c.tid==0) w->scan_done = 1;
    c.sync();
}

// Binarizer::findBlackWhite + findPCM16X0BW (binarizer.cpp:2603-2681): three windows, one inside each part.
SDV_HD void x0_find_black_white_cta(const Cta &c, X0Work *w, const u8 *px, const Geom &g)
{
    u32 *sprd = w->sprd;
    hist_clear(c, sprd);
    {
        const u16 span = g.scan_end;
        const u16 t = (u16)(span/8);
        u16 from = (u16)(span/5);
        hist_add(c, sprd, px, from, (u16)(from+t));
        from = (u16)(t*4+t/2);
        hist_add(c, sprd, px, from, (u16)(from+t));
        const u16 lim = (u16)(g.scan_end-span/64);
        hist_add(c, sprd, px, (u16)(lim-t), lim);
    }
    if(c.tid==0)
    {
        u8 bl, wh, st;
        bw_pick_levels(sprd, false, &bl, &wh, &st);
        w->was_bw_scanned = 1;
        w->o.black = bl; w->o.white = wh; w->o.bw_set = st;
    }
    c.sync();
}

// Binarizer::calcRefLevelBySweep + sweepRefLevel for one part of a PCM-16x0 line (binarizer.cpp:3551-4120), MODE_INSANE
// only; see p1_sweep_cta.  Every level resets VideoLine::scan_done and runs the full three-part grid search.
SDV_HD void x0_sweep_cta(const Cta &c, X0Work *w, const BinState *b, int part, const u8 *px, const Geom &g)
{
    X0Line *o = &w->o;
    c.sync();
    if(c.tid==0)
    {
        u8 lo = (u8)(o->black+1), hi = (u8)(o->white-1);
        if(MIN_REF_LVL>lo) lo = MIN_REF_LVL;
        if(MAX_REF_LVL<hi) hi = MAX_REF_LVL;
        w->sweep_low = lo; w->sweep_high = hi;
        w->hlim = 0; w->slim = SHIFT_SAFE;
        w->sweep_save = *o;
        x0_clear(&w->sweep_d);
    }
    for(int i=c.tid;i<256;i+=c.n) reset_crc_stats(&w->sw[i], 1);
    c.sync();
    const int lo = w->sweep_low, hi = w->sweep_high;
    for(int ref=hi;ref>=lo;ref--)
    {
        c.sync();
        if(c.tid==0)
        {
            X0Line *d = &w->sweep_d;
            // PCMLine::clear through the base pointer: words, control bit, part, picked-bit counts survive
            d->black = d->white = d->ref_low = d->ref = d->ref_high = 0;
            d->coords = coord_none();
            d->hyst = d->shift = 0;
            d->sweeped = d->coord_sweeped = d->by_ext = 0;
            d->calc_crc = 0;
            d->bw_set = d->coords_set = d->forced_bad = 0;
            d->service = 0;
            d->ppb.psm = INT_CALC_MULT; d->ppb.half = INT_CALC_MULT/2; d->ppb.ofs = 0;
            d->black = (u8)lo; d->white = (u8)hi; d->ref = (u8)ref;
            *o = *d;
            w->search_ok = x0_crc_ok(o) ? 2 : 0;
            if(w->search_ok!=2) w->scan_done = 0;         // VideoLine::scan_done is reset only where the level is searched (binarizer.cpp:3704)
        }
        c.sync();
        const bool skip = (w->search_ok==2);
        c.sync();
        if(!skip)
        {
            x0_find_coordinates_cta(c, w, px, g, b->mode, b->def_coord);
            if(c.tid==0) { if(o->coords_set) x0_read_pcm(px, g, b->mode, part, o, w->hlim, w->slim); }
        }
        if(c.tid==0)
        {
            if(o->picked_right) o->hyst = (u8)(o->hyst+HYST_DEPTH_MAX+2);
            else if(o->picked_left) o->hyst = (u8)(o->hyst+HYST_DEPTH_MAX+1);
            if(o->hyst>0x0F) o->hyst = 0x0F;
            CrcH *r = &w->sw[ref];
            if(x0_crc_ok(o)&&coord_valid(o->coords)) { r->result = REF_CRC_OK; r->start = o->coords.start; r->stop = o->coords.stop; r->hyst = o->hyst; r->shift = o->shift; r->crc = o->calc_crc; }
            else if(o->coords_set) { r->result = REF_BAD_CRC; r->start = o->coords.start; r->stop = o->coords.stop; r->hyst = o->hyst; r->shift = o->shift; r->crc = o->calc_crc; }
            w->sweep_d = *o;
        }
        c.sync();
    }
    if(c.tid==0)
    {
        *o = w->sweep_save;
        CrcH *sw = w->sw;
        CrcH stats[MAX_COLL_CRCS+1];
        u8 cnt = 0, span = SPAN_NOT_FOUND;
        const u8 fast_ref = pick_center_ref(o->black, o->white);
        reset_crc_stats(stats, MAX_COLL_CRCS+1);
        stats[0].hyst = 0; stats[0].shift = 0;
        for(u8 lvl=(u8)(o->white-1);lvl>o->black;lvl--) if(sw[lvl].result==REF_CRC_OK) update_crc_stats(stats, sw[lvl], &cnt);
        if(cnt>0)
        {
            find_most_frequent_crc(stats, &cnt, true);
            invalidate_non_frequent(sw, (u8)(o->black+1), (u8)(o->white-1), cnt, stats[0].crc);
            if(cnt>0)
            {
                if(stats[0].result<FINE_MIN_VALID_CRCS) span = SPAN_TOO_NARROW;
                else span = pick_level_by_stats(sw, &o->ref, (u8)(o->black+1), (u8)(o->white-1), REF_CRC_OK, 0x0F, SHIFT_MAX);
            }
        }
        if(span==SPAN_OK)
        {
            const CrcH t = sw[o->ref];
            o->sweeped = 1;
            if(t.stop>t.start) { o->coords.start = t.start; o->coords.stop = t.stop; }
            o->coords_set = 1;
            w->hlim = (t.hyst>HYST_DEPTH_MAX) ? HYST_DEPTH_MAX : t.hyst;
            w->slim = t.shift;
        }
        else
        {
            if(span==SPAN_TOO_NARROW)
            {
                span = pick_level_by_stats_opt(sw, &o->ref, (u8)(o->black+1), (u8)(o->white-1), REF_CRC_OK, w->hlim, w->slim);
                o->forced_bad = 1;
            }
            else span = pick_level_by_stats(sw, &o->ref, (u8)(o->black+1), (u8)(o->white-1), REF_NO_PCM, 0xFF, 0xFF);
            if(span==SPAN_OK)
            {
                const CrcH t = sw[o->ref];
                if(t.stop>t.start) { o->coords.start = t.start; o->coords.stop = t.stop; }
                o->coords_set = 1;
            }
            else if(bin_ref_preset(b))
            {
                o->ref = b->def_ref;
                if(coord_valid(b->def_coord)) o->coords = b->def_coord;
            }
            else
            {
                o->ref = fast_ref;
                if(coord_valid(b->def_coord)) o->coords = b->def_coord;
                else { o->coords.start = 0; o->coords.stop = (i16)g.scan_end; }
            }
            w->hlim = HYST_DEPTH_MIN; w->slim = SHIFT_MIN;
        }
    }
    c.sync();
}

// Binarizer::processLine for one part of a PCM-16x0 line (binarizer.cpp:443-1724).
// w->scan_done carries VideoLine::scan_done between the three parts of a line (set it before the first part).
SDV_HD void x0_process_line_cta(const Cta &c, X0Work *w, const BinState *b, int part, bool do_coord_search, const u8 *px, const Geom &g)
{
    X0Line *o = &w->o;
    c.sync();
    if(c.tid==0)
    {
        x0_clear(o);
        o->line_part = (u8)part;
        o->coords.start = 0; o->coords.stop = (i16)g.scan_end;
        w->proc_state = STG_REF_FIND;
        w->was_bw_scanned = 0;
        if(bin_bw_preset(b)) { o->black = b->def_black; o->white = b->def_white; o->bw_set = 1; }
        if(bin_ref_preset(b)) w->proc_state = coord_valid(b->def_coord) ? STG_INPUT_ALL : STG_INPUT_LEVEL;
        w->hlim = b->max_hyst; w->slim = b->max_shift;
        w->stage_count = 0;
        w->do_coord_search = do_coord_search ? 1 : 0;
        w->do_sweep = 0;
    }
    c.sync();
    for(;;)
    {
        c.sync();
        if(c.tid==0) w->stage_count++;
        const int st = w->proc_state;
        c.sync();
        if(st==STG_INPUT_ALL)
        {
            const bool need_bw = !o->bw_set;
            c.sync();
            if(need_bw) x0_find_black_white_cta(c, w, px, g);
            if(c.tid==0)
            {
                o->coords = b->def_coord;
                o->ref = b->def_ref;
                o->sweeped = 0;
                w->rp_go = 0;
                if(!o->bw_set) w->proc_state = STG_NO_GOOD;
                else if((b->def_ref>=o->white)||(b->def_ref<=o->black)) w->proc_state = STG_REF_FIND;
                else w->rp_go = 1;
            }
            c.sync();
            const bool go_read = w->rp_go!=0;
            c.sync();
            if(go_read)
            {
                x0_read_pcm_cta(c, w, px, g, b->mode, part, w->hlim, w->slim);
                if(c.tid==0)
                {
                    if(x0_crc_ok(o)) { o->by_ext = 1; w->proc_state = STG_DATA_OK; }
                    else w->proc_state = STG_REF_FIND;
                }
            }
        }
        else if(st==STG_INPUT_LEVEL)
        {
            const bool need_bw = !w->was_bw_scanned;
            c.sync();
            if(need_bw) x0_find_black_white_cta(c, w, px, g);
            if(c.tid==0)
            {
                o->coords.start = 0; o->coords.stop = (i16)g.scan_end;
                o->ref = b->def_ref;
                o->sweeped = 0;
                w->proc_state = o->bw_set ? STG_REF_FIND : STG_NO_GOOD;
            }
        }
        else if(st==STG_REF_FIND)
        {
            const bool need_bw = !w->was_bw_scanned;
            c.sync();
            if(need_bw) x0_find_black_white_cta(c, w, px, g);
            if(c.tid==0)
            {
                if(!o->bw_set) w->proc_state = STG_NO_GOOD;
                else if(b->mode==SDV_MODE_INSANE) { w->do_sweep = 1; w->proc_state = STG_REF_SWEEP_RUN; }
                else
                {
                    w->do_sweep = 0;
                    w->hlim = HYST_DEPTH_SAFE; w->slim = SHIFT_MIN;
                    o->ref = pick_center_ref(o->black, o->white);
                    if(coord_valid(b->def_coord)) o->coords = b->def_coord;
                    else { o->coords.start = 0; o->coords.stop = (i16)g.scan_end; }
                    w->proc_state = 0xFF;
                }
            }
            c.sync();
            const bool go = (w->proc_state==0xFF);
            c.sync();
            if(go)
            {
                if(w->do_coord_search&&FINE_EN_COORD_SEARCH) x0_find_coordinates_cta(c, w, px, g, b->mode, b->def_coord);
                if(c.tid==0)
                {
                    if(!o->coords_set)
                    {
                        if(b->mode==SDV_MODE_DRAFT) { w->hlim = 2; w->slim = SHIFT_MIN; }
                        else { w->hlim = HYST_DEPTH_SAFE; w->slim = SHIFT_SAFE; }
                    }
                    else { w->hlim = b->max_hyst; w->slim = SHIFT_SAFE; }
                    w->proc_state = STG_READ_PCM;
                }
            }
        }
        else if(st==STG_REF_SWEEP_RUN)
        {
            x0_sweep_cta(c, w, b, part, px, g);
            if(c.tid==0) w->proc_state = STG_READ_PCM;
        }
        else if(st==STG_READ_PCM)
        {
            const bool rd1 = o->coords_set!=0;
            c.sync();
            if(rd1) x0_read_pcm_cta(c, w, px, g, b->mode, part, w->hlim, w->slim);
            if(c.tid==0)
            {
                w->rp_go = 0;
                if(x0_crc_ok(o)) w->proc_state = STG_DATA_OK;
                else
                {
                    w->proc_state = STG_NO_GOOD;
                    if(coord_valid(b->def_coord)&&(!w->do_sweep)&&(!o->forced_bad)&&(!o->coords_set))
                        if(!coord_eq(o->coords, b->def_coord)) { o->coords = b->def_coord; w->rp_go = 1; }
                }
            }
            c.sync();
            const bool rd2 = w->rp_go!=0;
            c.sync();
            if(rd2)
            {
                x0_read_pcm_cta(c, w, px, g, b->mode, part, w->hlim, w->slim);
                if(c.tid==0) { if(x0_crc_ok(o)) w->proc_state = STG_DATA_OK; }
            }
        }
        else if(st==STG_DATA_OK)
        {
            if(c.tid==0)
            {
                if(o->forced_bad) w->proc_state = STG_NO_GOOD;
                else { o->coords_set = 1; w->proc_state = 0xFD; }
            }
            c.sync();
            const bool done = (w->proc_state==0xFD);
            c.sync();
            if(done) break;
        }
        else
        {   // STG_NO_GOOD
            if(c.tid==0) { if(x0_crc_ok(o)) x0_set_invalid_crc(o); }
            break;
        }
        c.sync();
        const bool overrun = w->stage_count>STG_MAX;
        c.sync();
        if(overrun) break;
    }
    c.sync();
}

}   // namespace sdv
