// pcm1_line.cuh -- PCM-1 line decode (the Binarizer operator for PCM1Line) as cooperative integer code.
//
// PCM-1 lines carry no START/STOP markers: the data coordinates are found by brute force, a 25 x 25 grid of
// (start, stop) offsets around a rough guess, each grid point one readPCMdata() with the bit picker forced
// (Binarizer::searchPCM1Data, binarizer.cpp:4123-4511).  Here every grid point is decoded by its own thread on a
// private copy of the line; the reference's per-row and per-column CRC votes are then replayed over the grid.
// The one sequential dependency of the reference loop -- a bit-picker CRC collision forces the line bad for every
// later grid point -- is restored afterwards from per-point collision flags.  With Cta{0,1} the same code is a
// sequential program (tests/hostemu).
#pragma once
#include "stc007_line.cuh"

namespace sdv {

enum { P1_BITS = 94, P1_WORD_BITS = 13, P1L_WORDS = 7, P1_CRC_SILENT = 0xECBF, P1_BIT_RANGE = 1<<12 };      // pcm1line.h:66-101
enum { P1_SEARCH_STEP_DIV = 4, P1_SEARCH_MAX_OFS = 12, P1_SEARCH_STEP_CNT = (P1_SEARCH_MAX_OFS+1)*2,       // binarizer.h:254-256
       P1_GRID = 2*P1_SEARCH_MAX_OFS+1 };
// (P1_LEFT_BIT_PICK / P1_RIGHT_BIT_PICK: bin_preset_t::left_bit_pick / right_bit_pick, sdv_common.cuh)

enum { P1F_PAD = 64, P1F_WORDS = (SDV_MAX_W+192)/32+2, P1F_DIAGS = 2*(P1_GRID-1)+1, P1F_LANES = P1_GRID+2, P1F_KEEP = 13+16 };

// PCM1Line + PCMLine payload (pcmline.h:132-160, pcm1line.h:103-110); bit positions are recomputed from [ppb].
struct P1Line
{
    u16 words[P1L_WORDS];                // L2 R2 L4 R4 L6 R6 (13 bit) + CRCC as read
    u16 calc_crc;
    Coord coords;
    u8 black, white, ref_low, ref, ref_high, hyst, shift, service;
    u8 picked_left, picked_right;
    u8 sweeped, coord_sweeped, by_ext, bw_set, coords_set, forced_bad;
    Ppb ppb;
};

SDV_HD bool p1_words_header(const u16 *w)
{   // PCM1Line::hasHeader (pcm1line.cpp:314-323)
    return (w[0]==0x0666)&&(w[1]==0x0CCC)&&(w[2]==0x1999)&&(w[3]==0x1333)&&(w[4]==0x0666)&&(w[5]==0x0CCC)&&(w[6]==0xCCCC);
}
SDV_HD u16 p1_calc_crc(const u16 *w)
{   // PCM1Line::calcCRC (pcm1line.cpp:158-166): CRC over the inverted words, result inverted
    // 78 message bits: the first 6 bit by bit, then 9 bytes
    const u32 i0 = (~(u32)w[0])&0x1FFFu, i1 = (~(u32)w[1])&0x1FFFu;
    const u64 lo = ((u64)(i1&0xFFFu)<<52)|((u64)((~(u32)w[2])&0x1FFFu)<<39)|((u64)((~(u32)w[3])&0x1FFFu)<<26)
                   |((u64)((~(u32)w[4])&0x1FFFu)<<13)|(u64)((~(u32)w[5])&0x1FFFu);
    u16 c = crc16_update(0xFFFF, (u16)(i0>>7), 6);
    c = crc16_byte(c, ((i0&0x7Fu)<<1)|(i1>>12));
    for(int k=7;k>=0;k--) c = crc16_byte(c, (u32)(lo>>(8*k))&0xFFu);
    return (u16)~c;
}
SDV_HD bool p1_crc_ok_ign(const P1Line *l) { return (l->calc_crc==l->words[6])||p1_words_header(l->words); }
SDV_HD bool p1_crc_ok(const P1Line *l) { return (!l->forced_bad)&&p1_crc_ok_ign(l); }
SDV_HD void p1_set_invalid_crc(P1Line *l) { l->words[6] = (u16)~l->calc_crc; }
SDV_HD int p1_get_ppb(const P1Line *l) { return (int)((l->ppb.psm/INT_CALC_MULT)&0xFF); }

// PCMLine::clear (pcmline.cpp:94-112), the part a service line conversion applies (setServiceLine is a base method).
SDV_HD void p1_base_clear(P1Line *l)
{
    l->black = l->white = l->ref_low = l->ref = l->ref_high = 0;
    l->coords = coord_none();
    l->hyst = l->shift = 0;
    l->sweeped = l->coord_sweeped = l->by_ext = 0;
    l->calc_crc = 0;
    l->bw_set = l->coords_set = l->forced_bad = 0;
    l->service = 0;
    l->ppb.psm = INT_CALC_MULT; l->ppb.half = INT_CALC_MULT/2; l->ppb.ofs = 0;
}
SDV_HD void p1_clear(P1Line *l)
{   // PCM1Line::clear (pcm1line.cpp:56-76)
    p1_base_clear(l);
    l->picked_left = l->picked_right = 0;
    for(int i=0;i<6;i++) l->words[i] = P1_BIT_RANGE;
    l->calc_crc = P1_CRC_SILENT;
    p1_set_invalid_crc(l);
}
SDV_HD void p1_set_serv_header(P1Line *l) { p1_base_clear(l); l->service = SDV_SRV_HEADER_LINE; }      // pcm1line.cpp:78-85

// PCMLine::setPPB / getVideoPixeBylCalc (pcmline.cpp:249-311,506-519) for the 94 bit cells of a PCM-1 line, bit offset 0.
SDV_HD Ppb p1_make_ppb(Coord c)
{
    Ppb p;
    p.psm = (u32)(c.stop-c.start);
    p.psm = (p.psm*INT_CALC_MULT+P1_BITS/2)/P1_BITS;
    p.ofs = c.start;
    p.half = (p.psm+1)/2;
    return p;
}
SDV_HD int p1_pixel_of_bit(Ppb p, int pcm_bit, int shift_px, int pixel_stop)
{
    i32 vp = (i32)(((u32)pcm_bit*p.psm)+p.half);
    vp = vp/INT_CALC_MULT;
    vp = vp+p.ofs+shift_px;
    if(vp<0) vp = 0;
    else if(vp>=pixel_stop) vp = pixel_stop-1;
    return vp;
}

// Binarizer::fillPCM1 (binarizer.cpp:7016-7131).
SDV_HDN void p1_fill(const u8 *px, int pixel_stop, Ppb ppb, int shift_stage, u8 low_ref, u8 high_ref, u16 *words /*[7]*/)
{
    bool prev_high = false;
    int sh = pix_shift(shift_stage);
    int bit = 0;
    for(int w=0;w<P1L_WORDS;w++)
    {
        int nb = (w<6) ? P1_WORD_BITS : 16;
        u32 acc = 0;
        for(int k=0;k<nb;k++, bit++)
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

// Binarizer::pickCutBitsUpPCM1 (binarizer.cpp:6116-6596): how many leading/trailing bit cells fall outside the video
// line, and -- on a bad CRC -- the brute-force search for the one patch of those bits that makes the CRC valid.
SDV_HDN void p1_pick_cut_bits(P1Line *l, int mode, int pixel_stop, int scan_end)
{
    l->picked_left = l->picked_right = 0;
    const int half = (p1_get_ppb(l)+1)/2;
    int lcnt = 0, rcnt = 0;
    int max_cut = P1_LEFT_BIT_PICK; if(mode==SDV_MODE_DRAFT) max_cut = max_cut/2;
    int first = 0;
    for(int idx=0;idx<max_cut;idx++)
    {
        int cur = p1_pixel_of_bit(l->ppb, idx, 0, pixel_stop);
        if((cur-first)>=half) break;
        if(idx==0) first = cur;
        lcnt = idx+1;
    }
    max_cut = P1_RIGHT_BIT_PICK; if(mode==SDV_MODE_DRAFT) max_cut = max_cut/2;
    first = scan_end;
    for(int idx=0;idx<max_cut;idx++)
    {
        int cur = p1_pixel_of_bit(l->ppb, P1_BITS-1-idx, 0, pixel_stop);
        if((first-cur)>=half) break;
        if(idx==0) first = cur;
        rcnt = idx+1;
    }
    if(p1_crc_ok(l)) { l->picked_left = (u8)lcnt; l->picked_right = (u8)rcnt; return; }     // forced run on a valid line: only count
    if((lcnt==0)&&(rcnt==0)) return;
    const u16 lorig = l->words[0], rorig = l->words[6];
    const int lrep = 1<<lcnt, rrep = 1<<rcnt;
    const u16 lclean = (u16)(lorig&(u16)~((lrep-1)<<(P1_WORD_BITS-lcnt)));
    const u16 rclean = (u16)(rorig&(u16)~(rrep-1));
    bool found = false, coll = false;
    u16 lfix = 0, rfix = 0;
    // The reference tries every left patch x right patch and recomputes the CRC each time.  Same outcome, no loop over CRCs:
    // the left patch changes the computed CRC linearly (message bit t of 78 contributes P1_CRC_BIT[t]), the right patch
    // replaces the low bits of the CRC that was READ -- so a left patch fits at most one right patch.  What the reference's
    // loops report depends only on how many (left, right) pairs fit: none, exactly one (that patch), or more (a collision).
    if(!l->forced_bad)
    {
        const u16 P1_CRC_BIT[4] = { 0x390D, 0x9496, 0x4A4B, 0xAD35 };
        if(lcnt>0) l->words[0] = lclean;
        const u16 base = p1_calc_crc(l->words);
        const bool hdr_mid = (l->words[1]==0x0CCC)&&(l->words[2]==0x1999)&&(l->words[3]==0x1333)&&(l->words[4]==0x0666)&&(l->words[5]==0x0CCC);
        const u16 rmask = (u16)(rrep-1);
        for(int i=0;(i<lrep)&&(!coll);i++)
        {
            u16 target = base;
            for(int t=0;t<lcnt;t++) if((i>>(lcnt-1-t))&1) target ^= P1_CRC_BIT[t];
            const u16 lpatch = (u16)(i<<(P1_WORD_BITS-lcnt));
            const u16 w0 = (lcnt>0) ? (u16)((lclean|lpatch)&0x1FFF) : lorig;
            int j1 = -1, j2 = -1;
            if((u16)(target&(u16)~rmask)==rclean) j1 = (int)(target&rmask);
            if(hdr_mid&&(w0==0x0666)&&((u16)(0xCCCC&(u16)~rmask)==rclean)) { j2 = (int)(0xCCCC&rmask); if(j2==j1) j2 = -1; }    // PCM1Line::isCRCValidIgnoreForced: a header line counts as valid
            for(int k=0;k<2;k++)
            {
                const int j = k ? j2 : j1;
                if(j<0) continue;
                if(found) { coll = true; break; }
                found = true; lfix = lpatch; rfix = (u16)j;
            }
        }
    }
    if(coll||(!found))
    {
        l->words[0] = lorig; l->words[6] = rorig;
        l->calc_crc = p1_calc_crc(l->words);
        if(coll) l->forced_bad = 1;
        return;
    }
    if(lcnt>0) l->words[0] = (u16)((lclean|lfix)&0x1FFF);
    if(rcnt>0) l->words[6] = (u16)(rclean|rfix);
    l->calc_crc = p1_calc_crc(l->words);
    l->picked_left = (u8)lcnt; l->picked_right = (u8)rcnt;
}

// Binarizer::fillDataWords for PCM-1 (binarizer.cpp:7560-7650); the bit picker is always forced (binarizer.cpp:82).
SDV_HD bool p1_fill_data_words(const u8 *px, const Geom &g, int mode, P1Line *l, int hyst, int shift)
{
    u8 low = get_low_level(l->ref, (u8)hyst), high = get_high_level(l->ref, (u8)hyst);
    l->ref_low = low; l->ref_high = high;
    if((low<=l->black)||(high>=l->white)) { p1_set_invalid_crc(l); return false; }
    l->hyst = (u8)hyst; l->shift = (u8)shift;
    p1_fill(px, g.W-1, l->ppb, shift, low, high, l->words);
    l->calc_crc = p1_calc_crc(l->words);
    p1_pick_cut_bits(l, mode, g.W-1, g.scan_end);
    return true;
}

// Binarizer::readPCMdata (binarizer.cpp:7695-8055) for a PCM-1 line: first (hysteresis, shift) with a valid CRC in
// lexicographic order, then the final fill with the winner or with (0,0).  Sequential: the line state (forced_bad set
// by a bit-picker collision) carries from one fill to the next.
SDV_HDN void p1_read_pcm(const u8 *px, const Geom &g, int mode, P1Line *l, int hlim, int slim)
{
    if(hlim>HYST_DEPTH_MAX) hlim = HYST_DEPTH_MAX;
    if(slim>SHIFT_MAX) slim = SHIFT_MAX;
    l->ppb = p1_make_ppb(l->coords);
    int win_h = 0, win_s = 0;
    if(!l->sweeped)
    {
        bool found = false;
        P1Line first;                   // the line as the (0,0) fill left it
        for(int h=0;(h<=hlim)&&(!found);h++)
        {
            bool invalid_hyst = false;
            for(int s=0;s<=slim;s++)
            {
                const bool filled = p1_fill_data_words(px, g, mode, l, h, s);
                if((h==0)&&(s==0)) first = *l;
                if(!filled) { invalid_hyst = true; break; }
                if(p1_crc_ok(l)) { found = true; win_h = h; win_s = s; break; }
            }
            if(invalid_hyst) break;
        }
        if(found&&(win_h==l->hyst)&&(win_s==l->shift)&&(!l->forced_bad)) return;      // the final fill would repeat the winning one bit for bit
        if(!found)
        {   // the final (0,0) fill repeats the first one: same samples, and a bit-picker patch cannot succeed now if it did
            // not then (it would have been the winner); only a forced-bad state picked up on the way stays
            const u8 fb = l->forced_bad;
            *l = first;
            l->forced_bad = (u8)(fb|first.forced_bad);
            return;
        }
    }
    else { win_h = hlim; win_s = slim; }
    p1_fill_data_words(px, g, mode, l, win_h, win_s);
}

// ------------------------------------------------------------------------------------------------ shared work area
struct P1Work
{
    P1Line o;                                   // the output line being built
    P1Line last;                                // line state after the last grid point of the coordinate search
    u32 sprd[256];                              // brightness histogram
    CrcH grid[P1_GRID][P1_SEARCH_STEP_CNT];     // scan_right_res of every left offset ([..][25] stays reset, as in the reference)
    CrcH left_res[P1_SEARCH_STEP_CNT];          // scan_left_res
    CrcH row_best[P1_GRID];                     // scan_right_crcs[0] of every left offset
    u8 row_valid[P1_GRID];
    u8 coll[P1_GRID*P1_GRID];                   // bit-picker collision at this grid point
    // scalars
    u8 proc_state, was_bw_scanned, hlim, slim, stage_count, do_coord_search, search_ok, pad0;
    i16 s_left_start, s_right_stop, s_step;
    Coord s_data_loc;
    int s_first_coll;
    int s_next;                                 // next grid point to hand out (dynamic distribution)
    // MODE_INSANE: reference level sweep (Binarizer::sweepRefLevel for PCM1Line)
    CrcH sw[256];
    P1Line sweep_d, sweep_save;                 // the sweep's dummy line (its words survive from level to level), the real line
    u8 do_sweep, sweep_low, sweep_high, pad1;
    // bit-sliced grid search (p1_search_fills_cta): per anti-diagonal of the grid one 32-lane word per bit cell
    u32 f_gbits[P1F_WORDS], f_ebits[P1F_WORDS]; // pixel > level / pixel == level, for pixel index -P1F_PAD .. (clamped to the line)
    u32 f_win[P1F_DIAGS][P1F_KEEP];             // bits 0..12 (first word) and 78..93 (CRCC as read) of every fill of the diagonal
    u32 f_crc[P1F_DIAGS][16];                   // computed CRC of every fill, bit-sliced
    u32 f_valid[P1F_DIAGS], f_hdrmid[P1F_DIAGS];// fills with a valid CRC (or the header pattern); fills whose words 1..5 are the header's
    u8 rp_flags[MAX_CAND+1]; u8 rp_go, rp_win, rp_forced;          // p1_read_pcm_cta: per (hysteresis, shift) candidate: bit 0 filled, 1 CRC valid, 2 collision
};

// p1_read_pcm by the whole group: the (hysteresis, shift) candidates are independent fills -- a fill rewrites everything of the line
// but the forced-bad state, which only a bit-picker collision sets -- so every candidate is tried on its own thread from the entry
// state and the reference's loop is replayed over three flag bits per candidate: stop at the first hysteresis depth whose levels
// touch black / white, the first candidate with a valid CRC wins unless a collision came before it.  The line is then filled once
// more with the winner (or with (0, 0), keeping a collision's forced-bad mark): one fill instead of up to 56 in a row.
SDV_HD void p1_read_pcm_cta(const Cta &c, P1Work *w, const u8 *px, const Geom &g, int mode, int hlim, int slim)
{
    P1Line *o = &w->o;
    c.sync();
    if(hlim>HYST_DEPTH_MAX) hlim = HYST_DEPTH_MAX;
    if(slim>SHIFT_MAX) slim = SHIFT_MAX;
    const int n = (hlim+1)*(slim+1);
    if(o->sweeped||(n<=2))
    {
        if(c.tid==0) p1_read_pcm(px, g, mode, o, hlim, slim);
        c.sync();
        return;
    }
    const P1Line entry = *o;
    P1Line mine = entry; int mine_q = -1;            // the last candidate this thread filled: the winner is not filled a second time
    for(int q=c.tid;q<n;q+=c.n)
    {
        P1Line t = entry;
        t.ppb = p1_make_ppb(t.coords);
        const bool filled = p1_fill_data_words(px, g, mode, &t, q/(slim+1), q%(slim+1));
        w->rp_flags[q] = (u8)((filled ? 1 : 0)|((filled&&p1_crc_ok(&t)) ? 2 : 0)|((t.forced_bad&&!entry.forced_bad) ? 4 : 0));
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
                if(forced) continue;
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
            mine.ppb = p1_make_ppb(mine.coords);
            p1_fill_data_words(px, g, mode, &mine, target/(slim+1), target%(slim+1));
        }
        if(w->rp_forced) mine.forced_bad = 1;
        *o = mine;
    }
    c.sync();
}

// Binarizer::findBlackWhite + findPCM1BW (binarizer.cpp:2560-2600,3116-3473).
SDV_HD void p1_find_black_white_cta(const Cta &c, P1Work *w, const u8 *px, const Geom &g)
{
    u32 *sprd = w->sprd;
    hist_clear(c, sprd);
    {
        u16 pixel_limit = g.scan_end;
        u16 search_lim = (u16)(g.scan_end-(u16)(pixel_limit/32));
        u16 from = (u16)(pixel_limit/8);
        hist_add(c, sprd, px, from, search_lim);
    }
    if(c.tid==0)
    {
        u8 bl, wh, st;
        bw_pick_levels(sprd, false, &bl, &wh, &st);     // do_ref_lvl_sweep is never set for PCM-1 below MODE_INSANE (binarizer.cpp:1105-1112)
        w->was_bw_scanned = 1;
        w->o.black = bl; w->o.white = wh; w->o.bw_set = st;
    }
    c.sync();
}

// ------------------------------------------------------------------------------------------------ the grid search, bit-sliced
// The grid visits (start, stop) = (ls + i, re - j), 25 x 25 points one pixel apart, and reads each with pixel shifts 0, +1, -1
// at hysteresis depth 0: 1875 fills of 94 bit cells + a CRC each in the reference.  A fill is determined by the cell pitch,
// i.e. by stop - start (constant along an anti-diagonal i + j), and by start + shift: the 27 fills of one anti-diagonal sample, for
// every bit cell, 27 CONSECUTIVE pixels.  So one thread takes an anti-diagonal and carries its 27 fills as the lanes of 32-bit words:
// per bit cell one window out of the "pixel above the level" bit string of the line, the same-level rule (a pixel exactly on the
// level keeps the previous bit) and the CRC-16 as lane-parallel logic (16 state words, 4 operations per message bit for all 27 fills).
// 49 threads x ~1100 operations replace 1875 x ~1500.  What the reference's per-point loop makes of the fills (first valid shift,
// bit picker on the cut-off cells, collisions) is replayed per grid point afterwards from the lane bits.
SDV_HD u32 p1f_window(const u32 *bits, int e)
{   // 32 bits starting at bit e
    const u32 lo = bits[e>>5], hi = bits[(e>>5)+1];
    const int sh = e&31;
    return sh ? ((lo>>sh)|(hi<<(32-sh))) : lo;
}
SDV_HD void p1_search_fills_cta(const Cta &c, P1Work *w, const u8 *px, const Geom &g, int ls, int re, int mode, int slim)
{
    const int level = w->o.ref;                 // hysteresis 0: low = high = the reference level (checked by the caller)
    const int last_px = g.W-2;                  // p1_fill clamps to pixel_stop - 1 = W - 2
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
    if((c.n>=P1F_DIAGS+32)&&(c.tid==c.n-1))
    {   // the line state the last grid point leaves behind (it becomes the output line): one ordinary read, done by a thread of a warp
        // that has no diagonal to carry, beside them
        P1Line t = w->o;
        t.coords.start = (i16)(ls+(P1_GRID-1)); t.coords.stop = (i16)(re-(P1_GRID-1));
        p1_read_pcm(px, g, mode, &t, 0, slim);
        w->last = t;
    }
    const u32 lanes = (1u<<P1F_LANES)-1u;
    for(int k=c.tid;k<P1F_DIAGS;k+=c.n)
    {
        Coord cc; cc.start = (i16)ls; cc.stop = (i16)(re-k);            // any point of the diagonal: the pitch depends on stop - start only
        const Ppb pp = p1_make_ppb(cc);
        u32 s[16];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for(int j=0;j<16;j++) s[j] = 0xFFFFFFFFu;
        u32 prev = 0, hdr_all = 0xFFFFFFFFu, hdr_mid = 0xFFFFFFFFu, mismatch = 0;
        const u16 hdr_words[7] = { 0x0666, 0x0CCC, 0x1999, 0x1333, 0x0666, 0x0CCC, 0xCCCC };
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for(int b=0;b<P1_BITS;b++)
        {   // (fully unrolled: the rotating index of the CRC state is a compile-time constant, the state stays in registers)
            const int wi = (b<6*P1_WORD_BITS) ? (b/P1_WORD_BITS) : 6;
            const int q = b-wi*P1_WORD_BITS;
            const int nb = (wi<6) ? P1_WORD_BITS : 16;
            const int e0 = (int)((((u32)b*pp.psm)+pp.half)/INT_CALC_MULT)+ls-1+P1F_PAD;     // lane m: start + shift = ls - 1 + m
            const u32 one = p1f_window(w->f_gbits, e0)|(p1f_window(w->f_ebits, e0)&prev);
            prev = one;
            const u32 want = ((hdr_words[wi]>>(nb-1-q))&1) ? 0xFFFFFFFFu : 0u;
            const u32 same = ~(one^want);
            hdr_all &= same;
            if((wi>=1)&&(wi<=5)) hdr_mid &= same;
            if(wi<6)
            {   // CRC over the inverted words: message bit = ~one.  Logical CRC bit j lives in s[(j - b) & 15].
                const int hi = (15-b)&15;
                const u32 fb = s[hi]^(~one);
                s[hi] = fb; s[(4-b)&15] ^= fb; s[(11-b)&15] ^= fb;
                if(wi==0) w->f_win[k][q] = one;
            }
            else
            {   // CRCC as read, MSB first: compare with the inverted CRC state (78 message bits in)
                const int j = 15-q;
                mismatch |= (~s[(j-78)&15])^one;
                w->f_win[k][13+q] = one;
            }
        }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for(int j=0;j<16;j++) w->f_crc[k][j] = ~s[(j-78)&15];
        w->f_valid[k] = ((~mismatch)|hdr_all)&lanes;
        w->f_hdrmid[k] = hdr_mid&lanes;
    }
    c.sync();
}
// The 16 bits lane m carries in 16 consecutive lane words, MSB first.
SDV_HD u16 p1f_lane_crc(const u32 *bits16, int m)
{
    u32 v = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for(int q=0;q<16;q++) v = (v<<1)|((bits16[q]>>m)&1u);
    return (u16)v;
}
// One grid point from the lane bits: what p1_read_pcm(hysteresis limit 0, shift limit slim) leaves of the line.
SDV_HD CrcH p1_search_point(const P1Work *w, const P1Line *o, int mode, int slim, int i, int j, int ls, int re, int pixel_stop, int scan_end, bool entry_forced, u8 *coll)
{
    const int k = i+j;
    P1Line t = *o;
    t.coords.start = (i16)(ls+i); t.coords.stop = (i16)(re-j);
    t.ppb = p1_make_ppb(t.coords);
    // cells cut off at the line edges (p1_pick_cut_bits, first part: depends on the point only)
    int lcnt = 0, rcnt = 0;
    {
        const int half = (p1_get_ppb(&t)+1)/2;
        int max_cut = P1_LEFT_BIT_PICK; if(mode==SDV_MODE_DRAFT) max_cut = max_cut/2;
        int first = 0;
        for(int idx=0;idx<max_cut;idx++)
        {
            const int cur = p1_pixel_of_bit(t.ppb, idx, 0, pixel_stop);
            if((cur-first)>=half) break;
            if(idx==0) first = cur;
            lcnt = idx+1;
        }
        max_cut = P1_RIGHT_BIT_PICK; if(mode==SDV_MODE_DRAFT) max_cut = max_cut/2;
        first = scan_end;
        for(int idx=0;idx<max_cut;idx++)
        {
            const int cur = p1_pixel_of_bit(t.ppb, P1_BITS-1-idx, 0, pixel_stop);
            if((first-cur)>=half) break;
            if(idx==0) first = cur;
            rcnt = idx+1;
        }
    }
    bool forced = entry_forced, found = false;
    u16 first_crc = 0, win_crc = 0; int win_s = 0; u8 pl = 0, pr = 0;
    for(int sidx=0;(sidx<=slim)&&(!found);sidx++)
    {
        const int m = i+pix_shift(sidx)+1;
        if(forced) continue;                                    // a forced-bad line: no fill counts, the picker does not run
        if((w->f_valid[k]>>m)&1u) { found = true; win_s = sidx; win_crc = p1f_lane_crc(w->f_win[k]+13, m); pl = (u8)lcnt; pr = (u8)rcnt; break; }
        if((lcnt==0)&&(rcnt==0)) continue;
        // the bit picker (p1_pick_cut_bits, second part) on this fill
        const u16 rd = p1f_lane_crc(w->f_win[k]+13, m);
        u16 w0 = 0, calc = 0;
        for(int q=0;q<13;q++) w0 = (u16)((w0<<1)|((w->f_win[k][q]>>m)&1u));
        for(int q=15;q>=0;q--) calc = (u16)((calc<<1)|((w->f_crc[k][q]>>m)&1u));
        const u16 P1_CRC_BIT[4] = { 0x390D, 0x9496, 0x4A4B, 0xAD35 };
        const int lrep = 1<<lcnt, rrep = 1<<rcnt;
        const u16 lclean = (u16)(w0&(u16)~((lrep-1)<<(P1_WORD_BITS-lcnt)));
        const u16 rclean = (u16)(rd&(u16)~(rrep-1)), rmask = (u16)(rrep-1);
        u16 base = calc;
        for(int tb=0;tb<lcnt;tb++) if((w0>>(P1_WORD_BITS-1-tb))&1) base ^= P1_CRC_BIT[tb];     // the computed CRC with the cut cells cleared
        const bool hdr_mid = ((w->f_hdrmid[k]>>m)&1u)!=0;
        bool pf = false, pc = false; u16 lfix = 0, rfix = 0;
        for(int ii=0;(ii<lrep)&&(!pc);ii++)
        {
            u16 target = base;
            for(int tb=0;tb<lcnt;tb++) if((ii>>(lcnt-1-tb))&1) target ^= P1_CRC_BIT[tb];
            const u16 lpatch = (u16)(ii<<(P1_WORD_BITS-lcnt));
            const u16 nw0 = (lcnt>0) ? (u16)((lclean|lpatch)&0x1FFF) : w0;
            int j1 = -1, j2 = -1;
            if((u16)(target&(u16)~rmask)==rclean) j1 = (int)(target&rmask);
            if(hdr_mid&&(nw0==0x0666)&&((u16)(0xCCCC&(u16)~rmask)==rclean)) { j2 = (int)(0xCCCC&rmask); if(j2==j1) j2 = -1; }
            for(int kk=0;kk<2;kk++)
            {
                const int jj = kk ? j2 : j1;
                if(jj<0) continue;
                if(pf) { pc = true; break; }
                pf = true; lfix = lpatch; rfix = (u16)jj;
            }
        }
        (void)lfix;
        if(pc) { forced = true; continue; }
        if(pf) { found = true; win_s = sidx; win_crc = (rcnt>0) ? (u16)(rclean|rfix) : rd; pl = (u8)lcnt; pr = (u8)rcnt; }
    }
    if(!found) first_crc = p1f_lane_crc(w->f_win[k]+13, i+1);      // the CRCC the shift-0 fill read (the line as the failed read leaves it)
    CrcH r;
    r.crc = found ? win_crc : first_crc; r.hyst = 0; r.shift = (u8)(found ? win_s : 0); r.start = t.coords.start; r.stop = t.coords.stop; r.pad = 0;
    if(pl&&pr) r.hyst = 0x0E; else if(pr) r.hyst = 0x0D; else if(pl) r.hyst = 0x0C;
    r.result = found ? REF_CRC_OK : REF_BAD_CRC;
    *coll = (forced&&(!entry_forced)) ? 1 : 0;
    return r;
}

// Binarizer::searchPCM1Data (binarizer.cpp:4123-4511).  Leaves the result in w->o; returns through w->search_ok.
SDV_HD void p1_search_data_cta(const Cta &c, P1Work *w, const u8 *px, const Geom &g, int mode, Coord data_loc_in)
{
    P1Line *o = &w->o;
    c.sync();
    if(c.tid==0)
    {
        Coord data_loc = data_loc_in;
        i16 step = 1, ls = 0, le = 0, rs = 0, re = 0;
        int guard = 2;
        while(guard>0)
        {
            o->ppb = p1_make_ppb(data_loc);
            u16 scan_step = (u16)p1_get_ppb(o);
            if(scan_step>=P1_SEARCH_STEP_DIV) scan_step = scan_step/P1_SEARCH_STEP_DIV; else scan_step = 1;
            u16 span = (u16)(scan_step*P1_SEARCH_MAX_OFS);
            step = (i16)scan_step;
            ls = (i16)(data_loc.start-span); le = (i16)(data_loc.start+span);
            rs = (i16)(data_loc.stop-span); re = (i16)(data_loc.stop+span);
            const int s0 = 0, s1 = g.scan_end;
            if(((ls<s0)&&(le<s0))||((ls>s0)&&(le>s0))||((rs<s1)&&(re<s1))||((rs>s1)&&(re>s1))) { data_loc.start = 0; data_loc.stop = (i16)g.scan_end; }
            else break;
            guard--;
        }
        w->s_left_start = ls; w->s_right_stop = re; w->s_step = step; w->s_data_loc = data_loc;
        w->s_first_coll = P1_GRID*P1_GRID;
        w->s_next = 0;
    }
    c.sync();
    const int slim = ((mode==SDV_MODE_NORMAL)||(mode==SDV_MODE_INSANE)) ? SHIFT_SAFE : 0;       // binarizer.cpp:4223-4243 (hysteresis 0)
    const int ls = w->s_left_start, re = w->s_right_stop, step = w->s_step;
    const bool entry_forced = o->forced_bad!=0;
    // grid points are handed out one at a time: their cost differs (one fill when the CRC is valid, up to three plus a
    // brute-force bit pick when not), a fixed split would leave most lanes waiting for the slowest
    // the bit-sliced search needs grid points one pixel apart, hysteresis-0 levels that coincide and lie strictly between black and white
    // (else every fill of the reference is refused), an unswept line and a window that stays inside the padded bit strings
    const u8 lev_lo = get_low_level(o->ref, 0), lev_hi = get_high_level(o->ref, 0);
    const bool fast = (step==1)&&(lev_lo==lev_hi)&&(lev_lo==o->ref)&&(lev_lo>o->black)&&(lev_hi<o->white)&&(!o->sweeped)&&(ls>=-(P1F_PAD-2))&&(re<=g.W+P1F_PAD);
    if(fast)
    {
#if defined(SDV_EMU_COUNTERS)
        g_emu_counters[0]++;            // tests/hostemu only: searches that took the bit-sliced path
#endif
        p1_search_fills_cta(c, w, px, g, ls, re, mode, slim);
        for(int p=c.tid;p<P1_GRID*P1_GRID;p+=c.n)
        {
            const int i = p/P1_GRID, j = p-i*P1_GRID;
            u8 cl = 0;
            w->grid[i][j] = p1_search_point(w, o, mode, slim, i, j, ls, re, g.W-1, g.scan_end, entry_forced, &cl);
            w->coll[p] = cl;
        }
        if((c.n<P1F_DIAGS+32)&&(c.tid==0))
        {   // (a group too small to have done it beside the diagonals)
            P1Line t = *o;
            t.coords.start = (i16)(ls+(P1_GRID-1)*step); t.coords.stop = (i16)(re-(P1_GRID-1)*step);
            p1_read_pcm(px, g, mode, &t, 0, slim);
            w->last = t;
        }
    }
    else for(;;)
    {
#if defined(SDV_EMU_COUNTERS)
        if(c.tid==0) g_emu_counters[1]++;
#endif
        const int p = grab_next(&w->s_next);
        if(p>=P1_GRID*P1_GRID) break;
        const int i = p/P1_GRID, j = p-i*P1_GRID;
        P1Line t = *o;
        t.coords.start = (i16)(ls+i*step); t.coords.stop = (i16)(re-j*step);       // PCMLine::coords.setCoordinates: plain assignment here (start < stop)
        p1_read_pcm(px, g, mode, &t, 0, slim);
        CrcH r;
        r.crc = t.words[6]; r.hyst = t.hyst; r.shift = t.shift; r.start = t.coords.start; r.stop = t.coords.stop; r.pad = 0;
        if(t.picked_left&&t.picked_right) r.hyst = 0x0E;
        else if(t.picked_right) r.hyst = 0x0D;
        else if(t.picked_left) r.hyst = 0x0C;
        r.result = p1_crc_ok(&t) ? REF_CRC_OK : REF_BAD_CRC;
        w->grid[i][j] = r;
        w->coll[p] = (t.forced_bad&&(!entry_forced)) ? 1 : 0;
        if(p==(P1_GRID*P1_GRID-1)) w->last = t;
    }
    for(int i=c.tid;i<P1_GRID;i+=c.n) reset_crc_stats(&w->grid[i][P1_SEARCH_STEP_CNT-1], 1);
    c.sync();
    // sequential dependency of the reference loop: after the first bit-picker collision the line stays forced bad
    for(int p=c.tid;p<P1_GRID*P1_GRID;p+=c.n) if(w->coll[p]) {
#if defined(__CUDA_ARCH__)
        atomicMin(&w->s_first_coll, p);
#else
        if(p<w->s_first_coll) w->s_first_coll = p;
#endif
    }
    c.sync();
    const int fc = w->s_first_coll;
    if(fc<P1_GRID*P1_GRID)
    {
        for(int p=fc+1+c.tid;p<P1_GRID*P1_GRID;p+=c.n) w->grid[p/P1_GRID][p%P1_GRID].result = REF_BAD_CRC;
        if((c.tid==0)&&(fc<(P1_GRID*P1_GRID-1)))
        {
            P1Line t = *o;
            t.forced_bad = 1;
            t.coords.start = (i16)(ls+(P1_GRID-1)*step); t.coords.stop = (i16)(re-(P1_GRID-1)*step);
            p1_read_pcm(px, g, mode, &t, 0, slim);
            w->last = t;
        }
        c.sync();
    }
    // right-coordinate vote of every left offset
    for(int i=c.tid;i<P1_GRID;i+=c.n)
    {
        CrcH stats[MAX_COLL_CRCS+1];
        u8 cnt = 0, ofs = 0xFF;
        reset_crc_stats(stats, MAX_COLL_CRCS);
        for(int j=0;j<P1_GRID;j++) if(w->grid[i][j].result==REF_CRC_OK) update_crc_stats(stats, w->grid[i][j], &cnt);
        if(cnt>0)
        {
            find_most_frequent_crc(stats, &cnt, true);
            invalidate_non_frequent(w->grid[i], 0, P1_SEARCH_STEP_CNT-1, cnt, stats[0].crc);
            if(cnt>0) if(pick_level_by_stats(w->grid[i], &ofs, 0, P1_SEARCH_STEP_CNT-1, REF_CRC_OK, 0x0F, SHIFT_MAX)!=SPAN_OK) cnt = 0;
        }
        CrcH lr;
        reset_crc_stats(&lr, 1);
        if(cnt>0)
        {
            lr = w->grid[i][ofs];
            lr.result = REF_CRC_OK;
            w->row_best[i] = stats[0];
            w->row_valid[i] = 1;
        }
        else
        {
            lr.result = REF_BAD_CRC; lr.crc = 0; lr.hyst = HYST_DEPTH_MAX; lr.shift = SHIFT_MAX;
            w->row_valid[i] = 0;
        }
        w->left_res[i] = lr;
    }
    if(c.tid==0) reset_crc_stats(&w->left_res[P1_SEARCH_STEP_CNT-1], 1);
    c.sync();
    // left-coordinate vote
    if(c.tid==0)
    {
        CrcH stats[MAX_COLL_CRCS+1];
        u8 cnt = 0, ofs = 0xFF;
        reset_crc_stats(stats, MAX_COLL_CRCS);
        for(int i=0;i<P1_GRID;i++)
            if(w->row_valid[i]) update_crc_stats_n(stats, w->row_best[i], &cnt, w->row_best[i].result);     // once per hit of the row's CRC
        if(cnt>0)
        {
            find_most_frequent_crc(stats, &cnt, true);
            invalidate_non_frequent(w->left_res, 0, P1_SEARCH_STEP_CNT-1, cnt, stats[0].crc);
            if(cnt>0) if(pick_level_by_stats(w->left_res, &ofs, 0, P1_SEARCH_STEP_CNT-1, REF_CRC_OK, 0x0F, SHIFT_MAX)!=SPAN_OK) cnt = 0;
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

// Binarizer::findPCM1Coordinates (binarizer.cpp:5601-5812): rough edges from the first/last level transition (or the
// coordinate history), then the grid search.
SDV_HD void p1_find_coordinates_cta(const Cta &c, P1Work *w, const u8 *px, const Geom &g, int mode, Coord history)
{
    c.sync();
    Coord dc = history;
    if(!coord_valid(history))
    {
        const P1Line *o = &w->o;
        const int margin = (int)((u16)g.scan_end/16);
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
    p1_search_data_cta(c, w, px, g, mode, dc);
}

// Binarizer::calcRefLevelBySweep + sweepRefLevel for a PCM-1 line (binarizer.cpp:3551-4120), MODE_INSANE only: the full
// coordinate search at every reference level between the black and the white level, then the CRC vote over the levels.
// The dummy line of the reference is cleared through its base class between levels, so its words (and with them the
// "is the CRC valid" test that decides whether a level is searched at all) carry over: kept in w->sweep_d.
SDV_HD void p1_sweep_cta(const Cta &c, P1Work *w, const BinState *b, const u8 *px, const Geom &g)
{
    P1Line *o = &w->o;
    c.sync();
    if(c.tid==0)
    {
        u8 lo = (u8)(o->black+1), hi = (u8)(o->white-1);
        if(MIN_REF_LVL>lo) lo = MIN_REF_LVL;
        if(MAX_REF_LVL<hi) hi = MAX_REF_LVL;
        w->sweep_low = lo; w->sweep_high = hi;
        w->hlim = 0; w->slim = SHIFT_SAFE;
        w->sweep_save = *o;
        p1_clear(&w->sweep_d);
    }
    for(int i=c.tid;i<256;i+=c.n) reset_crc_stats(&w->sw[i], 1);
    c.sync();
    const int lo = w->sweep_low, hi = w->sweep_high;
    for(int ref=hi;ref>=lo;ref--)
    {
        c.sync();
        if(c.tid==0)
        {
            P1Line *d = &w->sweep_d;
            p1_base_clear(d);
            d->black = (u8)lo; d->white = (u8)hi; d->ref = (u8)ref;
            *o = *d;                            // the search and the read work on w->o
            w->search_ok = p1_crc_ok(o) ? 2 : 0;        // 2: CRC "valid" by the carried words -> this level is not searched
        }
        c.sync();
        const bool skip = (w->search_ok==2);
        c.sync();
        if(!skip)
        {
            p1_find_coordinates_cta(c, w, px, g, b->mode, b->def_coord);
            if(c.tid==0) { if(o->coords_set) p1_read_pcm(px, g, b->mode, o, w->hlim, w->slim); }
        }
        if(c.tid==0)
        {
            if(o->picked_left&&o->picked_right) o->hyst = (u8)(o->hyst+HYST_DEPTH_MAX+3);
            else if(o->picked_right) o->hyst = (u8)(o->hyst+HYST_DEPTH_MAX+2);
            else if(o->picked_left) o->hyst = (u8)(o->hyst+HYST_DEPTH_MAX+1);
            if(o->hyst>0x0F) o->hyst = 0x0F;
            CrcH *r = &w->sw[ref];
            if(p1_crc_ok(o)&&coord_valid(o->coords)) { r->result = REF_CRC_OK; r->start = o->coords.start; r->stop = o->coords.stop; r->hyst = o->hyst; r->shift = o->shift; r->crc = o->calc_crc; }
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

// Binarizer::processLine for a PCM-1 line (binarizer.cpp:443-1724).  Result in w->o.
SDV_HD void p1_process_line_cta(const Cta &c, P1Work *w, const BinState *b, bool do_coord_search, const u8 *px, const Geom &g)
{
    P1Line *o = &w->o;
    c.sync();
    if(c.tid==0)
    {
        p1_clear(o);
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
            if(need_bw) p1_find_black_white_cta(c, w, px, g);
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
                p1_read_pcm_cta(c, w, px, g, b->mode, w->hlim, w->slim);
                if(c.tid==0)
                {
                    if(p1_crc_ok(o)) { o->by_ext = 1; w->proc_state = STG_DATA_OK; }
                    else w->proc_state = STG_REF_FIND;
                }
            }
        }
        else if(st==STG_INPUT_LEVEL)
        {
            const bool need_bw = !w->was_bw_scanned;
            c.sync();
            if(need_bw) p1_find_black_white_cta(c, w, px, g);
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
            if(need_bw) p1_find_black_white_cta(c, w, px, g);
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
                if(w->do_coord_search&&FINE_EN_COORD_SEARCH) p1_find_coordinates_cta(c, w, px, g, b->mode, b->def_coord);
                if(c.tid==0)
                {
                    if(!o->coords_set) { w->hlim = HYST_DEPTH_SAFE; w->slim = SHIFT_MIN; }
                    else { w->hlim = b->max_hyst; w->slim = b->max_shift; }
                    w->proc_state = STG_READ_PCM;
                }
            }
        }
        else if(st==STG_REF_SWEEP_RUN)
        {
            p1_sweep_cta(c, w, b, px, g);
            if(c.tid==0) w->proc_state = STG_READ_PCM;
        }
        else if(st==STG_READ_PCM)
        {
            const bool rd1 = o->coords_set!=0;
            c.sync();
            if(rd1) p1_read_pcm_cta(c, w, px, g, b->mode, w->hlim, w->slim);
            if(c.tid==0)
            {
                w->rp_go = 0;
                if(p1_crc_ok(o)) w->proc_state = STG_DATA_OK;
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
                p1_read_pcm_cta(c, w, px, g, b->mode, w->hlim, w->slim);
                if(c.tid==0) { if(p1_crc_ok(o)) w->proc_state = STG_DATA_OK; }
            }
        }
        else if(st==STG_DATA_OK)
        {
            if(c.tid==0)
            {
                if(o->forced_bad) w->proc_state = STG_NO_GOOD;
                else
                {
                    if(p1_words_header(o->words)) p1_set_serv_header(o);
                    w->proc_state = 0xFD;
                }
            }
            c.sync();
            const bool done = (w->proc_state==0xFD);
            c.sync();
            if(done) break;
        }
        else
        {   // STG_NO_GOOD
            if(c.tid==0) { if(p1_crc_ok(o)) p1_set_invalid_crc(o); }
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
