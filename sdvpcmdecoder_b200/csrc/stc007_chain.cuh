// stc007_chain.cuh -- the inter-line chain of VideoToDigital::doBinarize (videotodigital.cpp:698-1815) for STC-007.
//
// The chain is what makes line i depend on line i-1: the Binarizer presets fed back from the last good line, the
// duplicate-line check, the coordinate damper (median of the last 9 valid lines) and the per-frame coordinate
// averages.  Here it is a small state object in device memory that a single thread advances; the heavy per-line
// work is done by the whole block (stc007_line.cuh) or was done ahead by the bulk kernel.
#pragma once
#include "stc007_line.cuh"

namespace sdv {

struct ChainCtx
{
    // ---- first 64 bytes: what the host loop reads back after a launch (one small copy)
    int next_frame;                         // first frame not processed yet
    int stable;                             // 1: frames from next_frame on may be taken from the bulk kernel
    int first_unclean;                      // first frame the bulk kernel could not take (n_frames if none)
    int any_broken;                         // scratch of the deinterleaver
    unsigned long long lines_chain, lines_chain_fast, lines_swept;
    BinState bin;
    u8 pad_hdr[64-16-24-sizeof(BinState)];
    // ---- chain state proper
    u8 field_state, line_dup, m2, pad0;     // m2: TYPE_M2 tape (videotodigital.cpp:1145-1152)
    u16 last_words[8];                      // words of the previous line with PCM in this field (last_line)
    Coord last_valid[COORD_HISTORY_DEPTH];  // last_coord_list
    Coord long_valid[COORD_LONG_HISTORY];   // long_coord_list
    int n_last, n_long;
    Coord frame_avg;
    int n_fv, n_fi;
    Coord frame_valid[SDV_MAX_H];
    Coord frame_invalid[SDV_MAX_H];
};
struct ChainHdr { int next_frame, stable, first_unclean, any_broken; unsigned long long lines_chain, lines_chain_fast, lines_swept; BinState bin; };

// VideoToDigital::medianCoordinates (videotodigital.cpp:348-371): element n/2 of the list sorted by CoordinatePair::operator<
// (start ascending, stop descending: the key below).  Selection by value: count the entries below and equal to a candidate and
// step to the nearest value below / above until the candidate covers position n/2 -- one pass per distinct value on the way,
// where ranking every entry against every other cost n^2 loads of the chain context (global memory) per call.
SDV_HD u32 coord_key(Coord c) { return ((u32)(u16)((int)c.start+32768)<<16)|(u32)(u16)(65535-(u32)(u16)((int)c.stop+32768)); }
SDV_HD Coord median_small(const Coord *v, int n)
{
    if(n==0) return coord_none();
    const u32 target = (u32)(n/2);
    u32 cand = coord_key(v[0]);
    for(int guard=0;guard<=n;guard++)
    {
        u32 less = 0, eq = 0, below = 0, above = 0xFFFFFFFFu;
        for(int i=0;i<n;i++)
        {
            const u32 k = coord_key(v[i]);
            if(k<cand) { less++; if(k>=below) below = k; }
            else if(k==cand) eq++;
            else if(k<above) above = k;
        }
        if((less<=target)&&(target<less+eq)) break;
        cand = (target<less) ? below : above;
    }
    Coord r;
    r.start = (i16)((int)(cand>>16)-32768);
    r.stop = (i16)((int)(65535u-(cand&0xFFFFu))-32768);
    return r;
}
// Same for the per-frame lists (up to one entry per line), spread over the block.  Result in *out (shared or global).
SDV_HD void median_cta(const Cta &c, const Coord *v, int n, Coord *out, int *scratch)
{
    c.sync();
    if(c.tid==0) *scratch = 0;
    c.sync();
    if(n>0)
    {   // usual case on a steady tape: every entry is the same
        const Coord first = v[0];
        bool differ = false;
        for(int i=c.tid;i<n;i+=c.n) if(!coord_eq(v[i], first)) differ = true;
        if(differ) *scratch = 1;
    }
    c.sync();
    if((n>0)&&(*scratch==0)) { if(c.tid==0) *out = v[0]; c.sync(); return; }
    if((n==0)&&(c.tid==0)) *out = coord_none();
    if(n==0) { c.sync(); return; }
    // Selection by value instead of ranking every entry against every other (n up to one entry per line): in CoordinatePair's order
    // (start ascending, stop descending) a pair is the key (start << 16) | ~stop; count the entries below and equal to a candidate
    // value, and step to the nearest value below / above until the candidate covers position n/2.  A frame holds a handful of
    // distinct pairs, so this is a few passes over the list.
#if defined(__CUDA_ARCH__)
    __shared__ u32 sc[4];
#else
    u32 sc[4];
#endif
    u32 cand = coord_key(v[0]);
    const u32 target = (u32)(n/2);
    for(int guard=0;guard<=n+1;guard++)
    {
        if(c.tid==0) { sc[0] = 0; sc[1] = 0; sc[2] = 0; sc[3] = 0xFFFFFFFFu; }
        c.sync();
        u32 less = 0, eq = 0, below = 0, above = 0xFFFFFFFFu;
        for(int i=c.tid;i<n;i+=c.n)
        {
            const u32 k = coord_key(v[i]);
            if(k<cand) { less++; if(k>below) below = k; }
            else if(k==cand) eq++;
            else if(k<above) above = k;
        }
#if defined(__CUDA_ARCH__)
        if(less) { atomicAdd(&sc[0], less); atomicMax(&sc[2], below); }
        if(eq) atomicAdd(&sc[1], eq);
        if(above!=0xFFFFFFFFu) atomicMin(&sc[3], above);
#else
        sc[0] += less; sc[1] += eq; if(below>sc[2]) sc[2] = below; if(above<sc[3]) sc[3] = above;
#endif
        c.sync();
        const u32 n_less = sc[0], n_eq = sc[1], k_below = sc[2], k_above = sc[3];
        c.sync();
        if((n_less<=target)&&(target<n_less+n_eq)) break;
        cand = (target<n_less) ? k_below : k_above;
    }
    if(c.tid==0)
    {
        Coord r;
        r.start = (i16)((int)(cand>>16)-32768);
        r.stop = (i16)((int)(65535u-(cand&0xFFFFu))-32768);
        *out = r;
    }
    c.sync();
}

// line_dup: bit 0 = duplicate-line check, bit 1 = M2 sample format
SDV_HD void chain_reset(ChainCtx *x, int mode, int line_dup)
{
    bin_set_mode(&x->bin, mode);
    x->bin.def_coord = coord_none();
    bin_reset_good(&x->bin);
    x->field_state = FIELD_NEW; x->line_dup = (u8)(line_dup&1); x->m2 = (u8)((line_dup>>1)&1);
    for(int i=0;i<8;i++) x->last_words[i] = 0;
    x->n_last = x->n_long = 0; x->n_fv = x->n_fi = 0;
    x->frame_avg = coord_none();
    x->next_frame = 0; x->stable = 0; x->first_unclean = 0; x->any_broken = 0;
    x->lines_chain = x->lines_chain_fast = x->lines_swept = 0;
}

// Frame start (videotodigital.cpp:774-823); [first] = the frame that carries the NEW_FILE service line.
SDV_HD void chain_frame_start(ChainCtx *x, bool first)
{
    x->frame_avg = median_small(x->long_valid, x->n_long);
    if(coord_valid(x->frame_avg)) bin_set_coords2(&x->bin, x->frame_avg.start, x->frame_avg.stop);
    if(first)
    {
        x->n_last = x->n_long = x->n_fv = x->n_fi = 0;
        if(!coord_valid(x->frame_avg)) bin_reset_good(&x->bin);
    }
    x->field_state = FIELD_NEW;
}
// END_FIELD service line (videotodigital.cpp:1028-1048).
SDV_HD void chain_field_end(ChainCtx *x)
{
    x->field_state = FIELD_NEW;
    for(int i=0;i<8;i++) x->last_words[i] = 0;
}
// END_FRAME service line (videotodigital.cpp:1659-1723); the medians were computed by median_cta().
SDV_HD void chain_frame_end(ChainCtx *x, Coord med_valid, Coord med_invalid)
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

SDV_HD bool delta_warning(i32 ds, i32 de, int lim) { return (ds<=-lim)||(ds>=lim)||(de<=-lim)||(de>=lim); }

// What doBinarize does with one decoded line (videotodigital.cpp:1006-1522), thread 0 only.
SDV_HD void chain_line(ChainCtx *x, Line *line)
{
    if(line->service!=0)
    {
        if((line->service==SDV_SRV_CTRL_BLOCK)&&(x->field_state==FIELD_NEW)) x->field_state = FIELD_SAFE;
        return;
    }
    bool has_data = line_has_markers(line);
    bool has_pcm = line_crc_ok(line)||has_data;
    line->m2 = x->m2;
    if(has_pcm&&(x->field_state==FIELD_NEW)) x->field_state = FIELD_UNSAFE;
    if(line_crc_ok(line))
    {
        if(x->line_dup)
        {
            if(x->field_state==FIELD_UNSAFE)
            {   // first PCM line of a field without a preceding Control Block (en_first_line_dup)
                bin_set_good(&x->bin, line);
                if(FINE_FIRST_LINE_DUP) line->forced_bad = 1;
            }
            else
            {
                bool same = words_diff8(line->words, x->last_words)<=(BITS_PCM_DATA/32);
                if((!words_almost_silent(line->words, x->m2!=0))&&same) line->forced_bad = 1;
            }
        }
        if(line_crc_ok_ign(line))
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
                    if(delta_warning(ds, de, (int)(u8)(line_get_ppb(line)*3))) line->forced_bad = 1;
                }
            }
        }
        if(line_crc_ok(line)) bin_set_good(&x->bin, line);
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
    if(has_pcm) for(int i=0;i<8;i++) x->last_words[i] = line->words[i];
}

// The line object a preset-only decode produces when its first (hysteresis 0, shift 0) candidate has a valid CRC
// (STG_INPUT_ALL -> readPCMdata -> STG_DATA_OK, binarizer.cpp:774-931,1560-1640).
SDV_HD void line_from_fast(Line *l, const BinState *b, const u16 *words9)
{
    line_clear(l);
    for(int i=0;i<9;i++) l->words[i] = words9[i];
    l->calc_crc = words9[8];
    l->coords = b->def_coord;
    l->black = b->def_black; l->white = b->def_white;
    l->ref = l->ref_low = l->ref_high = b->def_ref;
    l->by_ext = 1; l->bw_set = 1; l->wflags = 1;
    l->ppb = make_ppb(b->def_coord);
    if(words_control_block(l->words)) line_set_serv_ctrl_blk(l);
}

// Is the chain in the steady state in which whole frames can be taken from the bulk kernel run with (ref, coords)?
// All remembered coordinates equal the preset, so neither the damper nor the frame-start average can move it.
SDV_HD bool chain_is_stable(const ChainCtx *x)
{
    if(!bin_fast_ready(&x->bin)) return false;
    if(x->n_last!=COORD_HISTORY_DEPTH) return false;
    for(int i=0;i<x->n_last;i++) if(!coord_eq(x->last_valid[i], x->bin.def_coord)) return false;
    if(x->n_long<1) return false;
    for(int i=0;i<x->n_long;i++) if(!coord_eq(x->long_valid[i], x->bin.def_coord)) return false;
    if(!coord_eq(x->frame_avg, x->bin.def_coord)) return false;
    return true;
}
// Do two chain states at a frame boundary lead to the same decode of everything that follows?  (Counters and the
// per-frame lists, empty between frames, aside.)
SDV_HD bool chain_state_equal(const ChainCtx *a, const ChainCtx *b)
{
    const BinState &p = a->bin, &q = b->bin;
    if((p.def_ref!=q.def_ref)||(p.def_black!=q.def_black)||(p.def_white!=q.def_white)||!coord_eq(p.def_coord, q.def_coord)) return false;
    if((p.max_hyst!=q.max_hyst)||(p.max_shift!=q.max_shift)||(p.mode!=q.mode)) return false;
    if((a->field_state!=b->field_state)||(a->line_dup!=b->line_dup)||(a->m2!=b->m2)) return false;
    for(int i=0;i<8;i++) if(a->last_words[i]!=b->last_words[i]) return false;
    if((a->n_last!=b->n_last)||(a->n_long!=b->n_long)) return false;
    for(int i=0;i<a->n_last;i++) if(!coord_eq(a->last_valid[i], b->last_valid[i])) return false;
    for(int i=0;i<a->n_long;i++) if(!coord_eq(a->long_valid[i], b->long_valid[i])) return false;
    if(!coord_eq(a->frame_avg, b->frame_avg)) return false;
    return (a->n_fv==b->n_fv)&&(a->n_fi==b->n_fi);
}
// The chain state at the head of a frame, reduced to what the frames after it can see (the per-frame lists are empty there).
struct ChainSnap
{
    BinState bin; u8 field_state, n_last, n_long, pad;
    u16 last_words[8];
    Coord last_valid[COORD_HISTORY_DEPTH], long_valid[COORD_LONG_HISTORY], frame_avg;
};
SDV_HD void chain_snap(const ChainCtx *x, ChainSnap *s)
{
    s->bin = x->bin; s->field_state = x->field_state; s->n_last = (u8)x->n_last; s->n_long = (u8)x->n_long; s->pad = 0;
    for(int i=0;i<8;i++) s->last_words[i] = x->last_words[i];
    for(int i=0;i<COORD_HISTORY_DEPTH;i++) s->last_valid[i] = (i<x->n_last) ? x->last_valid[i] : coord_none();
    for(int i=0;i<COORD_LONG_HISTORY;i++) s->long_valid[i] = (i<x->n_long) ? x->long_valid[i] : coord_none();
    s->frame_avg = x->frame_avg;
}
SDV_HD bool chain_snap_equal(const ChainSnap *a, const ChainSnap *b)
{
    const BinState &p = a->bin, &q = b->bin;
    if((p.def_ref!=q.def_ref)||(p.def_black!=q.def_black)||(p.def_white!=q.def_white)||!coord_eq(p.def_coord, q.def_coord)) return false;
    if((p.max_hyst!=q.max_hyst)||(p.max_shift!=q.max_shift)||(p.mode!=q.mode)) return false;
    if((a->field_state!=b->field_state)||(a->n_last!=b->n_last)||(a->n_long!=b->n_long)) return false;
    for(int i=0;i<8;i++) if(a->last_words[i]!=b->last_words[i]) return false;
    for(int i=0;i<COORD_HISTORY_DEPTH;i++) if(!coord_eq(a->last_valid[i], b->last_valid[i])) return false;
    for(int i=0;i<COORD_LONG_HISTORY;i++) if(!coord_eq(a->long_valid[i], b->long_valid[i])) return false;
    return coord_eq(a->frame_avg, b->frame_avg);
}
// Account for [n] clean frames decoded by the bulk kernel (every line valid with the preset coordinates).
SDV_HD void chain_skip_clean_frames(ChainCtx *x, int n)
{
    Coord c = x->bin.def_coord;
    for(int i=0;(i<n)&&(i<COORD_LONG_HISTORY);i++)
    {
        if(x->n_long==COORD_LONG_HISTORY) { for(int k=1;k<COORD_LONG_HISTORY;k++) x->long_valid[k-1] = x->long_valid[k]; x->n_long--; }
        x->long_valid[x->n_long++] = c;
    }
    x->frame_avg = c;
    x->field_state = FIELD_NEW;
    for(int i=0;i<8;i++) x->last_words[i] = 0;
}

// ------------------------------------------------------------------------------------------------ record export
SDV_HD void export_line(const Line *l, sdv_line_rec *r, sdv_line_aux *a)
{
    sdv_line_rec t;
    for(int i=0;i<9;i++) t.words[i] = l->words[i];
    u16 f = 0;
    if(line_crc_ok(l)) f |= SDV_LF_CRC_OK;
    if(line_crc_ok_ign(l)) f |= SDV_LF_CRC_OK_IGN;
    if(l->forced_bad) f |= SDV_LF_FORCED_BAD;
    if(l->bw_set) f |= SDV_LF_BW_SET;
    if(l->coords_set) f |= SDV_LF_COORDS_SET;
    if(l->sweeped) f |= SDV_LF_REF_SWEEP;
    if(l->by_ext) f |= SDV_LF_BY_EXT;
    if(line_has_markers(l)) f |= SDV_LF_MARKERS;
    if(line_has_start(l)) f |= SDV_LF_START_MARK;
    if(line_has_stop(l)) f |= SDV_LF_STOP_MARK;
    if(words_almost_silent(l->words, l->m2!=0)) f |= SDV_LF_ALMOST_SILENT;
    t.flags = f;
    t.ref = l->ref; t.black = l->black; t.white = l->white; t.hyst = l->hyst;
    t.data_start = l->coords.start; t.data_stop = l->coords.stop;
    t.shift = l->shift; t.service_type = l->service;
    t.mark_stages = (u8)(l->mst|(l->med<<4));
    t.reserved = 0;
#if defined(__CUDA_ARCH__)
    {   // record buffers are 16-byte aligned (checked at the C ABI): two 16-byte stores instead of sixteen 2-byte ones
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
        u.marker_start_bg = l->m_bg; u.marker_start_ed = l->m_ed; u.marker_stop_ed = l->m_stop;
        u16 m = ((!l->forced_bad)&&l->wflags) ? 0x1FF : 0;
        u.word_crc_mask = m; u.word_valid_mask = m;
        u.pad[0] = u.pad[1] = u.pad[2] = u.pad[3] = 0;
        *a = u;
    }
}

// ------------------------------------------------------------------------------------------------ a batch of preset-decoded lines
// Result of the preset decode of one line (look-ahead of the chain kernel).
struct FastRes { u16 words[9]; u16 ok; };
// What line i of the batch needs to know about the lines before it.
struct FastPlan { i16 prev; u16 m_incl; u8 fs_before; u8 cb; u16 pad; };

// chain_line() + export_line() for the leading lines of a batch whose preset decode was valid, done in parallel:
// such lines leave the Binarizer presets unchanged, so each one only needs (a) the field state before it, (b) the words
// of the previous non-service line and (c) how many coordinates were pushed to the damper before it -- a short scalar
// prefix pass by thread 0 -- after which every line is finished by its own thread.  Returns the number of lines taken.
SDV_HD int chain_fast_batch(const Cta &c, ChainCtx *x, const FastRes *fr, int nb, FastPlan *plan, int *s_n,
                            sdv_line_rec *recs, sdv_line_aux *aux)
{
    c.sync();
    for(int i=c.tid;i<nb;i+=c.n) plan[i].cb = (fr[i].ok&&words_control_block(fr[i].words)) ? 1 : 0;
    c.sync();
    if(c.tid==0)
    {
        int n = 0, prev = -1, m = 0;
        u8 fs = x->field_state;
        for(;n<nb;n++)
        {
            if(!fr[n].ok) break;
            plan[n].fs_before = fs; plan[n].prev = (i16)prev;
            if(plan[n].cb) { if(fs==FIELD_NEW) fs = FIELD_SAFE; }
            else { m++; prev = n; fs = FIELD_INIT; }
            plan[n].m_incl = (u16)m;
        }
        *s_n = n;
    }
    c.sync();
    const int n = *s_n;
    const Coord pc = x->bin.def_coord;
    for(int i=c.tid;i<n;i+=c.n)
    {
        Line l;
        line_from_fast(&l, &x->bin, fr[i].words);
        if(l.service==0)
        {
            l.m2 = x->m2;
            u8 fs = plan[i].fs_before;
            if(fs==FIELD_NEW) fs = FIELD_UNSAFE;
            if(x->line_dup)
            {
                if(fs==FIELD_UNSAFE) { if(FINE_FIRST_LINE_DUP) l.forced_bad = 1; }
                else
                {
                    const u16 *pw = (plan[i].prev>=0) ? fr[plan[i].prev].words : x->last_words;
                    if((words_diff8(l.words, pw)<=(BITS_PCM_DATA/32))&&(!words_almost_silent(l.words, x->m2!=0))) l.forced_bad = 1;
                }
            }
            // coordinate damper: window = last 9 of (history ++ m copies of the preset coordinates)
            const int m = plan[i].m_incl;
            int nwin = x->n_last+m; if(nwin>COORD_HISTORY_DEPTH) nwin = COORD_HISTORY_DEPTH;
            if(nwin>(COORD_HISTORY_DEPTH/2))
            {
                Coord win[COORD_HISTORY_DEPTH];
                const int n_new = (m<nwin) ? m : nwin, n_old = nwin-n_new;
                for(int q=0;q<n_old;q++) win[q] = x->last_valid[x->n_last-n_old+q];
                for(int q=n_old;q<nwin;q++) win[q] = pc;
                Coord target = median_small(win, nwin);
                if(!coord_valid(target)) target = x->frame_avg;
                if(coord_valid(target))
                {
                    i16 ds = (i16)(l.coords.start-target.start), de = (i16)(l.coords.stop-target.stop);
                    if(delta_warning(ds, de, (int)(u8)(line_get_ppb(&l)*3))) l.forced_bad = 1;
                }
            }
        }
        if(l.service==0)
        {   // per-frame list of valid coordinates (frame_coord_list): entry number = coordinates pushed before this line
            const int slot = x->n_fv+(int)plan[i].m_incl-1;
            if(slot<SDV_MAX_H) x->frame_valid[slot] = pc;
        }
        export_line(&l, recs+i, aux ? aux+i : (sdv_line_aux *)0);
    }
    c.sync();
    if((c.tid==0)&&(n>0))
    {
        const int m = plan[n-1].m_incl;
        const bool last_cb = plan[n-1].cb!=0;
        const int last = last_cb ? plan[n-1].prev : (n-1);
        if(last>=0) { for(int q=0;q<8;q++) x->last_words[q] = fr[last].words[q]; x->field_state = FIELD_INIT; }
        else if(x->field_state==FIELD_NEW) x->field_state = FIELD_SAFE;      // only Control Blocks so far
        for(int q=0;(q<m)&&(q<COORD_HISTORY_DEPTH);q++)
        {
            if(x->n_last==COORD_HISTORY_DEPTH) { for(int k=1;k<COORD_HISTORY_DEPTH;k++) x->last_valid[k-1] = x->last_valid[k]; x->n_last--; }
            x->last_valid[x->n_last++] = pc;
        }
        x->n_fv = (x->n_fv+m<SDV_MAX_H) ? (x->n_fv+m) : SDV_MAX_H;      // the entries were written by the line threads
        x->lines_chain += (unsigned long long)n; x->lines_chain_fast += (unsigned long long)n;
    }
    c.sync();
    return n;
}

}   // namespace sdv
