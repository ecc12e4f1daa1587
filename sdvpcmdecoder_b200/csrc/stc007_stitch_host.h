// stc007_stitch_host.h -- the frame-to-frame decision chain of STC007DataStitcher (host code of the library).
//
// Per frame the reference decides, from the previous frame's outcome (frasm_f0), the trims of this frame and the next
// (frasm_f1, frasm_f2) and the seam statistics (tryPadding / findPadding), how the two fields are stacked:
//   findFieldStitching   stc007datastitcher.cpp:2929-4276     -> Stitcher::find_field_stitching
//   getAssemblyFieldOrder                       4278-4423     -> Stitcher::assembly_order
//   fillFrameForOutput                          4588-5388     -> Stitcher::fill_frame
//   findPadding (ranking + acceptance)          1743-2054     -> pad_decide
//   detectVideoStandard                         2773-2925     -> Stitcher::detect_standard
// The data is a few bytes per frame and every step depends on the one before, so this part stays on the host; the line
// data behind it (trims, seam sweeps, deinterleave) is device work over all frames at once.  The seam statistics reach the
// chain through SeamOracle, which the library fills from stc007_seam_kernel results (and tests/hostemu from the host
// build of the same block logic).
#pragma once
#include <stdint.h>
#include <string.h>
#include <algorithm>
#include <vector>
#include "stc007_stitch.cuh"
#include "stc007_cwd.cuh"

namespace sdv {

// Seam kinds: which field vectors tryPadding / findPadding are given.
enum { SEAM_IN_TFF = 0,     // frame A odd  -> frame A even
       SEAM_IN_BFF = 1,     // frame A even -> frame A odd
       SEAM_TT = 2,         // frame A even -> frame B odd     (A TFF, B TFF)
       SEAM_BB = 3,         // frame A odd  -> frame B even    (A BFF, B BFF)
       SEAM_TB = 4,         // frame A even -> frame B even    (A TFF, B BFF)
       SEAM_BT = 5,         // frame A odd  -> frame B odd     (A BFF, B TFF)
       SEAM_KINDS = 6 };

struct PadStats { uint16_t index, valid, silent, unchecked, broken; };
inline bool pad_less(const PadStats &a, const PadStats &b)
{   // FieldStitchStats::operator< (frametrimset.cpp:312-371): a total order up to identical entries
    if(a.broken!=b.broken) return a.broken<b.broken;
    if(a.valid!=b.valid) return a.valid>b.valid;
    if(a.unchecked!=b.unchecked) return a.unchecked<b.unchecked;
    if(a.silent!=b.silent) return a.silent<b.silent;
    return a.index<b.index;
}

struct PadDecision { uint16_t padding; uint8_t result; uint8_t last_pad_counter; };

// findPadding's decision over the sweep statistics of one seam (stats[pad], pad = 0..max_padding-1, as tryPadding
// leaves them).  lpf: lines per field of the video standard (0 = unknown); lim: the unchecked-burst limit.
inline PadDecision pad_decide(const sdv_stitch_stats *stats, int max_padding, uint32_t f1_size, int lpf, int lim, bool ecc_on)
{
    enum { MAX_PAD = 32, UNCH_DELTA = 8, BURST_SILENCE = 8 };
    PadDecision o; o.padding = 0; o.result = SDV_DS_RET_NO_PAD; o.last_pad_counter = 0xFF;
    const uint32_t n1 = f1_size&0xFFFFu;
    if(lpf) o.padding = (uint16_t)((n1>(uint32_t)lpf) ? 0 : (lpf-n1));
    if(!ecc_on) return o;
    PadStats sd[MAX_PAD];
    for(int i=0;i<max_padding;i++) { sd[i].index = sd[i].valid = 0; sd[i].silent = sd[i].unchecked = sd[i].broken = 0xFF; }
    int min_broken = 0xFFFF, no_brk = 0;
    for(int pad=0;pad<max_padding;pad++)
    {   // the reference stops sweeping once an unbroken padding has been followed by a broken one: later entries keep their cleared values
        const sdv_stitch_stats &g = stats[pad];
        sd[pad].index = g.index; sd[pad].valid = g.valid; sd[pad].silent = g.silent; sd[pad].unchecked = g.unchecked; sd[pad].broken = g.broken;
        if(min_broken>sd[pad].broken) { min_broken = sd[pad].broken; if(min_broken==0) no_brk = pad; }
        else if(min_broken==0)
        {
            if((sd[no_brk].valid>0)&&(sd[no_brk].unchecked<lim)&&(sd[pad].broken>0)) break;
        }
    }
    std::sort(sd, sd+max_padding, pad_less);
    o.last_pad_counter = (uint8_t)sd[0].broken;
    if(sd[0].silent<BURST_SILENCE)
    {
        if(sd[0].unchecked<lim)
        {
            if((sd[0].broken<2)&&(sd[0].broken<sd[1].broken)) { o.result = SDV_DS_RET_OK; o.padding = sd[0].index; }
            else if((((int16_t)sd[0].valid-(int16_t)sd[1].valid)>UNCH_DELTA)&&(sd[0].broken==0)) { o.result = SDV_DS_RET_OK; o.padding = sd[0].index; }
        }
        else
        {   // nothing checkable at the top: rank by the valid runs among the paddings that are
            for(int pad=0;pad<max_padding;pad++)
            {
                sd[pad].broken = (uint16_t)min_broken;
                if(sd[pad].unchecked>=lim) sd[pad].broken = 0xFF;
            }
            std::sort(sd, sd+max_padding, pad_less);
            if((sd[0].unchecked<lim)&&(((int16_t)sd[0].valid-(int16_t)sd[1].valid)>UNCH_DELTA)) { o.result = SDV_DS_RET_OK; o.padding = sd[0].index; }
        }
    }
    else o.result = SDV_DS_RET_SILENCE;
    return o;
}

// detectAudioResolution (stc007datastitcher.cpp:2207-2770) with its 65-entry history (stats_resolution, getProbableResolution
// 2142-2204): from the detected resolution of the four fields of frames A and B (ST_RES_*) to their deinterleaver modes
// (SDV_RES_MODE_*).  A chain of its own: it depends on nothing but the fields' own detections, so the library runs it over
// all frames before the seam sweeps, which need its modes.
struct ResChain
{
    uint8_t hist[65]; uint8_t pos;
    void reset() { memset(hist, ST_RES_UNKNOWN, sizeof(hist)); pos = 0; }
    void push(uint8_t r) { hist[pos] = r; pos = (uint8_t)((pos+1)%65); }
    uint8_t probable() const
    {
        int c14 = 0, c16 = 0;
        for(int i=0;i<65;i++) { if(hist[i]==ST_RES_14BIT) c14++; if(hist[i]==ST_RES_16BIT) c16++; }
        if((c14>0)||(c16>0)) return (c14<c16) ? ST_RES_16BIT : ST_RES_14BIT;
        return ST_RES_UNKNOWN;
    }
    static uint8_t fixed(uint8_t r) { return (r==ST_RES_16BIT) ? SDV_RES_MODE_16BIT : SDV_RES_MODE_14BIT; }
    static uint8_t autom(uint8_t r) { return (r==ST_RES_16BIT) ? SDV_RES_MODE_16BIT_AUTO : SDV_RES_MODE_14BIT_AUTO; }
    // one frame's own pair when at least one of its fields is known
    static void pair(uint8_t o, uint8_t e, uint8_t *mo, uint8_t *me)
    {
        if(o==ST_RES_UNKNOWN) { *me = fixed(e); *mo = autom(e); }
        else if(e==ST_RES_UNKNOWN) { *mo = fixed(o); *me = autom(o); }
        else { *mo = fixed(o); *me = fixed(e); }
    }
    // out[0..3] = modes of A odd, A even, B odd, B even
    void step(uint8_t ao, uint8_t ae, uint8_t bo, uint8_t be, uint8_t *out)
    {
        if((ao==ST_RES_14BIT)||(ao==ST_RES_16BIT)) push(ao);
        if((ae==ST_RES_14BIT)||(ae==ST_RES_16BIT)) push(ae);
        if((ao==ST_RES_UNKNOWN)&&(ae==ST_RES_UNKNOWN))
        {
            if((bo==ST_RES_UNKNOWN)&&(be==ST_RES_UNKNOWN)) out[0] = out[1] = out[2] = out[3] = autom(probable());
            else if(bo==ST_RES_UNKNOWN) { out[3] = fixed(be); out[0] = out[1] = out[2] = autom(be); }
            else if(be==ST_RES_UNKNOWN) { out[2] = fixed(bo); out[0] = out[1] = out[3] = autom(bo); }
            else if((bo==be)&&(bo==ST_RES_16BIT)) { out[2] = out[3] = SDV_RES_MODE_16BIT; out[0] = out[1] = SDV_RES_MODE_16BIT_AUTO; }
            else { out[2] = fixed(bo); out[3] = fixed(be); out[0] = out[1] = SDV_RES_MODE_14BIT_AUTO; }
        }
        else
        {
            pair(ao, ae, &out[0], &out[1]);
            if((bo==ST_RES_UNKNOWN)&&(be==ST_RES_UNKNOWN)) out[2] = out[3] = autom(probable());
            else pair(bo, be, &out[2], &out[3]);
        }
    }
};

// Where the seam statistics come from.  try_padding: the DS_RET_* code of tryPadding(seam of [frame], padding);
// sweep: the statistics of paddings 0..31 of that seam.  Either returns false when the answer is not available yet
// (the library then computes the missing seams on the device and runs the frame again).
struct SeamOracle
{
    virtual bool try_padding(int frame, int kind, int padding, uint8_t *result) = 0;
    virtual bool sweep(int frame, int kind, const sdv_stitch_stats **stats32) = 0;
    virtual ~SeamOracle() {}
};

// FrameAsmSTC007, the fields the decisions use.
struct FrameSt
{
    uint16_t odd_lines, even_lines;
    uint8_t  order; bool order_preset, order_guessed;
    uint8_t  video_std; bool std_preset;
    uint16_t inner_pad, outer_pad;
    bool inner_ok, outer_ok, inner_silence, outer_silence;
    uint8_t tff_cnt, bff_cnt;
    uint8_t odd_res, even_res;  // deinterleaver modes of the fields (SDV_RES_MODE_*)
    void clear_misc()
    {
        odd_res = even_res = SDV_RES_MODE_14BIT;
        odd_lines = even_lines = 0; order = ST_ORDER_UNK; order_preset = order_guessed = false;
        video_std = ST_VID_UNKNOWN; std_preset = false; inner_pad = outer_pad = 0;
        inner_ok = outer_ok = false; inner_silence = outer_silence = true; tff_cnt = bff_cnt = 0;
    }
    bool is_tff() const { return order==ST_ORDER_TFF; }
    bool is_bff() const { return order==ST_ORDER_BFF; }
    bool order_set() const { return (order==ST_ORDER_TFF)||(order==ST_ORDER_BFF); }
    void set_order(uint8_t o) { if(!order_preset) order = o; }                 // setOrderTFF / setOrderBFF
    void set_order_unknown() { if(!order_preset) { order = ST_ORDER_UNK; order_guessed = false; } }
    void preset_order(uint8_t o) { order_preset = true; order_guessed = false; order = o; }
    void std_soft(uint8_t s) { if(!std_preset&&(s<3)) video_std = s; }
    uint16_t lines(int even) const { return even ? even_lines : odd_lines; }
    uint8_t res(int even) const { return even ? even_res : odd_res; }
};

struct StitchSettings
{
    uint8_t video_std;          // preset (ST_VID_UNKNOWN = detect by line count)
    uint8_t field_order;        // preset (ST_ORDER_UNK = detect)
    uint8_t res16;              // audio resolution preset: 0 = 14 bit, 1 = 16 bit, 2 = detected per field (the modes come with every step())
    uint8_t p_corr, q_corr;
    uint8_t max_unch14, max_unch16;
    uint8_t fix_cut_above;      // setFineTopLineFix
    uint8_t mask_seams;         // setFineMaskSeams
};

// State the reference keeps from frame to frame (everything else is rebuilt for every frame).
struct StitchCarry
{
    FrameSt f0;
    uint8_t last_pad_counter;
    uint8_t order_hist[65]; uint8_t order_pos;      // circarray<uint8_t, STATS_DEPTH> stats_field_order
    void reset()
    {
        f0.clear_misc(); last_pad_counter = 0xFF;
        memset(order_hist, ST_ORDER_UNK, sizeof(order_hist)); order_pos = 0;
    }
};

struct Stitcher
{
    StitchSettings set;
    StitchCarry st;
    SeamOracle *seams;
    bool missing;               // a seam answer was not available: the frame's result is void
    int cur;                    // frame A
    FrameSt f1, f2;

    int lpf_of(uint8_t std) const { return (std==ST_VID_PAL) ? ST_LINES_PF_PAL : ((std==ST_VID_NTSC) ? ST_LINES_PF_NTSC : 0); }
    uint8_t probable_order() const
    {
        int t = 0, b = 0;
        for(int i=0;i<65;i++) { if(st.order_hist[i]==ST_ORDER_TFF) t++; else if(st.order_hist[i]==ST_ORDER_BFF) b++; }
        if((t>0)||(b>0)) return (t<b) ? ST_ORDER_BFF : ST_ORDER_TFF;
        return ST_ORDER_UNK;
    }
    void push_order(uint8_t o) { st.order_hist[st.order_pos] = o; st.order_pos = (uint8_t)((st.order_pos+1)%65); }

    uint8_t try_pad(int kind, int padding)
    {
        uint8_t r = SDV_DS_RET_NO_PAD;
        if(padding>63) padding = 63;
        if(!seams->try_padding(cur, kind, padding, &r)) missing = true;
        return r;
    }
    uint8_t find_pad(int kind, uint32_t f1_size, uint16_t *padding)
    {
        const bool ecc = set.p_corr||set.q_corr;
        int max_padding = 32, lim = set.max_unch14;
        // getResolutionForSeam (stc007datastitcher.cpp:1256-1269) of the two fields of the seam
        static const int tab[SEAM_KINDS][3] = { {0, 0, 1}, {1, 0, 0}, {1, 1, 0}, {0, 1, 1}, {1, 1, 1}, {0, 1, 0} };   // field 1 parity, frame offset of field 2, field 2 parity
        const uint8_t sm = seam_res_mode(f1.res(tab[kind][0]), (tab[kind][1] ? f2 : f1).res(tab[kind][2]));
        const bool seam16 = (sm==SDV_RES_MODE_16BIT)||(sm==SDV_RES_MODE_16BIT_AUTO);
        if(seam16||!set.q_corr) { max_padding = 16; lim = set.max_unch16; }
        const sdv_stitch_stats *s = 0;
        if(ecc&&!seams->sweep(cur, kind, &s)) { missing = true; *padding = 0; st.last_pad_counter = 0xFF; return SDV_DS_RET_NO_PAD; }
        const PadDecision d = pad_decide(s, max_padding, f1_size, lpf_of(f1.video_std), lim, ecc);
        *padding = d.padding; st.last_pad_counter = d.last_pad_counter;
        return d.result;
    }

    // detectVideoStandard: frame A's standard and both frames' order presets.
    void detect_standard(const FrameTrim &ta, const FrameTrim &tb)
    {
        f1.video_std = ST_VID_UNKNOWN;
        if(set.video_std==ST_VID_UNKNOWN)
        {
            f1.std_preset = false;
            const int a = f1.odd_lines, b = f1.even_lines, c = f2.odd_lines, d = f2.even_lines;
            if((a>ST_LINES_PF_MAX_PAL)||(b>ST_LINES_PF_MAX_PAL)||(c>ST_LINES_PF_MAX_PAL)||(d>ST_LINES_PF_MAX_PAL)) f1.video_std = ST_VID_UNKNOWN;
            else if((a>ST_LINES_PF_MAX_NTSC)||(b>ST_LINES_PF_MAX_NTSC)||(c>ST_LINES_PF_MAX_NTSC)||(d>ST_LINES_PF_MAX_NTSC)) f1.video_std = ST_VID_PAL;
            else
            {
                const int max_line = std::max(ta.odd.max_line, ta.even.max_line);
                f1.video_std = (max_line<=(ST_LINES_PF_PAL-16)*2) ? ST_VID_NTSC : ST_VID_PAL;
            }
        }
        else { f1.std_preset = true; f1.video_std = set.video_std; }
        if(f1.video_std==ST_VID_UNKNOWN) f1.video_std = st.f0.video_std;
        if((set.field_order==ST_ORDER_TFF)||(set.field_order==ST_ORDER_BFF)) { f1.preset_order(set.field_order); f2.preset_order(set.field_order); }
        else { f2.order_preset = false; f2.set_order_unknown(); }
        (void)tb;
    }

    static int inner_kind(uint8_t order) { return (order==ST_ORDER_BFF) ? SEAM_IN_BFF : SEAM_IN_TFF; }
    static int outer_kind(uint8_t a, uint8_t b)
    {
        if(a==ST_ORDER_TFF) return (b==ST_ORDER_TFF) ? SEAM_TT : SEAM_TB;
        return (b==ST_ORDER_BFF) ? SEAM_BB : SEAM_BT;
    }
    static uint8_t other(uint8_t o) { return (o==ST_ORDER_TFF) ? ST_ORDER_BFF : ST_ORDER_TFF; }
    // first field of order o is the even one?
    static int first_even(uint8_t o) { return (o==ST_ORDER_BFF) ? 1 : 0; }

    // findFieldStitching.  The reference's stages, with the TFF / BFF twins folded into one by the order they assume.
    void find_field_stitching()
    {
        enum { S_TRY_PREV, S_TRY_SAME, S_A_PREPARE, S_A_PAD, S_AB_UNK, S_AB_SAME, S_AB_CROSS, S_END };
        const FrameSt &f0 = st.f0;
        int state = S_TRY_PREV, stages = 0;
        uint8_t X = ST_ORDER_TFF;       // the order frame A is assumed to have in the current stage
        bool en_sw = true;
        while(state!=S_END)
        {
            stages++;
            if(state==S_TRY_PREV)
            {
                state = S_A_PREPARE;
                if((f0.odd_lines==f1.odd_lines)&&(f0.even_lines==f1.even_lines)&&f0.inner_ok&&f0.outer_ok&&((!f1.order_preset)||(f0.order==f1.order)))
                {
                    f1.inner_silence = f1.outer_silence = f2.inner_silence = f2.outer_silence = true;
                    f2.inner_ok = f2.outer_ok = false; f2.inner_pad = f2.outer_pad = 0;
                    if((f1.odd_lines<ST_MIN_FILL_LINES_PF)&&(f1.even_lines<ST_MIN_FILL_LINES_PF))
                    {
                        f1.set_order_unknown(); f1.inner_ok = f1.outer_ok = false; f1.inner_pad = f1.outer_pad = 0;
                        state = S_END;
                    }
                    else
                    {
                        uint8_t r = SDV_DS_RET_NO_PAD;
                        if(f0.order_set()) r = try_pad(inner_kind(f0.order), f0.inner_pad);
                        if(r==SDV_DS_RET_OK)
                        {
                            f1.std_soft(f0.video_std);
                            f1.order = f0.order; f1.inner_pad = f0.inner_pad; f1.inner_ok = true; f1.inner_silence = false;
                            if(f1.is_tff()) f1.tff_cnt = st.last_pad_counter; else f1.bff_cnt = st.last_pad_counter;
                            X = f1.order; state = S_TRY_SAME;
                        }
                    }
                }
            }
            else if(state==S_TRY_SAME)
            {   // STG_TRY_TFF_TO_TFF / STG_TRY_BFF_TO_BFF: the previous outer padding between frame A and frame B
                uint8_t r = SDV_DS_RET_NO_PAD;
                if(f2.lines(first_even(X))>=ST_MIN_FILL_LINES_PF) r = try_pad(outer_kind(X, X), f0.outer_pad);
                if(r==SDV_DS_RET_OK)
                {
                    f1.outer_pad = f0.outer_pad; f1.outer_ok = true; f2.set_order(X); f1.outer_silence = false;
                    state = S_END;
                }
                else { state = S_AB_SAME; en_sw = false; }
            }
            else if(state==S_A_PREPARE)
            {
                f1.inner_ok = f1.outer_ok = false; f1.inner_pad = f1.outer_pad = 0; f1.tff_cnt = f1.bff_cnt = 0;
                const bool odd_short = f1.odd_lines<ST_MIN_FILL_LINES_PF, even_short = f1.even_lines<ST_MIN_FILL_LINES_PF;
                if(odd_short&&even_short) { if(!f1.order_preset) f1.set_order_unknown(); state = S_END; }
                else if(even_short||odd_short)
                {   // one field only: the order that would need the missing field second is out
                    const uint8_t needs = even_short ? ST_ORDER_TFF : ST_ORDER_BFF;
                    if(f1.order==needs) { f1.outer_ok = false; f1.outer_pad = 0; state = S_END; }
                    else { X = other(needs); state = S_AB_SAME; en_sw = false; }
                }
                else if(f1.order_set()) { X = f1.order; state = S_A_PAD; en_sw = false; }
                else
                {
                    const uint8_t p = probable_order();
                    X = (p==ST_ORDER_BFF) ? ST_ORDER_BFF : ST_ORDER_TFF;
                    state = S_A_PAD; en_sw = true;
                }
            }
            else if(state==S_A_PAD)
            {   // STG_A_PAD_TFF / STG_A_PAD_BFF: padding between the fields of frame A, assuming order X
                f1.inner_pad = 0;
                const uint8_t r = find_pad(inner_kind(X), f1.lines(first_even(X)), &f1.inner_pad);
                if(X==ST_ORDER_TFF) f1.tff_cnt = st.last_pad_counter; else f1.bff_cnt = st.last_pad_counter;
                f1.inner_silence = false;
                if(r==SDV_DS_RET_OK) { f1.set_order(X); f1.inner_ok = true; state = S_AB_SAME; en_sw = false; }
                else if(r==SDV_DS_RET_SILENCE) { f1.inner_silence = f1.outer_silence = true; f1.inner_ok = false; f1.inner_pad = 0; state = S_END; }
                else
                {
                    f1.inner_pad = 0;
                    if(f1.order==X) { f1.inner_ok = false; state = S_AB_SAME; en_sw = false; }
                    else if(en_sw) { X = other(X); en_sw = false; }
                    else state = S_AB_UNK;
                }
            }
            else if(state==S_AB_UNK)
            {
                f1.inner_pad = 0; f1.inner_ok = false; f1.set_order_unknown();
                const uint8_t p = probable_order();
                X = (p==ST_ORDER_BFF) ? ST_ORDER_BFF : ST_ORDER_TFF;
                state = S_AB_SAME; en_sw = true;
            }
            else if(state==S_AB_SAME)
            {   // STG_AB_TFF_TO_TFF / STG_AB_BFF_TO_BFF: frame A and frame B both of order X
                const int fe = first_even(X);
                if((f2.odd_lines<ST_MIN_FILL_LINES_PF)&&(f2.even_lines<ST_MIN_FILL_LINES_PF))
                { f1.outer_pad = 0; f1.outer_ok = false; f2.inner_ok = false; state = S_END; }
                else if(f2.lines(fe)<ST_MIN_FILL_LINES_PF)
                {
                    if(!f1.order_preset) state = S_AB_CROSS;
                    else { f1.outer_pad = 0; f1.outer_ok = false; f2.inner_ok = false; state = S_END; }
                }
                else
                {
                    const uint8_t r = find_pad(outer_kind(X, X), f1.lines(!fe), &f1.outer_pad);
                    f1.outer_silence = false;
                    if(r==SDV_DS_RET_OK)
                    {
                        f1.outer_ok = true; f2.set_order(X); state = S_END;
                        if(!f1.order_set()) f1.set_order(X);
                        else if(f1.order==other(X)) f1.outer_ok = false;
                    }
                    else if(r==SDV_DS_RET_SILENCE) { f1.outer_silence = true; f1.outer_pad = 0; f1.outer_ok = false; state = S_END; }
                    else if(f2.lines(!fe)<ST_MIN_FILL_LINES_PF) { f1.outer_pad = 0; f1.outer_ok = false; f2.inner_ok = false; state = S_END; }
                    else if(!f1.order_preset) state = S_AB_CROSS;
                    else { f1.outer_pad = 0; f1.outer_ok = false; state = S_END; }
                }
            }
            else if(state==S_AB_CROSS)
            {   // STG_AB_TFF_TO_BFF / STG_AB_BFF_TO_TFF: frame A of order X, frame B of the other
                const uint8_t r = find_pad(outer_kind(X, other(X)), f1.lines(!first_even(X)), &f1.outer_pad);
                f1.outer_silence = false;
                if(r==SDV_DS_RET_OK)
                {
                    f1.outer_ok = true; f2.set_order(other(X)); state = S_END;
                    if(!f1.order_set()) f1.set_order(X);
                    else if(f1.order==other(X)) f1.outer_ok = false;
                }
                else if(r==SDV_DS_RET_SILENCE) { f1.outer_silence = true; f1.outer_pad = 0; f1.outer_ok = false; f2.inner_ok = false; state = S_END; }
                else
                {
                    f1.outer_pad = 0; f1.outer_ok = false; f2.inner_ok = false;
                    if(en_sw&&(f1.even_lines>=ST_MIN_FILL_LINES_PF)) { X = other(X); state = S_AB_SAME; en_sw = false; }   // (the reference tests the even field in both twins)
                    else state = S_END;
                }
            }
            if((state!=S_END)&&(stages>14)) break;      // STG_PAD_MAX
        }
    }

    // getAssemblyFieldOrder
    uint8_t assembly_order()
    {
        uint8_t o = ST_ORDER_UNK;
        if(f1.order_set()) { o = f1.order; if(!f1.order_preset) push_order(o); }
        else if(f2.order_preset&&f2.order_set()) o = f2.order;
        else if(st.f0.order_set()&&st.f0.outer_ok) o = st.f0.order;
        if((o!=ST_ORDER_TFF)&&(o!=ST_ORDER_BFF))
        {
            const uint8_t p = probable_order();
            if((p==ST_ORDER_TFF)||(p==ST_ORDER_BFF)) o = p;
            else if(f1.tff_cnt<f1.bff_cnt) o = ST_ORDER_TFF;
            else if(f1.tff_cnt>f1.bff_cnt) o = ST_ORDER_BFF;
            else o = ST_ORDER_TFF;
        }
        if(!f1.order_set()) { f1.order = o; if(!f1.order_preset) f1.order_guessed = true; }
        return o;
    }

    // fillFrameForOutput: the frame's five segments.  The reference's ten branches reduce to: where the lines that are
    // missing to 2 x lines-per-field go (in front, between the fields, behind), or which field loses the excess.
    void fill_frame(const FrameTrim &ta, FrameAsm *out)
    {
        const uint8_t order = assembly_order();
        const bool bff = (order!=ST_ORDER_TFF);
        const FieldTrim &t1 = bff ? ta.even : ta.odd, &t2 = bff ? ta.odd : ta.even;
        if(st.f0.order_set()&&(st.f0.order!=(bff ? ST_ORDER_BFF : ST_ORDER_TFF))) st.f0.outer_ok = false;
        const int T = (f1.video_std==ST_VID_PAL) ? ST_LINES_PF_PAL : ST_LINES_PF_NTSC;     // LINES_PF_DEFAULT = NTSC
        int c1 = t1.data_lines, c2 = t2.data_lines;
        if(c1>T) c1 = T;
        if(c2>T) c2 = T;
        int pre = 0, skip1 = 0, n1 = c1, inner = 0, skip2 = 0, n2 = c2, outer = 0;
        int pre_acct = 0;       // the lines put in front count as inner (1) or outer (2) padding in what the frame remembers
        auto cut = [](int cnt, int excess, int *n) { const uint16_t k = (uint16_t)(cnt-excess); *n = (k<=cnt) ? k : -1; };   // uint16_t wrap: the reference's bounds check then adds nothing
        const int ip = f1.inner_pad, op = f1.outer_pad;
        if(st.f0.outer_ok)
        {
            if(f1.inner_ok&&f1.outer_ok)
            {
                const int tot = c1+c2+ip+op;
                if(tot==2*T) { inner = ip; outer = op; }
                else if(tot<2*T) { inner = ip; outer = op+(2*T-tot); f1.outer_ok = false; }
                else
                {
                    const int t2s = c1+c2+ip;
                    inner = ip;
                    if(2*T>=t2s) outer = 2*T-t2s;
                    else cut(c2, t2s-2*T, &n2);
                    f1.outer_ok = false;
                }
            }
            else if(f1.inner_ok)
            {
                const int t2s = c1+c2+ip;
                inner = ip;
                if(2*T>=t2s) outer = 2*T-t2s;
                else cut(c2, t2s-2*T, &n2);
            }
            else if(f1.outer_ok)
            {
                const int t2s = c1+c2+op;
                outer = op;
                if(2*T>=t2s) inner = 2*T-t2s;
                else { skip2 = t2s-2*T; cut(c2, skip2, &n2); }
            }
            else
            {
                if(2*T>=c1+c2) { inner = T-c1; outer = T-c2; }
                else cut(c2, c1+c2-2*T, &n2);
            }
        }
        else if(f1.inner_ok)
        {
            if(f1.outer_ok)
            {
                const int tot = c1+c2+ip+op;
                inner = ip; outer = op;
                if(2*T>=tot) { pre = 2*T-tot; pre_acct = 1; }
                else { skip1 = tot-2*T; cut(c1, skip1, &n1); }
            }
            else
            {
                const int t2s = c1+c2+ip;
                inner = ip;
                if(2*T>=t2s) outer = 2*T-t2s;
                else cut(c2, t2s-2*T, &n2);
            }
        }
        else if(f1.outer_ok)
        {
            const int t2s = c1+c2+op;
            outer = op;
            if(2*T>=t2s) inner = 2*T-t2s;
            else cut(c1, t2s-2*T, &n1);
        }
        else
        {
            if(2*T>=c1+c2)
            {
                if(set.fix_cut_above&&(c1>0)&&(c2>0))
                {
                    if(bff) { pre = 1; pre_acct = 2; inner = T-(c1+1); outer = T-c2; }
                    else { inner = T-c1+1; outer = T-(c2+1); }
                }
                else { inner = T-c1; outer = T-c2; }
            }
            else
            {   // unreachable with both fields capped at T; kept for the shape of the reference
                if(c1<T) inner = T-c1;
                if(c2<T) outer = T-c2;
            }
        }
        // a cut that wrapped, or a skip beyond the field: addLinesFromField's bounds check fails and adds nothing
        if((n1<0)||(skip1+n1>(int)t1.data_lines)) n1 = 0;
        if((n2<0)||(skip2+n2>(int)t2.data_lines)) n2 = 0;
        if(inner<0) inner = 0;      // (only with the top-line fix on a full field, where the reference queues 65535 lines)
        if(outer<0) outer = 0;
        // line numbers of the empty lines: they continue the count of the data line before them
        FrameAsm fa; memset(&fa, 0, sizeof(fa));
        const int first_num = bff ? 2 : 1, second_num = bff ? 1 : 2;
        int last = first_num;
        fa.line0_pre = (uint16_t)last; last += 2*pre;
        if(n1>0) { const int e = skip1+n1-1, j = t1.first+e+((e>=(int)t1.hole) ? 1 : 0); last = 2*j+1+(bff ? 1 : 0)+2; }
        fa.line0_inner = (uint16_t)last;
        last = second_num;
        if(n2>0) { const int e = skip2+n2-1, j = t2.first+e+((e>=(int)t2.hole) ? 1 : 0); last = 2*j+1+(bff ? 0 : 1)+2; }
        fa.line0_outer = (uint16_t)last;
        fa.pre = (uint16_t)pre; fa.n1 = (uint16_t)n1; fa.inner = (uint16_t)inner; fa.n2 = (uint16_t)n2; fa.outer = (uint16_t)outer;
        fa.skip1 = (uint16_t)skip1; fa.skip2 = (uint16_t)skip2;
        fa.first_even = bff ? 1 : 0;
        fa.first1 = t1.first; fa.first2 = t2.first; fa.hole1 = t1.hole; fa.hole2 = t2.hole;
        fa.total = (uint16_t)(pre+n1+inner+n2+outer);
        fa.mask = 0;
        if(set.mask_seams)
        {
            if((!f1.inner_ok)&&(!f1.inner_silence)) fa.mask |= 1;
            if((!st.f0.outer_ok)&&(!st.f0.outer_silence)) fa.mask |= 2;
        }
        // what the reference remembers of the paddings is what it queued
        f1.inner_pad = (uint16_t)(inner+((pre_acct==1) ? pre : 0));
        f1.outer_pad = (uint16_t)(outer+((pre_acct==2) ? pre : 0));
        *out = fa;
    }

    // One frame of doFrameReassemble: frame A = [frame] with trims ta, frame B with trims tb (all-zero trims behind the last
    // frame of the file).  Returns false when a seam answer was missing (state unchanged, call again once it is there).
    bool step(int frame, const FrameTrim &ta, const FrameTrim &tb, FrameAsm *out, const uint8_t *res4 = 0)
    {
        const StitchCarry saved = st;
        missing = false; cur = frame;
        f1.clear_misc(); f2.clear_misc();
        if(res4) { f1.odd_res = res4[0]; f1.even_res = res4[1]; f2.odd_res = res4[2]; f2.even_res = res4[3]; }
        else f1.odd_res = f1.even_res = f2.odd_res = f2.even_res = set.res16 ? SDV_RES_MODE_16BIT : SDV_RES_MODE_14BIT;
        f1.odd_lines = ta.odd.data_lines; f1.even_lines = ta.even.data_lines;
        f2.odd_lines = tb.odd.data_lines; f2.even_lines = tb.even.data_lines;
        detect_standard(ta, tb);
        find_field_stitching();
        if(missing) { st = saved; return false; }
        fill_frame(ta, out);
        st.f0 = f1;
        return true;
    }
};

// ------------------------------------------------------------------------------------------------ CWD: what the chain kernel is told
// fillNextFieldForCWD for the frame the chain just stepped over ([r] = its final FrameAsmSTC007 state, [tb] = the trims of the
// frame behind it, whose records start at rec_base): the preview lines appended to the queue for the CWD passes.
inline void cwd_next_field(const FrameSt &r, const FrameTrim &tb, size_t rec_base, int H, CwdStep *o)
{
    o->nf_first = 0; o->nf_cnt = 0; o->nf_hole = ST_NO_HOLE; o->pad = 0;
    if(!r.outer_ok||!r.order_set()) return;
    const bool even = !r.is_tff();
    const FieldTrim &t = even ? tb.even : tb.odd;
    o->nf_first = (uint32_t)(rec_base+(even ? H/2 : 0)+t.first);
    o->nf_cnt = (uint16_t)((t.data_lines>112) ? 112 : t.data_lines);
    o->nf_hole = t.hole;
}
// Frames CWD can touch, as chains of consecutive frames.  patch[f] != 0: frame f holds a line CWD may write into (CRC wrong,
// coordinates valid, not forced bad).  A frame is "dirty" when it or the frame before it holds one (the queue still carries the
// last 112 lines of that frame), when the lines handed over by the previous call were patched, or when it follows a dirty frame
// too short to have pushed that frame's predecessors out of the queue.  Everything else CWD leaves exactly as it is.
inline void cwd_plan_chains(const uint8_t *patch, const FrameAsm *fa, int n_done, bool carry_patched, std::vector<int> *chains, std::vector<uint8_t> *dirty)
{
    dirty->assign((size_t)n_done, 0);
    chains->clear();
    for(int f=0;f<n_done;f++)
    {
        bool d = patch[f]!=0;
        if(f==0) d = d||carry_patched;
        else d = d||(patch[f-1]!=0)||((*dirty)[f-1]&&(fa[f-1].total<224));
        (*dirty)[f] = d ? 1 : 0;
        if(d)
        {
            if((f>0)&&(*dirty)[f-1]) (*chains)[chains->size()-1]++;
            else { chains->push_back(f); chains->push_back(1); }
        }
    }
}

}   // namespace sdv
