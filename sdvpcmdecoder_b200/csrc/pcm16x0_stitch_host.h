// pcm16x0_stitch_host.h -- the decisions of PCM16X0DataStitcher's SI and EI padding searches (host code of the library).
//
//   findEIFrameStitching pcm16x0datastitcher.cpp:3588-4117   -> X0PadChain::frame_ei
//   findEIPadding                              2649-2994   -> X0PadChain::find_ei_padding
//   conditionEIFramePadding                    2997-3464   -> X0PadChain::condition_ei
//   findEIDataAlignment                        3467-3585   -> X0PadChain::find_ei_data_alignment
//   findSIPadding        pcm16x0datastitcher.cpp:1557-2245   -> X0PadChain::find_si_padding
//   findSIDataAlignment                        2246-2377   -> X0PadChain::frame
//   getProbablePadding / updatePadStats        4355-4423   -> X0PadChain::probable / push
// What they decide over -- trySIPadding for the paddings 0..34, the offset of the zeroed control bits, the interleave block
// the field starts in -- is computed per field on the device (x0_sipad_scan_cta, pcm16x0_stitch.cuh); what is left is a few
// comparisons per field and the 65-field history of accepted paddings, which makes every field depend on the ones before.
#pragma once
#include <stdint.h>
#include <string.h>
#include <algorithm>
#include "pcm16x0_stitch.cuh"

namespace sdv {

struct X0PadChain
{
    enum { DEPTH = 65, INVALID = 0xFF };
    uint8_t hist[DEPTH]; int pos;
    bool p_corr;

    void reset() { memset(hist, INVALID, sizeof(hist)); pos = 0; }
    void push(uint8_t pad) { hist[pos] = pad; pos = (pos+1)%DEPTH; }
    uint8_t probable() const
    {
        int cnt[81]; memset(cnt, 0, sizeof(cnt));
        int total = 0;
        for(int i=0;i<DEPTH;i++) if((hist[i]!=INVALID)&&(hist[i]<81)) { cnt[hist[i]]++; total++; }
        if(!total) return INVALID;
        int best = 0, idx = INVALID;
        for(int i=0;i<81;i++) if(cnt[i]>best) { best = cnt[i]; idx = i; }
        return (uint8_t)idx;
    }
    static bool stat_less(const sdv_stitch_stats &a, const sdv_stitch_stats &b)
    {   // FieldStitchStats::operator<
        if(a.broken!=b.broken) return a.broken<b.broken;
        if(a.valid!=b.valid) return a.valid>b.valid;
        if(a.unchecked!=b.unchecked) return a.unchecked<b.unchecked;
        if(a.silent!=b.silent) return a.silent<b.silent;
        return a.index<b.index;
    }
    // findSIPadding for one field.  Returns the DS_RET_* code; geo = top padding, lines cut at the head, lines kept.
    uint8_t find_si_padding(const X0PadScan &sc, X0FieldGeo *geo)
    {
        int f_size = sc.n_sub, cut_lines = 0;
        uint8_t res = SDV_DS_RET_NO_PAD;
        int top_padding = (X0S_SUBLINES_PF-f_size)/3;
        geo->top_pad = (int16_t)top_padding; geo->cut = 0; geo->lines = (int16_t)(f_size/3); geo->pad = 0;
        if(f_size<X0S_MIN_FILL_SI) return SDV_DS_RET_NO_DATA;
        int pad_top = 0, pad_bottom = X0S_SUBLINES_PF-f_size;
        bool lock = false;
        const int zero_ofs = sc.zero_ofs, iblk = sc.iblk_num;
        if(p_corr)
        {
            const uint8_t pad = probable();
            if(pad!=INVALID)
            {
                if((pad<X0S_MAX_PAD_SI)&&(sc.st[pad].result==SDV_DS_RET_OK))
                {
                    lock = true; push(pad); pad_top = pad;
                    pad_bottom = (pad_bottom>=pad_top) ? (pad_bottom-pad_top) : 0;
                    res = SDV_DS_RET_OK;
                }
                else pad_bottom = X0S_SUBLINES_PF-f_size;
            }
            if(!lock)
            {
                int min_broken = sc.st[0].broken;
                for(int p=0;p<X0S_MAX_PAD_SI;p++) if(sc.st[p].broken<min_broken) min_broken = sc.st[p].broken;
                sdv_stitch_stats cand[X0S_MAX_PAD_SI]; int n = 0;
                for(int p=0;p<X0S_MAX_PAD_SI;p++) if((sc.st[p].broken==min_broken)&&(sc.st[p].valid>0)) cand[n++] = sc.st[p];
                if(n>0)
                {
                    std::sort(cand, cand+n, stat_less);
                    if(cand[0].unchecked<=X0S_BURST_SI)
                    {
                        if(cand[0].silent<X0S_BURST_SI)
                        {
                            if(min_broken==0) res = (cand[0].valid>X0S_MIN_VALID_SI) ? SDV_DS_RET_OK : SDV_DS_RET_NO_PAD;
                            else res = SDV_DS_RET_BROKE;
                            lock = true; pad_top = cand[0].index;
                            pad_bottom = (pad_bottom>=pad_top) ? (pad_bottom-pad_top) : 0;
                            push((uint8_t)pad_top);
                        }
                        else res = SDV_DS_RET_SILENCE;
                    }
                }
            }
        }
        if(lock)
        {
            int last_ofs = iblk*X0_BLOCKS_ITL;
            if(last_ofs<pad_top)
            {
                last_ofs = (iblk+1)*X0_BLOCKS_ITL-pad_top;
                pad_top = 0;
                cut_lines += last_ofs; f_size -= 3*last_ofs;       // cutFieldTop
            }
            else if(last_ofs>pad_top) pad_top += (iblk-1)*X0_BLOCKS_ITL;
            pad_top *= 3;
            pad_bottom = X0S_SUBLINES_PF-pad_top;
            if(pad_bottom>=f_size) pad_bottom -= f_size;
            else { pad_bottom = f_size-pad_bottom; f_size -= pad_bottom; pad_bottom = 0; }
        }
        else if(zero_ofs>=0)
        {
            pad_top = pad_bottom = 0;
            int last_ofs = 3+iblk*X0_SUBLINES_ITL-zero_ofs;
            if(last_ofs>0) pad_top = last_ofs;
            else if(last_ofs<0) { const int c = (-last_ofs)/3; cut_lines += c; f_size -= 3*c; }
            last_ofs = X0S_SUBLINES_PF-(pad_top+f_size);
            if(last_ofs>0) pad_bottom = last_ofs;
            else if(last_ofs<0) f_size -= -last_ofs;
        }
        else { pad_bottom = 0; pad_top = X0S_SUBLINES_PF-f_size; }
        if(f_size<0) f_size = 0;
        geo->top_pad = (int16_t)(pad_top/3); geo->cut = (int16_t)cut_lines; geo->lines = (int16_t)(f_size/3);
        return res;
    }
    // findSIDataAlignment for one frame: geo[0] odd field, geo[1] even field; returns mask_seams (padding not OK, not silence).
    bool frame(const X0PadScan &odd, const X0PadScan &even, X0FieldGeo *geo, uint8_t *results /*[2], may be NULL*/)
    {
        const uint8_t ro = find_si_padding(odd, &geo[0]);
        bool padding_ok = (ro==SDV_DS_RET_OK), silence = (!padding_ok)&&(ro==SDV_DS_RET_SILENCE);
        const uint8_t re = find_si_padding(even, &geo[1]);
        if(re!=SDV_DS_RET_OK)
        {
            padding_ok = false;
            if(ro==SDV_DS_RET_SILENCE) silence = true;         // (the reference tests the odd field's result here too)
        }
        if(results) { results[0] = ro; results[1] = re; }
        return (!padding_ok)&&(!silence);
    }

    // ------------------------------------------------------------------------------------------------ EI format
    // One field of the frame while findEIFrameStitching works on it: sub-lines kept, lines cut at the head, paddings in lines.
    struct EIField { int size, cut, top, bottom; };
    // findEIPadding (2649-2994) over the statistics tryEIPadding left for the 81 paddings.  Returns DS_RET_*; *pad = the padding
    // it locked on (-1: none).
    uint8_t find_ei_padding(const X0EIScan &sc, int *pad)
    {
        *pad = -1;
        if(!p_corr) return SDV_DS_RET_NO_PAD;
        uint8_t res = SDV_DS_RET_NO_PAD;
        int min_broken = sc.st[0].broken;
        for(int p=0;p<X0S_MAX_PAD_EI;p++) if(sc.st[p].broken<min_broken) min_broken = sc.st[p].broken;
        sdv_stitch_stats cand[X0S_MAX_PAD_EI]; int n = 0;
        for(int p=0;p<X0S_MAX_PAD_EI;p++) if((sc.st[p].broken==min_broken)&&(sc.st[p].valid>0)) cand[n++] = sc.st[p];
        if(n==0) return res;
        std::sort(cand, cand+n, stat_less);
        if(cand[0].unchecked>X0S_BURST_EI) return res;
        if(cand[0].silent>=X0S_BURST_EI) return SDV_DS_RET_SILENCE;
        if(min_broken==0) res = (cand[0].valid>X0S_MIN_VALID_EI) ? SDV_DS_RET_OK : SDV_DS_RET_NO_PAD;
        else res = SDV_DS_RET_BROKE;
        *pad = cand[0].index;
        push((uint8_t)cand[0].index);
        return res;
    }
    // conditionEIFramePadding (2997-3464): the padding between the fields is known (a.bottom), spread it and the rest of the
    // 2 x 245 lines over the four paddings by the position of the zeroed control bits.  a: first field of the frame, b: second.
    static void condition_ei(EIField &a, EIField &b, int zero_a, int zero_b, int iblk_b)
    {
        bool pos_lock = false;
        const int inter = a.bottom;
        int zero_ofs = zero_b, last_ofs;
        if(zero_ofs>=0)
        {
            pos_lock = true;
            zero_ofs = b.size-zero_ofs;
            last_ofs = (X0_BLOCKS_ITL-2)*3-zero_ofs;
            if(last_ofs<0) b.size -= -last_ofs;
            else if(last_ofs>0) b.bottom += last_ofs/3;
            last_ofs = (7-iblk_b-1)*X0_SUBLINES_ITL;
            b.bottom += last_ofs/3;
            last_ofs = X0S_LINES_PF-b.size/3;
            last_ofs -= b.bottom;
            if(last_ofs<0)
            {
                last_ofs = -last_ofs;
                zero_ofs = (last_ofs/X0_BLOCKS_ITL+1)*X0_BLOCKS_ITL;
                last_ofs = b.bottom-zero_ofs;
                if(last_ofs<0) { b.top = b.bottom = 0; pos_lock = false; }
                else
                {
                    b.bottom = last_ofs;
                    last_ofs = X0S_LINES_PF-b.size/3;
                    last_ofs -= b.bottom;
                }
            }
            if(last_ofs>inter)
            {
                if((last_ofs-inter)<2) { b.top = inter; b.bottom += last_ofs-inter; }
                else { b.top = b.bottom = 0; pos_lock = false; }
            }
            else if(pos_lock) b.top = last_ofs;
        }
        if(pos_lock)
        {
            a.bottom = inter-b.top;
            zero_ofs = (a.size+b.size)/3+a.bottom+b.top+b.bottom;
            zero_ofs = 2*X0S_LINES_PF-zero_ofs;
            if(zero_ofs<0) { a.top = a.bottom = b.top = b.bottom = 0; pos_lock = false; }
            else a.top = zero_ofs;
        }
        if(!pos_lock)
        {
            zero_ofs = zero_a;
            if(zero_ofs>=0)
            {
                pos_lock = true;
                zero_ofs -= (zero_ofs/X0_SUBLINES_ITL)*X0_SUBLINES_ITL;
                zero_ofs = (X0S_SUBLINES_PF+2*3-zero_ofs)/3;
                a.top = zero_ofs;
                zero_ofs = X0S_LINES_PF-a.top-a.size/3;
                if(zero_ofs<0) pos_lock = false;
                else
                {
                    a.bottom = zero_ofs;
                    zero_ofs = inter-a.bottom;
                    if(zero_ofs<0) pos_lock = false;
                    else
                    {
                        b.top = zero_ofs;
                        zero_ofs = X0S_LINES_PF-(b.size/3+b.top);
                        if(zero_ofs<0) { b.bottom = 0; b.size -= (-zero_ofs)*3; }
                        else b.bottom = zero_ofs;
                    }
                }
            }
        }
        if(!pos_lock)
        {
            b.top = inter/2;
            a.bottom = (inter*3-b.top*3)/3;
            zero_ofs = X0S_LINES_PF-(a.size/3+a.bottom);
            if(zero_ofs<0)
            {
                a.top = 0;
                a.bottom = X0S_LINES_PF-a.size/3;
                b.top = inter-a.bottom;
            }
            else a.top = zero_ofs;
            zero_ofs = X0S_LINES_PF-(b.size/3+b.top);
            if(zero_ofs<0) { b.bottom = 0; b.size -= (-zero_ofs)*3; }
            else b.bottom = zero_ofs;
        }
    }
    // findEIDataAlignment (3467-3585): one field placed by its zeroed control bits alone.
    static uint8_t find_ei_data_alignment(EIField &f, int zero_ofs, int iblk)
    {
        if(zero_ofs<0) return SDV_DS_RET_NO_PAD;
        f.top = f.bottom = 0;
        zero_ofs = f.size-zero_ofs;
        int last_ofs = (X0_BLOCKS_ITL-2)*3-zero_ofs;
        if(last_ofs<0) f.size -= -last_ofs;
        else if(last_ofs>0) f.bottom += last_ofs/3;
        last_ofs = (7-iblk-1)*X0_SUBLINES_ITL;
        f.bottom += last_ofs/3;
        last_ofs = X0S_LINES_PF-f.size/3;
        last_ofs -= f.bottom;
        if(last_ofs<0)
        {
            last_ofs = -last_ofs;
            if((last_ofs<X0_BLOCKS_ITL)&&(last_ofs<f.size)) { f.cut += last_ofs; f.size -= 3*last_ofs; return SDV_DS_RET_OK; }    // cutFieldTop
            return SDV_DS_RET_NO_PAD;
        }
        f.top += last_ofs;
        return SDV_DS_RET_OK;
    }
    // findEIFrameStitching for one frame.  geo[0] odd field, geo[1] even field (geo.pad = bottom padding); returns mask_seams
    // (padding not OK, frame not silent).  result (may be NULL): [0] the DS_RET_* of the padding search (SDV_DS_RET_OK when the
    // padding of the history held), [1] the padding between the fields or 0xFF.
    bool frame_ei(const X0EIScan &sc, bool bff, X0FieldGeo *geo, uint8_t *result)
    {
        const int f1 = bff ? 1 : 0, f2 = 1-f1;
        EIField fld[2];
        for(int k=0;k<2;k++) { fld[k].size = sc.n_sub[k]; fld[k].cut = 0; fld[k].top = 0; fld[k].bottom = 0; }
        bool padding_ok = false, silence = false, aligned = false;
        uint8_t res = SDV_DS_RET_NO_PAD; int inter = -1;
        // STG_TRY_PREVIOUS
        const uint8_t prob = probable();
        if((prob!=INVALID)&&(sc.st[prob].result==SDV_DS_RET_OK))
        {
            push(prob);
            fld[f1].bottom = prob; fld[f2].top = 0;
            inter = prob; res = SDV_DS_RET_OK; aligned = true;
        }
        else
        {   // STG_FULL_PREPARE
            const int odd = fld[0].size, even = fld[1].size;
            const bool too_few = ((odd<X0S_MIN_FILL_EI)&&(even<X0S_MIN_FILL_EI))||((odd+even)<2*X0S_MIN_FILL_EI);
            if((!too_few)&&(fld[f1].size>=X0S_MIN_FILL_EI))
            {   // STG_INTERPAD_TFF / STG_INTERPAD_BFF
                fld[0].top = (X0S_SUBLINES_PF-odd)/3; fld[1].top = (X0S_SUBLINES_PF-even)/3;
                res = find_ei_padding(sc, &inter);
                if(inter>=0) { fld[f1].bottom = inter; fld[f2].top = 0; }
                if(res==SDV_DS_RET_OK) aligned = true;
                else { if(res==SDV_DS_RET_SILENCE) silence = true; fld[f1].bottom = 0; }
            }
        }
        if(aligned)
        {   // STG_ALIGN_TFF / STG_ALIGN_BFF
            condition_ei(fld[f1], fld[f2], sc.zero_ofs[f1], sc.zero_ofs[f2], sc.iblk[f2]);
            padding_ok = true;
        }
        else
        {   // STG_FB_CTRL_EST
            for(int k=0;k<2;k++)
                if(find_ei_data_alignment(fld[k], sc.zero_ofs[k], sc.iblk[k])!=SDV_DS_RET_OK)
                {
                    fld[k].bottom = 0;
                    fld[k].top = (X0S_SUBLINES_PF-fld[k].size)/3;
                }
        }
        for(int k=0;k<2;k++)
        {
            if(fld[k].size<0) fld[k].size = 0;
            geo[k].top_pad = (int16_t)fld[k].top; geo[k].cut = (int16_t)fld[k].cut; geo[k].lines = (int16_t)(fld[k].size/3); geo[k].pad = (int16_t)fld[k].bottom;
        }
        if(result) { result[0] = res; result[1] = (uint8_t)((inter>=0) ? inter : INVALID); }
        return (!padding_ok)&&(!silence);
    }
};

}   // namespace sdv
