// pcm16x0_stitch_host.h -- the decisions of PCM16X0DataStitcher's SI padding search (host code of the library).
//
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
};

}   // namespace sdv
