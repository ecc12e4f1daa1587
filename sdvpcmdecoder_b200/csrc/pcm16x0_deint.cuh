// pcm16x0_deint.cuh -- PCM-16x0 (SI format) deinterleave + P correction, one thread per data block.
//
// PCM16X0Deinterleaver::processBlock / setWordData / fixByP (pcm16x0deinterleaver.cpp:128-912) with the PCM16X0DataBlock
// state it works on (pcm16x0datablock.cpp:98-470, 476-800; word <-> line map 1029-1157): a data block takes sub-lines
// i, i+35 and i+70 of a 105 sub-line interleave block (LINE_1, LINE_2 = parity, LINE_3); word j of each sub-line goes
// to sub-block j; L/R swap lines with the block's order (odd/even) and the sub-block; one erased word per sub-block is
// rebuilt from P = L xor R; words whose bits were guessed by the Binarizer's bit picker are distrusted when the parity
// check fails.  Block state lives in registers: 9 words and bit masks over (sub-block, line).
// Sample flags as PCM16X0DataStitcher::outputDataBlock computes them (pcm16x0datastitcher.cpp:4998-5085).
#pragma once
#include "sdv_common.cuh"

namespace sdv {

enum { X0_LINE_1 = 0, X0_LINE_2, X0_LINE_3 };
enum { X0_WORD_L = 0, X0_WORD_R, X0_WORD_P };
enum { X0_AUD_ORIG = 0, X0_AUD_FIX_P, X0_AUD_BROKEN };
enum { X0_STG_CRC_CHECK = 0, X0_STG_P_CORR, X0_STG_BAD_BLOCK, X0_STG_NO_CHECK, X0_STG_DATA_OK, X0_STG_CONVERT_MAX };
enum { X0_FIX_NOT_NEED = 0, X0_FIX_BROKEN, X0_FIX_DONE, X0_NO_ERR = 64 };
enum { X0_SUBLINES_ITL = 105, X0_BLOCKS_ITL = 35, X0_OFS = 35 };              // SI format (pcm16x0datablock.h:40,67-69)
enum { X0_SUBLINES_EI = 1470, X0_BLOCKS_EI = 490, X0_OFS_EI = 490 };          // EI format: one unit = one frame (41,70-72)

struct X0Block
{
    u16 words[3][3];            // [sub-block][line]
    u32 crc, valid;             // bit (3*sub-block + line)
    u32 picked_left, picked_crc;    // bit per line
    u8 state[3];
    bool order_even;
};
SDV_HD int x0_line(const X0Block *b, int blk, int word)
{   // getWordToLine
    if(word==X0_WORD_P) return X0_LINE_2;
    const bool l_on_3 = (blk==1) ? b->order_even : !b->order_even;
    if(word==X0_WORD_L) return l_on_3 ? X0_LINE_3 : X0_LINE_1;
    return l_on_3 ? X0_LINE_1 : X0_LINE_3;
}
SDV_HD u16 x0_get(const X0Block *b, int blk, int line)
{
    u16 r = 0;
    for(int l=0;l<3;l++) r = (l==line) ? b->words[blk][l] : r;
    return r;
}
SDV_HD void x0_fix_word(X0Block *b, int blk, int word, u16 w)
{
    const int line = x0_line(b, blk, word);
    for(int l=0;l<3;l++) b->words[blk][l] = (l==line) ? w : b->words[blk][l];
    b->valid |= 1u<<(3*blk+line);
}
SDV_HD void x0_mark_bad(X0Block *b, int blk, int line) { const u32 m = 1u<<(3*blk+line); b->crc &= ~m; b->valid &= ~m; b->picked_left &= ~(1u<<line); }
SDV_HD void x0_mark_broken(X0Block *b, int blk)
{
    for(int i=0;i<3;i++) if((blk>=3)||(i==blk)) { const u32 m = 7u<<(3*i); b->crc &= ~m; b->valid &= ~m; b->state[i] = X0_AUD_BROKEN; }
}
SDV_HD bool x0_picked_sample(const X0Block *b, int blk, int word) { return (blk==0) ? (((b->picked_left>>x0_line(b, blk, word))&1u)!=0) : false; }
SDV_HD int x0_picked_audio(const X0Block *b, int blk) { return (blk==0) ? ((x0_picked_sample(b, 0, X0_WORD_L) ? 1 : 0)+(x0_picked_sample(b, 0, X0_WORD_R) ? 1 : 0)) : 0; }
SDV_HD bool x0_picked_parity(const X0Block *b, int blk) { if((blk==0)&&((b->picked_left>>X0_LINE_2)&1u)) return true; return ((b->picked_crc>>X0_LINE_2)&1u)!=0; }
SDV_HD int x0_fix_by_p(X0Block *b, int blk, int bad_ptr, u16 mask)
{
    const u16 check = (u16)(b->words[blk][0]^b->words[blk][1]^b->words[blk][2]);
    if(check==0) { if(bad_ptr!=X0_NO_ERR) x0_fix_word(b, blk, bad_ptr, x0_get(b, blk, x0_line(b, blk, bad_ptr))); return X0_FIX_NOT_NEED; }
    if(bad_ptr==X0_NO_ERR) return X0_FIX_BROKEN;
    if((mask&check)==0) { x0_fix_word(b, blk, bad_ptr, (u16)(check^x0_get(b, blk, x0_line(b, blk, bad_ptr)))); return X0_FIX_DONE; }
    return X0_FIX_BROKEN;
}

struct X0Cfg { u8 ignore_crc, force_check, p_corr; };

// One data block.  sub[l] = the sub-line on LINE_(l+1).
SDV_HD void x0_process_block(X0Block *b, const sdv_pcm16x0_subline *s1, const sdv_pcm16x0_subline *s2, const sdv_pcm16x0_subline *s3,
                             bool even_order, X0Cfg cfg)
{
    const sdv_pcm16x0_subline *s[3] = { s1, s2, s3 };
    b->crc = b->valid = b->picked_left = b->picked_crc = 0;
    b->order_even = even_order;
    int pick_cnt = 0;
    for(int line=0;line<3;line++)
    {
        const bool ok = cfg.ignore_crc ? ((s[line]->flags&SDV_X0F_HAS_DATA)!=0) : ((s[line]->flags&SDV_X0F_CRC_OK)!=0);
        for(int sb=0;sb<3;sb++) { b->words[sb][line] = s[line]->words[sb]; if(ok) { b->crc |= 1u<<(3*sb+line); b->valid |= 1u<<(3*sb+line); } }
        if(s[line]->picked_left) b->picked_left |= 1u<<line;
        if(s[line]->flags&SDV_X0F_PICKED_RIGHT) b->picked_crc |= 1u<<line;
        pick_cnt += s[line]->picked_left;
    }
    pick_cnt &= 0xFF;
    for(int blk=0;blk<3;blk++)
    {
        b->state[blk] = X0_AUD_ORIG;
    }
    for(int blk=0;blk<3;blk++)
    {
        int st = X0_STG_CRC_CHECK, stage_count = 0;
        const u32 c3 = (b->crc>>(3*blk))&7u;
        const int err_total = 3-(int)((c3&1u)+((c3>>1)&1u)+((c3>>2)&1u));
        const int err_audio = 2-(int)((c3&1u)+((c3>>2)&1u));
        u16 pick_mask = 0;
        for(;;)
        {
            stage_count++;
            if(st==X0_STG_CRC_CHECK)
            {
                if(err_total>1) st = X0_STG_BAD_BLOCK;
                else if(cfg.p_corr)
                {
                    if(cfg.force_check) st = X0_STG_P_CORR;
                    else if(err_total>0) st = (err_audio>0) ? X0_STG_P_CORR : X0_STG_DATA_OK;
                    else st = X0_STG_DATA_OK;
                }
                else
                {
                    if(err_audio>0) st = X0_STG_BAD_BLOCK;
                    else if(cfg.force_check) st = X0_STG_NO_CHECK;
                    else st = X0_STG_DATA_OK;
                }
            }
            else if(st==X0_STG_P_CORR)
            {
                int bad_ptr = X0_NO_ERR;
                if(!((b->crc>>(3*blk+x0_line(b, blk, X0_WORD_L)))&1u)) bad_ptr = X0_WORD_L;
                else if(!((b->crc>>(3*blk+x0_line(b, blk, X0_WORD_R)))&1u)) bad_ptr = X0_WORD_R;
                else if(!((b->crc>>(3*blk+X0_LINE_2))&1u)) bad_ptr = X0_WORD_P;
                if(bad_ptr!=X0_WORD_P)
                {
                    const int fix_result = x0_fix_by_p(b, blk, bad_ptr, pick_mask);
                    if(fix_result==X0_FIX_BROKEN)
                    {
                        const int pa = x0_picked_audio(b, blk);
                        if(pa>1) { x0_mark_bad(b, blk, X0_LINE_1); x0_mark_bad(b, blk, X0_LINE_3); st = X0_STG_BAD_BLOCK; }
                        else if(pa==1)
                        {
                            if(x0_picked_parity(b, blk)) { x0_mark_bad(b, blk, X0_LINE_1); x0_mark_bad(b, blk, X0_LINE_3); st = X0_STG_BAD_BLOCK; }
                            else
                            {
                                if((b->picked_left>>X0_LINE_1)&1u) { x0_mark_bad(b, blk, X0_LINE_1); st = X0_STG_P_CORR; }
                                else if((b->picked_left>>X0_LINE_3)&1u) { x0_mark_bad(b, blk, X0_LINE_3); st = X0_STG_P_CORR; }
                                else { st = X0_STG_BAD_BLOCK; x0_mark_broken(b, 3); }
                                if(pick_cnt>0) { pick_mask = (u16)(16-pick_cnt); pick_mask = (u16)(1<<pick_mask); pick_mask--; }
                            }
                        }
                        else
                        {
                            if(x0_picked_parity(b, blk)) { x0_mark_bad(b, blk, X0_LINE_2); st = X0_STG_NO_CHECK; }
                            else { st = X0_STG_BAD_BLOCK; x0_mark_broken(b, blk); }
                        }
                    }
                    else if(fix_result==X0_FIX_NOT_NEED) st = X0_STG_DATA_OK;
                    else { st = X0_STG_DATA_OK; b->state[blk] = X0_AUD_FIX_P; }
                }
                else st = X0_STG_NO_CHECK;
            }
            else break;
            if(stage_count>X0_STG_CONVERT_MAX) break;
        }
    }
}

// 3 sample pairs + flags + audio states of one processed block.
SDV_HD void x0_output(const X0Block *b, i16 *smp /*[6]*/, u8 *fl /*[6]*/, u8 *st /*[3] or NULL*/)
{
    const bool all_valid = ((b->valid&0x5u)==0x5u)&&(((b->valid>>3)&0x5u)==0x5u)&&(((b->valid>>6)&0x5u)==0x5u);     // LINE_1 and LINE_3 of every sub-block
    for(int sb=0;sb<3;sb++)
    {
        const bool broken = b->state[sb]==X0_AUD_BROKEN;
        const bool bstate = (!broken)&&all_valid;
        for(int ch=0;ch<2;ch++)
        {
            const int line = x0_line(b, sb, ch ? X0_WORD_R : X0_WORD_L);
            const bool v = (!broken)&&(((b->valid>>(3*sb+line))&1u)!=0);
            const bool fx = bstate&&(((b->crc>>(3*sb+line))&1u)!=0);
            smp[2*sb+ch] = (i16)x0_get(b, sb, line);
            fl[2*sb+ch] = (u8)((bstate ? SDV_SF_BLOCK_OK : 0)|(v ? SDV_SF_WORD_VALID : 0)|(fx ? SDV_SF_WORD_FIXED : 0));
        }
        if(st) st[sb] = b->state[sb];
    }
}

#if defined(__CUDACC__)
__global__ void __launch_bounds__(256) pcm16x0_deint_kernel(const sdv_pcm16x0_subline *sub, long long n_blocks, X0Cfg cfg, int ei,
                                                            i16 *samples, u8 *sflags, u8 *states)
{
    const long long b = (long long)blockIdx.x*blockDim.x+threadIdx.x;
    if(b>=n_blocks) return;
    const int per = ei ? X0_BLOCKS_EI : X0_BLOCKS_ITL, unit = ei ? X0_SUBLINES_EI : X0_SUBLINES_ITL, ofs = ei ? X0_OFS_EI : X0_OFS;
    const long long m = b/per; const int i = (int)(b-m*per);
    const sdv_pcm16x0_subline *base = sub+m*unit+i;
    const sdv_pcm16x0_subline s1 = base[0], s2 = base[ofs], s3 = base[2*ofs];
    X0Block blk;
    x0_process_block(&blk, &s1, &s2, &s3, (i&1)!=0, cfg);
    i16 smp[6]; u8 fl[6]; u8 st[3];
    x0_output(&blk, smp, fl, st);
    u32 *d = (u32 *)(samples+b*6);
    d[0] = (u32)(u16)smp[0]|((u32)(u16)smp[1]<<16); d[1] = (u32)(u16)smp[2]|((u32)(u16)smp[3]<<16); d[2] = (u32)(u16)smp[4]|((u32)(u16)smp[5]<<16);
    if(sflags) { u16 *f = (u16 *)(sflags+b*6); f[0] = (u16)(fl[0]|(fl[1]<<8)); f[1] = (u16)(fl[2]|(fl[3]<<8)); f[2] = (u16)(fl[4]|(fl[5]<<8)); }
    if(states) { states[b*3] = st[0]; states[b*3+1] = st[1]; states[b*3+2] = st[2]; }
}
#endif

}   // namespace sdv
