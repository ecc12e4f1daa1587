/* oracle/sdv_oracle_pcm16x0.c -- TEST INFRASTRUCTURE ONLY (see sdv_oracle.h).
 *
 * Sequential C restatement of the PCM-16x0 (SI format) deinterleaver: PCM16X0Deinterleaver::processBlock / setWordData /
 * fixByP (pcm16x0deinterleaver.cpp:128-912), the PCM16X0DataBlock state it manipulates (pcm16x0datablock.cpp:98-470,
 * 476-800, word <-> line map 1029-1157) and the sample flags of PCM16X0DataStitcher::outputDataBlock
 * (pcm16x0datastitcher.cpp:4998-5085), driven per interleave block as performDeinterleave does (5204-5346).
 */
#include <string.h>
#include <stdbool.h>
#include "sdv_oracle.h"

enum { LINE_1 = 0, LINE_2, LINE_3 };
enum { WORD_L = 0, WORD_R, WORD_P };
enum { AUD_ORIG = 0, AUD_FIX_P, AUD_BROKEN };
enum { STG_CRC_CHECK = 0, STG_P_CORR, STG_BAD_BLOCK, STG_NO_CHECK, STG_DATA_OK, STG_CONVERT_MAX };
enum { FIX_NOT_NEED = 0, FIX_BROKEN, FIX_DONE };
enum { NO_ERR_INDEX = 64 };

typedef struct
{
    uint16_t words[3][3];
    bool word_crc[3][3], word_valid[3][3];
    bool picked_left[3], picked_crc[3];
    uint8_t audio_state[3];
    bool order_even;
} blk_t;

/* pcm16x0datablock.cpp:1029-1157 */
static int word_to_line(const blk_t *b, int blk, int word)
{
    if(word==WORD_P) return LINE_2;
    bool l_on_line3 = (blk==1) ? b->order_even : !b->order_even;       /* sub-blocks 1 and 3: L on LINE_3 in odd order */
    if(word==WORD_L) return l_on_line3 ? LINE_3 : LINE_1;
    return l_on_line3 ? LINE_1 : LINE_3;
}
static bool crc_ok(const blk_t *b, int blk, int word) { return b->word_crc[blk][word_to_line(b, blk, word)]; }
static bool valid(const blk_t *b, int blk, int word) { return b->word_valid[blk][word_to_line(b, blk, word)]; }
static uint16_t get_word(const blk_t *b, int blk, int word) { return b->words[blk][word_to_line(b, blk, word)]; }
static void fix_word(blk_t *b, int blk, int word, uint16_t w) { int l = word_to_line(b, blk, word); b->words[blk][l] = w; b->word_valid[blk][l] = true; }
static void mark_bad(blk_t *b, int blk, int line) { b->word_crc[blk][line] = false; b->word_valid[blk][line] = false; b->picked_left[line] = false; }
static void mark_broken(blk_t *b, int blk)
{
    for(int i=0;i<3;i++)
        if((blk>=3)||(i==blk)) { for(int l=0;l<3;l++) { b->word_valid[i][l] = false; b->word_crc[i][l] = false; } b->audio_state[i] = AUD_BROKEN; }
}
static bool picked_sample(const blk_t *b, int blk, int word) { return (blk==0) ? b->picked_left[word_to_line(b, blk, word)] : false; }
static int picked_audio(const blk_t *b, int blk) { return (blk==0) ? ((picked_sample(b, 0, WORD_L) ? 1 : 0)+(picked_sample(b, 0, WORD_R) ? 1 : 0)) : 0; }
static bool picked_parity(const blk_t *b, int blk) { if((blk==0)&&b->picked_left[LINE_2]) return true; return b->picked_crc[LINE_2]; }

static int fix_by_p(blk_t *b, int blk, int bad_ptr, uint16_t mask)
{
    uint16_t check = (uint16_t)(b->words[blk][0]^b->words[blk][1]^b->words[blk][2]);
    if(check==0) { if(bad_ptr!=NO_ERR_INDEX) fix_word(b, blk, bad_ptr, get_word(b, blk, bad_ptr)); return FIX_NOT_NEED; }
    if(bad_ptr==NO_ERR_INDEX) return FIX_BROKEN;
    if((mask&check)==0) { fix_word(b, blk, bad_ptr, (uint16_t)(check^get_word(b, blk, bad_ptr))); return FIX_DONE; }
    return FIX_BROKEN;
}

static void process_block(blk_t *b, const uint16_t *w1, const uint16_t *w2, const uint16_t *w3, const bool ok[3],
                          const bool pl[3], const bool pr[3], int pick_cnt, bool even_order, bool force, bool en_p)
{
    const uint16_t *w[3] = { w1, w2, w3 };
    memset(b, 0, sizeof(*b));
    b->order_even = even_order;
    for(int line=0;line<3;line++)
    {
        for(int sb=0;sb<3;sb++) { b->words[sb][line] = w[line][sb]; b->word_crc[sb][line] = b->word_valid[sb][line] = ok[line]; }
        b->picked_left[line] = pl[line]; b->picked_crc[line] = pr[line];
    }
    for(int blk=0;blk<3;blk++)
    {
        int st = STG_CRC_CHECK, stage_count = 0, bad_ptr, fix_result;
        int err_total = 0, err_audio = 0;
        uint16_t pick_mask = 0;
        for(int l=0;l<3;l++) if(!b->word_crc[blk][l]) err_total++;
        if(!b->word_crc[blk][LINE_1]) err_audio++;
        if(!b->word_crc[blk][LINE_3]) err_audio++;
        do
        {
            stage_count++;
            if(st==STG_CRC_CHECK)
            {
                if(err_total>1) st = STG_BAD_BLOCK;
                else if(en_p)
                {
                    if(force) st = STG_P_CORR;
                    else if(err_total>0) st = (err_audio>0) ? STG_P_CORR : STG_DATA_OK;
                    else st = STG_DATA_OK;
                }
                else
                {
                    if(err_audio>0) st = STG_BAD_BLOCK;
                    else if(force) st = STG_NO_CHECK;
                    else st = STG_DATA_OK;
                }
            }
            else if(st==STG_P_CORR)
            {
                bad_ptr = NO_ERR_INDEX;
                if(!crc_ok(b, blk, WORD_L)) bad_ptr = WORD_L;
                else if(!crc_ok(b, blk, WORD_R)) bad_ptr = WORD_R;
                else if(!crc_ok(b, blk, WORD_P)) bad_ptr = WORD_P;
                if(bad_ptr!=WORD_P)
                {
                    fix_result = fix_by_p(b, blk, bad_ptr, pick_mask);
                    if(fix_result==FIX_BROKEN)
                    {
                        if(picked_audio(b, blk)>1) { mark_bad(b, blk, LINE_1); mark_bad(b, blk, LINE_3); st = STG_BAD_BLOCK; }
                        else if(picked_audio(b, blk)==1)
                        {
                            if(picked_parity(b, blk)) { mark_bad(b, blk, LINE_1); mark_bad(b, blk, LINE_3); st = STG_BAD_BLOCK; }
                            else
                            {
                                if(b->picked_left[LINE_1]) { mark_bad(b, blk, LINE_1); st = STG_P_CORR; }
                                else if(b->picked_left[LINE_3]) { mark_bad(b, blk, LINE_3); st = STG_P_CORR; }
                                else { st = STG_BAD_BLOCK; mark_broken(b, 3); }
                                if(pick_cnt>0) { pick_mask = (uint16_t)(16-pick_cnt); pick_mask = (uint16_t)(1<<pick_mask); pick_mask--; }
                            }
                        }
                        else
                        {
                            if(picked_parity(b, blk)) { mark_bad(b, blk, LINE_2); st = STG_NO_CHECK; }
                            else { st = STG_BAD_BLOCK; mark_broken(b, blk); }
                        }
                    }
                    else if(fix_result==FIX_NOT_NEED) st = STG_DATA_OK;
                    else { st = STG_DATA_OK; b->audio_state[blk] = AUD_FIX_P; }
                }
                else st = STG_NO_CHECK;
            }
            else break;
            if(stage_count>STG_CONVERT_MAX) break;
        }
        while(1);
    }
}

static int deint_pcm16x0_fmt(const uint16_t *words, const uint8_t *flags, const uint8_t *picked_left, int n_itl, int ignore_crc,
                            int force_check, int p_corr, int ei, int16_t *out_samples, uint8_t *out_flags, uint8_t *out_state)
{
    /* SI: units of 105 sub-lines, 35 data blocks, line offset 35; EI: units of 1470 (one frame), 490 blocks, offset 490
       (pcm16x0datablock.h:40-41,67-72; pcm16x0deinterleaver.cpp:223-236) */
    const int unit = ei ? 1470 : 105, nblk = ei ? 490 : 35, lofs = ei ? 490 : 35;
    int o = 0;
    for(int m=0;m<n_itl;m++)
    {
        for(int i=0;i<nblk;i++)
        {
            blk_t b;
            bool ok[3], pl[3], pr[3];
            int pick_cnt = 0;
            const uint16_t *w[3];
            for(int l=0;l<3;l++)
            {
                size_t k = (size_t)m*unit+i+(size_t)lofs*l;
                w[l] = words+3*k;
                ok[l] = ignore_crc ? ((flags[k]&2)!=0) : ((flags[k]&1)!=0);
                pl[l] = picked_left[k]!=0; pr[l] = (flags[k]&8)!=0;
                pick_cnt += picked_left[k];
            }
            process_block(&b, w[0], w[1], w[2], ok, pl, pr, (uint8_t)pick_cnt, (i&1)!=0, force_check!=0, p_corr!=0);
            bool all_valid = true;
            for(int sb=0;sb<3;sb++) if(!b.word_valid[sb][LINE_1]||!b.word_valid[sb][LINE_3]) all_valid = false;
            for(int sb=0;sb<3;sb++)
            {
                bool broken = b.audio_state[sb]==AUD_BROKEN;
                bool bstate = (!broken)&&all_valid;
                for(int ch=0;ch<2;ch++)
                {
                    int word = ch ? WORD_R : WORD_L;
                    bool v = (!broken)&&valid(&b, sb, word);
                    bool fx = bstate&&crc_ok(&b, sb, word);
                    out_samples[o*6+2*sb+ch] = (int16_t)get_word(&b, sb, word);
                    out_flags[o*6+2*sb+ch] = (uint8_t)((bstate ? 1 : 0)|(v ? 2 : 0)|(fx ? 4 : 0));
                }
                out_state[o*3+sb] = b.audio_state[sb];
            }
            o++;
        }
    }
    return o;
}

int sdvo_deint_pcm16x0(const uint16_t *words, const uint8_t *flags, const uint8_t *picked_left, int n_itl, int ignore_crc,
                       int force_check, int p_corr, int16_t *out_samples, uint8_t *out_flags, uint8_t *out_state)
{
    return deint_pcm16x0_fmt(words, flags, picked_left, n_itl, ignore_crc, force_check, p_corr, 0, out_samples, out_flags, out_state);
}
int sdvo_deint_pcm16x0_ei(const uint16_t *words, const uint8_t *flags, const uint8_t *picked_left, int n_units, int ignore_crc,
                          int force_check, int p_corr, int16_t *out_samples, uint8_t *out_flags, uint8_t *out_state)
{
    return deint_pcm16x0_fmt(words, flags, picked_left, n_units, ignore_crc, force_check, p_corr, 1, out_samples, out_flags, out_state);
}
