/* oracle/sdv_oracle_deint.c -- TEST INFRASTRUCTURE ONLY (see sdv_oracle.h).
 *
 * Sequential C restatement of STC007Deinterleaver::processBlock with P/Q correction
 * (stc007deinterleaver.cpp:4-75 matrices, 286-1123 state machine, 1126-1294 fill, 1297-1374 codes,
 * 1376-1464 fixByP, 1468-2048 fixByQ, 2052-2088 multMatrix) and of the STC007DataBlock flag logic
 * (stc007datablock.cpp:53-70, 79-201, 507-562, 565-700).  CWD is out of scope (always off).
 */
#include <string.h>
#include <stdbool.h>
#include "sdv_oracle.h"

enum { W_L0 = 0, W_R0, W_L1, W_R1, W_L2, W_R2, W_P0, W_Q0, W_CNT };
enum { RES_14BIT = 0, RES_16BIT = 1 };
enum { AUD_ORIG = 0, AUD_FIX_P, AUD_FIX_Q, AUD_BROKEN };
enum { RES_MODE_14BIT = 0, RES_MODE_14BIT_AUTO, RES_MODE_16BIT_AUTO, RES_MODE_16BIT };
enum { NO_ERR_INDEX = 64, MAX_PASSES = 3 };
enum { STG_DATA_FILL = 0, STG_ERROR_CHECK, STG_TASK_SELECTION, STG_CWD_CORR, STG_P_CORR, STG_Q_CORR, STG_BAD_BLOCK, STG_NO_CHECK, STG_DATA_OK, STG_CONVERT_MAX };
enum { FIX_NOT_NEED = 0, FIX_SWITCH_P, FIX_BROKEN, FIX_NA, FIX_DONE };

typedef struct
{
    uint16_t words[W_CNT];
    bool line_crc[W_CNT], word_valid[W_CNT];
    uint8_t resolution, audio_state;
} block_t;

/* T = multiplication by x modulo x^14 + x^8 + 1 (TP1_MATRIX rows, stc007deinterleaver.cpp:8-11). */
static uint16_t t_fwd(uint16_t v) { uint16_t fb = (v>>13)&1; return (uint16_t)(((v<<1)&0x3FFF)^fb^(fb<<8)); }
/* T^-1 (TN1_MATRIX rows, stc007deinterleaver.cpp:32-35). */
static uint16_t t_inv(uint16_t v) { uint16_t lb = v&1; return (uint16_t)(((v>>1)^(lb<<13)^(lb<<7))&0x3FFF); }
static uint16_t t_pow(uint16_t v, int k) { v &= 0x3FFF; while(k>0) { v = t_fwd(v); k--; } while(k<0) { v = t_inv(v); k++; } return v; }
/* (T^k + I)^-1 for k=1..5 as bit matrices (TPkIN1_MATRIX, stc007deinterleaver.cpp:56-75), row r = mask of input bits. */
static const uint16_t TPIN1[5][14] =
{
    { 0x3FFE, 0x3FFC, 0x3FF8, 0x3FF0, 0x3FE0, 0x3FC0, 0x3F80, 0x3F00, 0x01FF, 0x03FF, 0x07FF, 0x0FFF, 0x1FFF, 0x3FFF },
    { 0x1554, 0x2AA8, 0x1550, 0x2AA0, 0x1540, 0x2A80, 0x1500, 0x2A00, 0x0155, 0x02AA, 0x0555, 0x0AAA, 0x1555, 0x2AAA },
    { 0x1248, 0x2490, 0x0920, 0x1240, 0x2480, 0x0900, 0x1200, 0x2400, 0x1A49, 0x3492, 0x2924, 0x1249, 0x2492, 0x0924 },
    { 0x0445, 0x088A, 0x1115, 0x222A, 0x0455, 0x08AA, 0x1155, 0x22AA, 0x0111, 0x0222, 0x0444, 0x0888, 0x1111, 0x2222 },
    { 0x1AD7, 0x35AF, 0x2B5E, 0x16BD, 0x2D7B, 0x1AF7, 0x35EF, 0x2BDE, 0x0D6B, 0x1AD6, 0x35AD, 0x2B5A, 0x16B5, 0x2D6B }
};
static uint16_t mult_matrix(const uint16_t *m, uint16_t v)
{   /* multMatrix/bitXOR: parity over the low 14 bits only (stc007deinterleaver.cpp:2052-2088) */
    uint16_t r = 0;
    for(int bit=0;bit<14;bit++) { uint16_t t = (uint16_t)(m[bit]&v&0x3FFF); if(__builtin_parity(t)) r |= (uint16_t)(1<<bit); }
    return r;
}

static uint16_t calc_p(const block_t *b) { return (uint16_t)(b->words[0]^b->words[1]^b->words[2]^b->words[3]^b->words[4]^b->words[5]); }
static uint16_t calc_q(const block_t *b)
{
    uint16_t q = 0;
    for(int i=0;i<6;i++) q ^= t_pow(b->words[i], 6-i);
    return q;
}
static uint16_t synd_p(const block_t *b) { return (uint16_t)(calc_p(b)^b->words[W_P0]); }
static uint16_t synd_q(const block_t *b) { return (uint16_t)(calc_q(b)^b->words[W_Q0]); }

static void blk_set_word(block_t *b, uint8_t i, uint16_t w, bool line_valid) { b->words[i] = w; b->line_crc[i] = b->word_valid[i] = line_valid; }
static void blk_mark_broken(block_t *b)
{
    uint8_t lim = (b->resolution==RES_16BIT) ? W_P0 : W_Q0;
    for(uint8_t i=0;i<=lim;i++) b->word_valid[i] = b->line_crc[i] = false;
    b->audio_state = AUD_BROKEN;
}
static uint8_t blk_err_audio_src(const block_t *b) { uint8_t n = 0; for(int i=0;i<=W_R2;i++) if(!b->line_crc[i]) n++; return n; }
static uint8_t blk_err_total_src(const block_t *b)
{
    uint8_t n = 0, lim = (b->resolution==RES_16BIT) ? W_P0 : W_Q0;
    for(int i=0;i<=lim;i++) if(!b->line_crc[i]) n++;
    return n;
}

static void recalc_p(block_t *b)
{
    uint16_t oldp = b->words[W_P0], p = calc_p(b);
    if(oldp!=p) { blk_set_word(b, W_P0, p, b->line_crc[W_P0]); b->word_valid[W_P0] = true; }
    else b->word_valid[W_P0] = true;
}

static uint8_t fix_by_p(block_t *b, uint8_t first_bad)
{
    uint16_t check;
    b->audio_state = AUD_ORIG;
    check = synd_p(b);
    if(check==0) { if(first_bad!=NO_ERR_INDEX) b->word_valid[first_bad] = true; return FIX_NOT_NEED; }
    else if(first_bad==NO_ERR_INDEX) return FIX_BROKEN;
    else
    {
        uint16_t fix = (uint16_t)(check^b->words[first_bad]);
        blk_set_word(b, first_bad, fix, false);
        b->word_valid[first_bad] = true;
        return FIX_DONE;
    }
}

static uint8_t fix_by_q(block_t *b, uint8_t first_bad, uint8_t second_bad)
{
    bool fix_found = false;
    uint16_t sp = 0, sq, e1 = 0, e2 = 0;
    b->audio_state = AUD_ORIG;
    if(second_bad==NO_ERR_INDEX) if(!b->word_valid[W_P0]) second_bad = W_P0;
    sq = synd_q(b);
    if(second_bad==W_P0)
    {
        if(sq==0)
        {
            if(first_bad!=NO_ERR_INDEX) b->word_valid[first_bad] = true;
            recalc_p(b);
            return FIX_NOT_NEED;
        }
    }
    else
    {
        sp = synd_p(b);
        if((sp==0)&&(sq==0))
        {   /* the reference calls setValid(NO_ERR_INDEX) here when no markers are set: index check makes it a no-op */
            if(first_bad<W_CNT) b->word_valid[first_bad] = true;
            if(second_bad<W_CNT) b->word_valid[second_bad] = true;
            return FIX_NOT_NEED;
        }
    }
    if((second_bad!=W_P0)&&(!b->word_valid[W_P0])) return FIX_NA;
    if(first_bad==NO_ERR_INDEX) return FIX_BROKEN;
    else if(second_bad==NO_ERR_INDEX) return FIX_SWITCH_P;
    if(first_bad<=W_R2)
    {
        if(second_bad==W_P0)
        {   /* one audio word + P: e1 = T^-(6-i) * Sq */
            e1 = t_pow(sq, -(6-first_bad));
            fix_found = true;
        }
        else if((second_bad<=W_R2)&&(second_bad>first_bad))
        {   /* two audio words i<j: e1 = (T^(j-i)+I)^-1 * (T^-(6-j)*Sq ^ Sp), e2 = e1 ^ Sp */
            e1 = t_pow(sq, -(6-second_bad));
            e1 ^= sp;
            e1 = mult_matrix(TPIN1[second_bad-first_bad-1], e1);
            e2 = (uint16_t)(e1^sp);
            fix_found = true;
        }
    }
    if(fix_found)
    {
        uint16_t old1 = b->words[first_bad], old2, fw1 = (uint16_t)(old1^e1), fw2;
        if(e1!=0) { blk_set_word(b, first_bad, fw1, false); b->word_valid[first_bad] = true; }
        else b->word_valid[first_bad] = true;
        old2 = b->words[second_bad];
        if(second_bad==W_P0) e2 = (uint16_t)(old2^calc_p(b));
        fw2 = (uint16_t)(old2^e2);
        if(e2!=0) { blk_set_word(b, second_bad, fw2, false); b->word_valid[second_bad] = true; }
        else b->word_valid[second_bad] = true;
        if((e1==0)&&(e2==0)) return FIX_NOT_NEED;
        return FIX_DONE;
    }
    return FIX_BROKEN;
}

/* words: pointer to line 0 of the block ([n][8] u16), stride 16 lines between block words. */
static void set_word_data(block_t *b, const uint16_t *words, const uint8_t *crc_ok, int s, uint8_t res, bool ignore_crc)
{
    memset(b, 0, sizeof(*b));
    if(res==RES_14BIT)
    {
        for(int k=0;k<8;k++)
        {
            int ln = s+16*k;
            bool ok = ignore_crc ? ((crc_ok[ln]&2)!=0) : ((crc_ok[ln]&1)!=0);
            blk_set_word(b, (uint8_t)k, words[ln*8+k], ok);
        }
    }
    else
    {
        static const uint8_t s_ofs[7] = { 12, 10, 8, 6, 4, 2, 0 };
        for(int k=0;k<7;k++)
        {
            int ln = s+16*k;
            bool ok = ignore_crc ? ((crc_ok[ln]&2)!=0) : ((crc_ok[ln]&1)!=0);      /* word and S-word share the line flag */
            uint16_t f1 = (uint16_t)(words[ln*8+k]<<2);
            uint16_t sw = (uint16_t)((words[ln*8+7]>>s_ofs[k])&0x3);
            blk_set_word(b, (uint8_t)k, (uint16_t)(f1+sw), ok);
        }
        blk_set_word(b, W_Q0, 0, true);
    }
    b->resolution = res;
}

static void process_block(block_t *blk, const uint16_t *words, const uint8_t *crc_ok, int s, int res_mode,
                          bool ignore_crc, bool force_check, bool en_p, bool en_q)
{
    uint8_t run_res, stage_count = 0, fill_passes, all_errs = 0, aud_errs = 0, first_bad = NO_ERR_INDEX, second_bad = NO_ERR_INDEX, fix_result, st;
    if(res_mode==RES_MODE_14BIT) { run_res = RES_14BIT; fill_passes = MAX_PASSES; }
    else if(res_mode==RES_MODE_14BIT_AUTO) { run_res = RES_14BIT; fill_passes = 0; }
    else if(res_mode==RES_MODE_16BIT_AUTO) { run_res = RES_16BIT; fill_passes = 0; }
    else { run_res = RES_16BIT; fill_passes = MAX_PASSES; }
    st = STG_DATA_FILL;
    do
    {
        stage_count++;
        if(st==STG_DATA_FILL)
        {
            set_word_data(blk, words, crc_ok, s, run_res, ignore_crc);
            blk->audio_state = AUD_ORIG;
            fill_passes++;
            st = STG_ERROR_CHECK;
        }
        else if(st==STG_ERROR_CHECK)
        {
            first_bad = second_bad = NO_ERR_INDEX;
            for(uint8_t i=W_L0;i<=W_R2;i++)
                if(!blk->line_crc[i]) { if(first_bad==NO_ERR_INDEX) first_bad = i; else if(second_bad==NO_ERR_INDEX) { second_bad = i; break; } }
            aud_errs = blk_err_audio_src(blk);
            all_errs = blk_err_total_src(blk);
            st = STG_TASK_SELECTION;
        }
        else if(st==STG_TASK_SELECTION)
        {
            st = STG_BAD_BLOCK;
            if(all_errs<=2)
            {
                if(aud_errs==0)
                {
                    if(!force_check) st = STG_DATA_OK;
                    else if(en_p) st = STG_P_CORR;
                    else st = STG_NO_CHECK;
                }
                else if(aud_errs==1) { if(en_p) st = STG_P_CORR; }
                else if(aud_errs==2) { if(run_res==RES_14BIT) { if(en_q) st = STG_Q_CORR; } }
            }
        }
        else if(st==STG_P_CORR)
        {
            st = STG_BAD_BLOCK;
            if(blk->word_valid[W_P0])
            {
                fix_result = fix_by_p(blk, first_bad);
                if(fix_result==FIX_BROKEN) blk_mark_broken(blk);
                else
                {
                    st = STG_DATA_OK;
                    if(fix_result==FIX_DONE) blk->audio_state = AUD_FIX_P;
                    else if(fix_result==FIX_NOT_NEED) { if(first_bad<W_P0) blk->audio_state = AUD_FIX_P; }
                    if((run_res==RES_14BIT)&&en_q)
                    {
                        if(blk->word_valid[W_Q0])
                        {
                            if(force_check) { if(synd_q(blk)!=0) { st = STG_BAD_BLOCK; blk_mark_broken(blk); } }
                        }
                        else
                        {
                            uint16_t q = calc_q(blk);
                            if(blk->words[W_Q0]!=q) { blk_set_word(blk, W_Q0, q, blk->line_crc[W_Q0]); blk->word_valid[W_Q0] = true; }
                            else blk->word_valid[W_Q0] = true;
                        }
                    }
                }
            }
            else
            {
                if(run_res==RES_14BIT)
                {
                    if(en_q) st = STG_Q_CORR;
                    else if(aud_errs==0) st = STG_NO_CHECK;
                }
                else if(aud_errs==0) st = STG_NO_CHECK;
            }
        }
        else if(st==STG_Q_CORR)
        {
            st = STG_BAD_BLOCK;
            if(blk->word_valid[W_Q0])
            {
                fix_result = fix_by_q(blk, first_bad, second_bad);
                if(!blk->line_crc[W_P0]) second_bad = W_P0;
                if(fix_result==FIX_DONE) { st = STG_DATA_OK; blk->audio_state = AUD_FIX_Q; }
                else if(fix_result==FIX_NOT_NEED) { st = STG_DATA_OK; if(first_bad<W_P0) blk->audio_state = AUD_FIX_Q; }
                else if(fix_result==FIX_SWITCH_P) st = STG_P_CORR;
                else if(fix_result==FIX_BROKEN) blk_mark_broken(blk);
            }
            else if(first_bad==NO_ERR_INDEX)
            {
                st = STG_NO_CHECK;
                blk_set_word(blk, W_P0, calc_p(blk), false); blk->word_valid[W_P0] = true;
                blk_set_word(blk, W_Q0, calc_q(blk), false); blk->word_valid[W_Q0] = true;
            }
        }
        else if(st==STG_BAD_BLOCK)
        {
            if(fill_passes>=MAX_PASSES) break;
            run_res = (run_res==RES_16BIT) ? RES_14BIT : RES_16BIT;
            st = STG_DATA_FILL;
        }
        else break;     /* STG_NO_CHECK, STG_DATA_OK */
        if(stage_count>(STG_CONVERT_MAX*MAX_PASSES)) break;
    }
    while(1);
}

int sdvo_deint_stc007(const uint16_t *words, const uint8_t *crc_ok, int n, int res_mode,
                      int ignore_crc, int force_check, int p_corr, int q_corr, sdvo_block_rec *out)
{
    int nb = 0;
    for(int s=0;s+112<n;s++)
    {
        block_t b;
        sdvo_block_rec *r = &out[nb++];
        process_block(&b, words, crc_ok, s, res_mode, ignore_crc!=0, force_check!=0, p_corr!=0, q_corr!=0);
        memset(r, 0, sizeof(*r));
        uint8_t audio_bad = 0; bool silent = true;
        for(int i=0;i<8;i++)
        {
            r->words[i] = b.words[i];
            if(b.line_crc[i]) r->line_crc |= (uint8_t)(1<<i);
            if(b.word_valid[i]) r->word_valid |= (uint8_t)(1<<i);
        }
        for(int i=0;i<6;i++)
        {
            if(!b.word_valid[i]) audio_bad++;
            r->samples[i] = (b.resolution==RES_16BIT) ? (int16_t)b.words[i] : (int16_t)(uint16_t)(b.words[i]<<2);
            if(r->samples[i]!=0) silent = false;
        }
        r->audio_state = b.audio_state; r->resolution = b.resolution;
        if(audio_bad==0) r->flags |= 1;
        if(b.audio_state==AUD_BROKEN) r->flags |= 2;
        if(b.audio_state==AUD_FIX_P) r->flags |= 4;
        if(b.audio_state==AUD_FIX_Q) r->flags |= 8;
        if(silent) r->flags |= 16;
    }
    return nb;
}

/* STC007DataStitcher::tryPadding (stc007datastitcher.cpp:1417-1740): the seam of two fields with [padding] empty lines
 * between them is deinterleaved block by block (forced parity check) and the bursts of valid / silent / unchecked /
 * BROKEN blocks are counted.  out [n_pad][6] = index, valid, silent, unchecked, broken, return code
 * (0 NO_DATA, 1 SILENCE, 2 BROKE, 3 NO_PAD, 4 OK: stc007datastitcher.h:209-216). */
enum { MAX_BURST_SILENCE = 8, MAX_BURST_BROKEN = 1, SEAM_LINES = 112+8 };
int sdvo_try_padding(const uint16_t *w1, const uint8_t *ok1, int n1, const uint16_t *w2, const uint8_t *ok2, int n2,
                     int n_pad, int res_mode, int ignore_crc, int p_corr, int q_corr, int lim14, int lim16, uint16_t *out)
{
    static uint16_t qw[(2*SEAM_LINES+64)*8]; static uint8_t qok[2*SEAM_LINES+64];
    for(int pad=0;pad<n_pad;pad++)
    {
        uint16_t *o = out+pad*6;
        memset(o, 0, 6*sizeof(uint16_t));
        int n = 0;
        int start1 = (n1>(SEAM_LINES-pad)) ? (n1-(SEAM_LINES-pad)) : 0;
        for(int i=start1;i<n1;i++) { memcpy(qw+n*8, w1+i*8, 16); qok[n] = ok1[i]; n++; }
        for(int i=0;i<pad;i++) { memset(qw+n*8, 0, 16); qok[n] = 0; n++; }
        int cnt2 = (n2>SEAM_LINES) ? SEAM_LINES : n2;
        for(int i=0;i<cnt2;i++) { memcpy(qw+n*8, w2+i*8, 16); qok[n] = ok2[i]; n++; }
        /* fewer than 112 lines: DS_RET_NO_DATA and the caller's FieldStitchStats stays as clear() left it (frametrimset.cpp:
         * 374-378).  Exactly 112 lines: no block fits, the run counters are all zero (-> NO_PAD); whether the statistics are
         * written depends on an uninitialised flag in the reference (run_lock, stc007datastitcher.cpp:1424/1563) -- the
         * compiled reference writes them, so does this. */
        if(n<112) { o[2] = o[3] = o[4] = 0xFF; o[5] = 0; continue; }
        int valid_cnt = 0, silence_cnt = 0, uncheck_cnt = 0, broken_cnt = 0, valid_max = 0, silence_max = 0, uncheck_max = 0;
        int lim = q_corr ? lim14 : lim16;
        for(int s=0;s+112<n;s++)
        {
            block_t b;
            process_block(&b, qw, qok, s, res_mode, ignore_crc!=0, true, p_corr!=0, q_corr!=0);
            bool bvalid = true, silent = true, broken = b.audio_state==AUD_BROKEN;
            for(int i=0;i<6;i++)
            {
                if(!b.word_valid[i]) bvalid = false;
                int16_t smp = (b.resolution==RES_16BIT) ? (int16_t)b.words[i] : (int16_t)(uint16_t)(b.words[i]<<2);
                if(smp!=0) silent = false;
            }
            int errs = 0, limw = (b.resolution==RES_16BIT) ? W_P0 : W_Q0;
            for(int i=0;i<=limw;i++) if(!b.line_crc[i]) errs++;
            bool can_force = (!broken)&&((b.resolution==RES_14BIT) ? (errs<=1) : (errs==0));
            if(bvalid&&(!silent)&&can_force) valid_cnt++;
            else if(valid_cnt>valid_max) valid_max = valid_cnt;
            if(silent) { silence_cnt++; if(silence_cnt>=MAX_BURST_SILENCE) valid_cnt = 0; }
            else { if(silence_cnt>silence_max) silence_max = silence_cnt; silence_cnt = 0; }
            bool unch = q_corr ? ((!can_force)||(b.audio_state==AUD_FIX_Q)) : (b.audio_state==AUD_FIX_P);
            if(unch) { uncheck_cnt++; if(uncheck_cnt>=lim) valid_cnt = 0; }
            else { if(uncheck_cnt>uncheck_max) uncheck_max = uncheck_cnt; uncheck_cnt = 0; }
            if(broken) { broken_cnt++; if(broken_cnt>=MAX_BURST_BROKEN) valid_cnt = 0; }
        }
        if(valid_cnt>valid_max) valid_max = valid_cnt;
        if(silence_cnt>silence_max) silence_max = silence_cnt;
        if(uncheck_cnt>uncheck_max) uncheck_max = uncheck_cnt;
        o[0] = (uint16_t)pad; o[1] = (uint16_t)valid_max; o[2] = (uint16_t)silence_max; o[3] = (uint16_t)uncheck_max; o[4] = (uint16_t)broken_cnt;
        if(broken_cnt>=MAX_BURST_BROKEN) o[5] = 2;
        else if(silence_max>MAX_BURST_SILENCE) o[5] = 1;
        else if(uncheck_max>lim) o[5] = 3;
        else if(valid_max==0) o[5] = 3;
        else o[5] = 4;
    }
    return n_pad;
}


/* STC007DataStitcher::findPadding (stc007datastitcher.cpp:1743-2054): sweep the paddings (stopping early once a
 * padding without BROKEN blocks has been followed by one with), order the runs with FieldStitchStats::operator<
 * (frametrimset.cpp:312-371; untouched entries keep FieldStitchStats::clear()'s values, 374-378) and accept the best
 * one only if it stands out.  video_std: 1 PAL / 2 NTSC (frametrimset.h VID_*).  out[0] = padding, out[1] = DS_RET_* code,
 * out[2] = last_pad_counter. */
typedef struct { uint16_t index, valid, silent, unchecked, broken; } fss_t;
static int fss_less(const fss_t *a, const fss_t *b)
{
    if(a->broken!=b->broken) return a->broken<b->broken;
    if(a->valid!=b->valid) return a->valid>b->valid;
    if(a->unchecked!=b->unchecked) return a->unchecked<b->unchecked;
    if(a->silent!=b->silent) return a->silent<b->silent;
    return a->index<b->index;
}
static void fss_sort(fss_t *v, int n)
{   /* the order is total up to identical entries: any sort gives the reference's std::sort result */
    for(int i=1;i<n;i++)
    {
        fss_t t = v[i]; int j = i;
        while((j>0)&&fss_less(&t, &v[j-1])) { v[j] = v[j-1]; j--; }
        v[j] = t;
    }
}
int sdvo_find_padding(const uint16_t *w1, const uint8_t *ok1, int n1, const uint16_t *w2, const uint8_t *ok2, int n2,
                      int video_std, int resolution_16bit, int res_mode, int ignore_crc, int p_corr, int q_corr,
                      int lim14, int lim16, uint16_t *out)
{
    enum { MAX_PAD_14 = 32, MAX_PAD_16 = 16, UNCH_DELTA = 8, RET_SILENCE = 1, RET_NO_PAD = 3, RET_OK = 4 };
    int padding = 0, res = RET_NO_PAD, last_cnt = 0xFF;
    const int lpf = (video_std==1) ? 294 : ((video_std==2) ? 245 : 0);
    if(lpf) padding = (n1>lpf) ? 0 : (lpf-n1);
    int max_padding = MAX_PAD_14, lim = lim14&0xFF;
    if(resolution_16bit||!q_corr) { max_padding = MAX_PAD_16; lim = lim16&0xFF; }
    if(p_corr||q_corr)
    {
        fss_t sd[MAX_PAD_14];
        for(int i=0;i<max_padding;i++) { sd[i].index = sd[i].valid = 0; sd[i].silent = sd[i].unchecked = sd[i].broken = 0xFF; }
        int min_broken = 0xFFFF, no_brk = 0;
        for(int pad=0;pad<max_padding;pad++)
        {
            uint16_t o[6*MAX_PAD_14];
            /* tryPadding of one padding = the last row of a sweep 0..pad */
            sdvo_try_padding(w1, ok1, n1, w2, ok2, n2, pad+1, res_mode, ignore_crc, p_corr, q_corr, lim14, lim16, o);
            const uint16_t *r = o+6*pad;
            sd[pad].index = r[0]; sd[pad].valid = r[1]; sd[pad].silent = r[2]; sd[pad].unchecked = r[3]; sd[pad].broken = r[4];
            if(min_broken>sd[pad].broken) { min_broken = sd[pad].broken; if(min_broken==0) no_brk = pad; }
            else if(min_broken==0)
            {
                if((sd[no_brk].valid>0)&&(sd[no_brk].unchecked<lim)&&(sd[pad].broken>0)) break;
            }
        }
        fss_sort(sd, max_padding);
        last_cnt = sd[0].broken&0xFF;
        if(sd[0].silent<MAX_BURST_SILENCE)
        {
            if(sd[0].unchecked<lim)
            {
                if((sd[0].broken<2)&&(sd[0].broken<sd[1].broken)) { res = RET_OK; padding = sd[0].index; }
                else if((((int16_t)sd[0].valid-(int16_t)sd[1].valid)>UNCH_DELTA)&&(sd[0].broken==0)) { res = RET_OK; padding = sd[0].index; }
            }
            else
            {
                for(int pad=0;pad<max_padding;pad++)
                {
                    sd[pad].broken = (uint16_t)min_broken;
                    if(sd[pad].unchecked>=lim) sd[pad].broken = 0xFF;
                }
                fss_sort(sd, max_padding);
                if(sd[0].unchecked<lim)
                {
                    if(((int16_t)sd[0].valid-(int16_t)sd[1].valid)>UNCH_DELTA) { res = RET_OK; padding = sd[0].index; }
                }
            }
        }
        else res = RET_SILENCE;
    }
    out[0] = (uint16_t)padding; out[1] = (uint16_t)res; out[2] = (uint16_t)last_cnt;
    return res;
}
