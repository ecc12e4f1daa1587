/* oracle/sdv_oracle_stc007.c -- TEST INFRASTRUCTURE ONLY (see sdv_oracle.h).
 *
 * Sequential C restatement of the reference's STC-007 line decode:
 *   Binarizer::processLine and everything below it (binarizer.cpp:443-1724, 1771-2383, 2385-3548, 3551-4120,
 *   5275-5595, 6047-6113, 7322-7445, 7560-8055), PCMLine / STC007Line arithmetic (pcmline.cpp:223-311,461-519,
 *   stc007line.cpp:69-98,245-341,493-504,582-606), and the inter-line chain of VideoToDigital::doBinarize
 *   (videotodigital.cpp:698-1815) for the STC-007 format with default fine settings (binarizer.cpp:48-65).
 */
#include <string.h>
#include <stdbool.h>
#include <stdlib.h>
#include "sdv_oracle.h"

/* ------------------------------------------------------------------------------------------- constants */
enum { NO_COORD_LEFT = -32768, NO_COORD_RIGHT = 32767 };
enum { INT_CALC_MULT = 128 };
enum { BITS_PCM_DATA = 128, BITS_IN_LINE = 137, BITS_BETWEEN = 132, PS_STAGES = 5 };
enum { HYST_DEPTH_MIN = 0, HYST_DEPTH_SAFE = 4, HYST_DEPTH_MAX = 10, SHIFT_MIN = 0, SHIFT_SAFE = 2, SHIFT_MAX = 4 };
enum { MAX_COLL_CRCS = 32 };
enum { REF_NO_PCM = 0, REF_BAD_CRC, REF_CRC_COLL, REF_CRC_OK };
enum { SPAN_NOT_FOUND = 0, SPAN_TOO_NARROW, SPAN_OK };
enum { STG_INPUT_ALL = 0, STG_INPUT_LEVEL, STG_REF_FIND, STG_REF_SWEEP_RUN, STG_READ_PCM, STG_DATA_OK, STG_NO_GOOD, STG_MAX };
enum { MARK_ST_START = 0, MARK_ST_TOP_1, MARK_ST_BOT_1, MARK_ST_TOP_2, MARK_ST_BOT_2 };
enum { MARK_ED_START = 0, MARK_ED_TOP, MARK_ED_BOT, MARK_ED_LEN_OK };
enum { SRV_NO = 0, SRV_CTRL_BLOCK = 7 };
static const int8_t PIX_SHIFT[PS_STAGES] = { 0, 1, -1, 2, -2 };   /* pcmline.h:60-71 (BG == ED tables) */

/* Fine settings, bin_preset_t::reset (binarizer.cpp:48-65). */
enum { MAX_BLACK_LVL = 160, MIN_WHITE_LVL = 28, MIN_CONTRAST = 10, MIN_REF_LVL = 7, MAX_REF_LVL = 240,
       MIN_VALID_CRCS = 5, MARK_MAX_DIST = 6 };

typedef struct { int16_t start, stop; uint8_t reference; } coord_t;

static coord_t coord_none(void) { coord_t c; c.start = NO_COORD_LEFT; c.stop = NO_COORD_RIGHT; c.reference = 0; return c; }
static bool coord_valid(coord_t c) { return (c.start!=NO_COORD_LEFT)&&(c.stop!=NO_COORD_RIGHT)&&(c.start<c.stop); }
static bool coord_set(coord_t *c, int16_t s, int16_t e) { if(e>s) { c->start = s; c->stop = e; return true; } return false; }
static bool coord_ne(coord_t a, coord_t b) { return (a.start!=b.start)||(a.stop!=b.stop); }
/* CoordinatePair::operator< (frametrimset.cpp:63-98). */
static bool coord_less(coord_t a, coord_t b)
{
    if(a.start<b.start) return true;
    if(a.start==b.start)
    {
        if(a.stop>b.stop) return true;
        if(a.stop==b.stop) return a.reference<b.reference;
    }
    return false;
}

typedef struct { uint8_t result; uint16_t crc; uint8_t hyst_dph, shift_stg; int16_t data_start, data_stop; } crc_handler_t;

/* STC007Line + PCMLine state (pcmline.h:132-160, stc007line.h:154-166). */
typedef struct
{
    uint32_t frame_number; uint16_t line_number;
    uint8_t black_level, white_level, ref_low, ref_level, ref_high;
    coord_t coords;
    uint8_t hysteresis_depth, shift_stage;
    bool ref_level_sweeped, data_by_ext_tune;
    uint16_t calc_crc;
    bool blk_wht_set, coords_set, forced_bad;
    uint8_t service_type;
    uint16_t pixel_start, pixel_stop; int16_t pixel_start_offset; uint32_t pixel_size_mult, halfpixel_size_mult;
    uint8_t mark_st_stage, mark_ed_stage;
    uint16_t marker_start_bg, marker_start_ed, marker_stop_ed;
    uint16_t pix[PS_STAGES][BITS_PCM_DATA];
    bool word_crc[9], word_valid[9];
    uint16_t words[9];
} stc_line;

/* Binarizer state (binarizer.h:277-305). */
typedef struct
{
    const uint8_t *px; uint16_t line_length;
    uint8_t in_def_black, in_def_white, in_def_reference; coord_t in_def_coord;
    uint8_t in_max_hysteresis_depth, in_max_shift_stages;
    bool do_coord_search, do_ref_lvl_sweep;
    uint8_t bin_mode;
    uint8_t hysteresis_depth_lim, shift_stages_lim;
    uint16_t scan_start, scan_end, mark_start_max, mark_end_min, estimated_ppb;
    bool was_BW_scanned;
    crc_handler_t shift_crcs[SHIFT_MAX+1], hyst_crcs[HYST_DEPTH_MAX+1], crc_stats[MAX_COLL_CRCS+1];
} binarizer;

/* ------------------------------------------------------------------------------------------- CRC */
static uint16_t crc16_update(uint16_t crc, uint16_t data, uint8_t bits)
{
    for(uint8_t i=0;i<bits;i++)
    {
        bool in = (data&(1u<<(bits-1)))!=0;
        bool msb = (crc&0x8000)!=0;
        crc = (uint16_t)(crc<<1);
        if(in!=msb) crc ^= 0x1021;
        data = (uint16_t)(data<<1);
    }
    return crc;
}
uint16_t sdvo_crc_stc007(const uint16_t *w8) { uint16_t c = 0xFFFF; for(int i=0;i<8;i++) c = crc16_update(c, w8[i], 14); return c; }
uint16_t sdvo_crc_pcm1(const uint16_t *w6)
{   /* CRC over inverted 13-bit words, result inverted (pcm1line.cpp:158-171). */
    uint16_t c = 0xFFFF; for(int i=0;i<6;i++) c = crc16_update(c, (uint16_t)((~w6[i])&0x1FFF), 13); return (uint16_t)~c;
}
uint16_t sdvo_crc_pcm16x0(const uint16_t *w3) { uint16_t c = 0xFFFF; for(int i=0;i<3;i++) c = crc16_update(c, w3[i], 16); return c; }

/* ------------------------------------------------------------------------------------------- line object */
static void line_calc_crc(stc_line *l) { l->calc_crc = sdvo_crc_stc007(l->words); }
static void line_set_invalid_crc(stc_line *l) { l->words[8] = (uint16_t)~l->calc_crc; }
static bool line_crc_ok_ign(const stc_line *l) { return l->calc_crc==l->words[8]; }
static bool line_crc_ok(const stc_line *l) { return (!l->forced_bad)&&line_crc_ok_ign(l); }
static bool line_has_start(const stc_line *l) { return l->mark_st_stage==MARK_ST_BOT_2; }
static bool line_has_stop(const stc_line *l) { return l->mark_ed_stage==MARK_ED_LEN_OK; }
static bool line_has_markers(const stc_line *l) { return line_has_start(l)&&line_has_stop(l); }

static void line_clear(stc_line *l)
{   /* STC007Line::clear (stc007line.cpp:69-98) */
    memset(l, 0, sizeof(*l));
    l->coords = coord_none();
    l->pixel_stop = 1;
    l->pixel_size_mult = INT_CALC_MULT;
    l->halfpixel_size_mult = INT_CALC_MULT/2;
    l->calc_crc = 0xA96A;
    line_set_invalid_crc(l);
}

/* PCMLine::clear (pcmline.cpp:94-112) is NOT virtual: called through a PCMLine pointer it resets only the base
   fields and leaves words[], marker stages/coordinates and the pixel table of the STC007Line untouched. */
static void line_base_clear(stc_line *l)
{
    l->frame_number = 0; l->line_number = 0;
    l->black_level = l->white_level = 0;
    l->ref_low = l->ref_level = l->ref_high = 0;
    l->coords = coord_none();
    l->hysteresis_depth = l->shift_stage = 0;
    l->ref_level_sweeped = l->data_by_ext_tune = false;
    l->calc_crc = 0;
    l->blk_wht_set = l->coords_set = l->forced_bad = false;
    l->service_type = SRV_NO;
    l->pixel_start = 0; l->pixel_stop = 1; l->pixel_start_offset = 0;
    l->pixel_size_mult = INT_CALC_MULT; l->halfpixel_size_mult = INT_CALC_MULT/2;
}

static void line_set_source_pixels(stc_line *l, uint16_t s, uint16_t e)
{
    if(e>s) { if(BITS_BETWEEN<=(e-s)) { l->pixel_start = s; l->pixel_stop = e; } }
}

static void line_apply_crc_per_word(stc_line *l)
{
    bool ok = line_crc_ok(l);
    for(int i=0;i<9;i++) l->word_crc[i] = l->word_valid[i] = ok;
}

static uint16_t line_pixel_calc(const stc_line *l, uint8_t pcm_bit, uint8_t shift_stage, uint8_t bit_ofs)
{   /* PCMLine::getVideoPixeBylCalc (pcmline.cpp:249-311) */
    int32_t vp;
    pcm_bit = (uint8_t)(pcm_bit+bit_ofs);
    if(pcm_bit>=BITS_IN_LINE) pcm_bit = BITS_IN_LINE-1;
    vp = (int32_t)((pcm_bit*l->pixel_size_mult)+l->halfpixel_size_mult);
    vp = vp/INT_CALC_MULT;
    vp = vp+l->pixel_start_offset;
    vp += PIX_SHIFT[shift_stage];
    if(vp<l->pixel_start) vp = l->pixel_start;
    else if(vp>=l->pixel_stop) vp = l->pixel_stop-1;
    return (uint16_t)vp;
}

static void line_calc_ppb(stc_line *l, coord_t c)
{   /* PCMLine::calcPPB / setPPB (pcmline.cpp:223-234,506-519), STC007Line::calcCoordinates (stc007line.cpp:1051) */
    l->pixel_size_mult = (uint32_t)(c.stop-c.start);
    l->pixel_size_mult = (l->pixel_size_mult*INT_CALC_MULT+BITS_BETWEEN/2)/BITS_BETWEEN;
    l->pixel_start_offset = c.start;
    l->halfpixel_size_mult = (l->pixel_size_mult+1)/2;
    for(uint8_t s=0;s<PS_STAGES;s++)
        for(uint8_t b=0;b<BITS_PCM_DATA;b++) l->pix[s][b] = line_pixel_calc(l, b, s, 3);
}

static uint8_t line_get_ppb(const stc_line *l) { return (uint8_t)(l->pixel_size_mult/INT_CALC_MULT); }

static bool line_has_control_block(const stc_line *l)
{   /* stc007line.cpp:493-504 */
    return (l->words[0]==0x3333)&&(l->words[1]==0x0CCC)&&(l->words[2]==0x3333)&&(l->words[3]==0x0CCC)
           &&(l->words[4]==0x0000)&&((l->words[7]&0x0FF0)==0x0000);
}

static void line_set_serv_ctrl_blk(stc_line *l)
{   /* stc007line.cpp:101-131 */
    uint16_t w4 = l->words[4], w5 = l->words[5], w6 = l->words[6], w7 = l->words[7];
    uint32_t fr = l->frame_number; uint16_t ln = l->line_number;
    line_clear(l);
    l->frame_number = fr; l->line_number = ln;
    l->words[4] = w4; l->words[5] = w5; l->words[6] = w6; l->words[7] = w7;
    line_calc_crc(l);
    l->words[8] = l->calc_crc;
    l->service_type = SRV_CTRL_BLOCK;
}

static int16_t line_sample(const stc_line *l, uint8_t idx) { return (int16_t)(uint16_t)(l->words[idx]<<2); }
static bool line_almost_silent(const stc_line *l)
{   /* stc007line.cpp:582-606 (non-M2) */
    uint8_t cnt = 0;
    for(uint8_t i=0;i<6;i++) { int16_t s = line_sample(l, i); if(!(s>=16)&&!(s<-16)) cnt++; }
    return cnt>=2;
}
static uint8_t line_words_diff(const stc_line *a, const stc_line *b)
{   /* stc007line.cpp:329-356: the XOR is truncated to 8 bits (reference quirk) */
    uint8_t cnt = 0;
    for(uint8_t i=0;i<8;i++)
    {
        uint8_t diff = (uint8_t)(a->words[i]^b->words[i]);
        if(diff!=0) for(uint8_t bit=0;bit<=16;bit++) if((diff&(1<<bit))!=0) cnt++;
    }
    return cnt;
}

/* ------------------------------------------------------------------------------------------- binarizer */
static void bin_set_mode(binarizer *b, uint8_t mode)
{   /* binarizer.cpp:120-152 */
    if(mode==SDVO_MODE_DRAFT) { b->bin_mode = mode; b->in_max_hysteresis_depth = HYST_DEPTH_SAFE; b->in_max_shift_stages = SHIFT_MIN; }
    else if(mode==SDVO_MODE_FAST) { b->bin_mode = mode; b->in_max_hysteresis_depth = 7; b->in_max_shift_stages = SHIFT_SAFE; }
    else if(mode==SDVO_MODE_INSANE) { b->bin_mode = mode; b->in_max_hysteresis_depth = HYST_DEPTH_MAX; b->in_max_shift_stages = SHIFT_MAX; }
    else { b->bin_mode = SDVO_MODE_NORMAL; b->in_max_hysteresis_depth = HYST_DEPTH_SAFE; b->in_max_shift_stages = SHIFT_SAFE; }
}
static void bin_set_bw(binarizer *b, uint8_t bl, uint8_t wh)
{
    if((bl<wh)&&(bl<MAX_BLACK_LVL)&&(wh>MIN_WHITE_LVL)&&(wh!=0)) { b->in_def_black = bl; b->in_def_white = wh; }
    else b->in_def_black = b->in_def_white = 0;
}
static void bin_set_coords(binarizer *b, coord_t c) { if(coord_valid(c)) b->in_def_coord = c; else b->in_def_coord = coord_none(); }
static void bin_set_coords2(binarizer *b, int16_t s, int16_t e)
{
    coord_t t = coord_none();
    if((s<e)&&(e!=0)&&(s!=NO_COORD_LEFT)&&(e!=NO_COORD_RIGHT)) coord_set(&t, s, e);
    bin_set_coords(b, t);
}
static void bin_reset_good(binarizer *b) { b->in_def_reference = 0; bin_set_coords2(b, 0, 0); bin_set_bw(b, 0, 0); }
static void bin_set_good(binarizer *b, const stc_line *l)
{
    if(line_crc_ok_ign(l)) { b->in_def_reference = l->ref_level; bin_set_coords(b, l->coords); bin_set_bw(b, l->black_level, l->white_level); }
}
static bool bin_ref_preset(const binarizer *b) { return b->in_def_reference>=MIN_REF_LVL; }
static bool bin_bw_preset(const binarizer *b)
{
    if((b->in_def_white>MIN_WHITE_LVL)&&(b->in_def_black<MAX_BLACK_LVL))
    {
        if(bin_ref_preset(b)) { if((b->in_def_reference<=b->in_def_black)||(b->in_def_reference>=b->in_def_white)) return false; }
        return true;
    }
    return false;
}
static void bin_init(binarizer *b)
{
    memset(b, 0, sizeof(*b));
    b->in_def_coord = coord_none();
    b->do_coord_search = true;
    b->do_ref_lvl_sweep = false;
    b->mark_end_min = 0xFFFF;
    bin_set_mode(b, SDVO_MODE_FAST);
}

static uint8_t get_low_level(uint8_t lvl, uint8_t diff) { if(lvl>diff) lvl = (uint8_t)(lvl-diff); else lvl = 1; return lvl; }
static uint8_t get_high_level(uint8_t lvl, uint8_t diff) { if(lvl<(255-diff)) lvl = (uint8_t)(lvl+diff); else lvl = 254; return lvl; }
static uint8_t pick_center_ref(uint8_t bl, uint8_t wh)
{   /* binarizer.cpp:3504-3548 */
    uint8_t d = (uint8_t)(wh-bl), r;
    if(d>=MIN_CONTRAST) { d = d/2; r = (uint8_t)(d+bl); if(r<MIN_REF_LVL) r = MIN_REF_LVL; else if(r>MAX_REF_LVL) r = MAX_REF_LVL; }
    else { if(wh<MAX_REF_LVL) r = MAX_REF_LVL; else r = MIN_REF_LVL; }
    return r;
}

/* ---- CRC statistics (binarizer.cpp:1771-1950) */
static void reset_crc_stats(crc_handler_t *a, uint16_t n, uint8_t *cnt)
{
    for(uint16_t i=0;i<n;i++) { a[i].result = 0; a[i].data_start = a[i].data_stop = 0; a[i].crc = 0; a[i].hyst_dph = a[i].shift_stg = 0x0f; }
    if(cnt) *cnt = 0;
}
static void update_crc_stats(crc_handler_t *a, crc_handler_t in, uint8_t *cnt)
{
    bool found = false;
    if(*cnt>=MAX_COLL_CRCS) *cnt = MAX_COLL_CRCS-1;
    for(uint8_t i=1;i<=*cnt;i++) if(a[i].crc==in.crc) { a[i].result++; found = true; break; }
    if(!found)
    {
        (*cnt)++;
        if(*cnt<MAX_COLL_CRCS) { a[*cnt].crc = in.crc; a[*cnt].hyst_dph = in.hyst_dph; a[*cnt].shift_stg = in.shift_stg; a[*cnt].result++; }
    }
}
static void find_most_frequent_crc(crc_handler_t *a, uint8_t *cnt, bool skip_equal)
{
    a[0].result = 0; a[0].data_start = 0; a[0].data_stop = 0; a[0].hyst_dph = 0; a[0].shift_stg = 0;
    if(*cnt>=MAX_COLL_CRCS) *cnt = MAX_COLL_CRCS-1;
    for(uint8_t i=1;i<=*cnt;i++)
        if(a[i].result>a[0].result) { a[0].result = a[i].result; a[0].crc = a[i].crc; a[0].hyst_dph = a[i].hyst_dph; a[0].shift_stg = a[i].shift_stg; a[0].data_start = i; }
    if(skip_equal)
        for(uint8_t i=1;i<=*cnt;i++)
            if(a[0].data_start!=i) if(a[0].result<=(2*a[i].result)) { a[0].result = 0; a[0].hyst_dph = 0; a[0].shift_stg = 0; break; }
    if(a[0].result==0) *cnt = 0;
}
static void invalidate_non_frequent(crc_handler_t *a, uint8_t lo, uint8_t hi, uint8_t cnt, uint16_t target)
{
    uint8_t idx = hi;
    while(idx>=lo)
    {
        if(a[idx].result==REF_CRC_OK) if((cnt==0)||(a[idx].crc!=target)) a[idx].result = REF_CRC_COLL;
        if(idx==lo) break;
        idx--;
    }
}

/* binarizer.cpp:1953-2131 */
static uint8_t pick_level_by_stats(crc_handler_t *c, uint8_t *res, uint8_t low_lvl, uint8_t high_lvl, uint8_t target, uint8_t max_hyst, uint8_t max_shift)
{
    bool good = false, range_lock = false, second_lock = false;
    uint8_t idx, low_depth = 0xFF, low_shift = 0xFF, low_ref = 0, high_ref = 0, tst_low = 0, tst_high = 0, picked;
    idx = high_lvl;
    while(idx>=low_lvl)
    {
        if((c[idx].result==target)&&(c[idx].hyst_dph<=max_hyst)&&(c[idx].shift_stg<=max_shift))
        {
            good = true;
            if(c[idx].hyst_dph<low_depth) { low_depth = c[idx].hyst_dph; low_shift = c[idx].shift_stg; high_ref = idx; }
            else if(c[idx].hyst_dph==low_depth) { if(c[idx].shift_stg<low_shift) { low_shift = c[idx].shift_stg; high_ref = idx; } }
        }
        if(idx==low_lvl) break;
        idx--;
    }
    if(!good) return SPAN_NOT_FOUND;
    idx = high_ref;
    while(idx>=low_lvl)
    {
        if((c[idx].result==target)&&(c[idx].hyst_dph==low_depth)&&(c[idx].shift_stg==low_shift))
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
    picked = (uint8_t)(high_ref-low_ref); picked = picked/2; picked = (uint8_t)(low_ref+picked);
    *res = picked;
    return SPAN_OK;
}

/* binarizer.cpp:2134-2300 */
static uint8_t pick_level_by_stats_opt(crc_handler_t *c, uint8_t *res, uint8_t low_lvl, uint8_t high_lvl, uint8_t target, uint8_t max_hyst, uint8_t max_shift)
{
    bool range_lock = false, good = false;
    uint8_t idx, hold_cnt, same_cnt, low_depth, low_shift = 0, high_shift = 0, low_ref = 0, high_ref = 0, picked;
    idx = high_lvl;
    while(idx>=low_lvl)
    {
        if((c[idx].result==target)&&(c[idx].hyst_dph<=max_hyst)&&(c[idx].shift_stg<=max_shift))
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
        if((c[idx].result==target)&&(c[idx].hyst_dph<=max_hyst)&&(c[idx].shift_stg<=max_shift))
        {
            good = true;
            if(low_depth>c[idx].hyst_dph) { low_depth = c[idx].hyst_dph; low_shift = c[idx].shift_stg; low_ref = high_ref = idx; hold_cnt = MIN_VALID_CRCS; }
            else if(low_depth==c[idx].hyst_dph)
            {
                if(low_shift>c[idx].shift_stg) { low_shift = c[idx].shift_stg; low_ref = high_ref = idx; same_cnt = MIN_VALID_CRCS; hold_cnt = MIN_VALID_CRCS; }
                else if(low_shift==c[idx].shift_stg) { low_ref = idx; same_cnt--; if(same_cnt==0) { hold_cnt = 0; break; } }
                else { hold_cnt--; if(hold_cnt==0) break; }
            }
            else { hold_cnt--; if(hold_cnt==0) break; }
        }
        if(idx==low_lvl) break;
        idx--;
    }
    if(good) { picked = (uint8_t)(high_ref-low_ref); picked = picked/2; picked = (uint8_t)(low_ref+picked); *res = picked; return SPAN_OK; }
    return SPAN_NOT_FOUND;
}

/* ---- AGC (binarizer.cpp:2385-2500, 2684-3070, 3116-3473) */
static uint16_t most_frequent_count(const uint16_t *s) { uint16_t m = 0; for(int i=255;i>=0;i--) if(s[i]>m) m = s[i]; return m; }
static uint8_t usefull_low(const uint16_t *s)
{
    uint8_t lev = 0, lowest = 0; bool found = false;
    uint16_t mf = most_frequent_count(s)/64;
    while(lev<MAX_BLACK_LVL) { if(s[lev]>mf) { lowest = lev; found = true; break; } lev++; }
    if(!found) while(lev<MAX_BLACK_LVL) { if(s[lev]>0) { lowest = lev; break; } lev++; }
    return lowest;
}
static uint8_t usefull_high(const uint16_t *s)
{
    uint8_t lev = 255, highest = 255;
    uint16_t mf = most_frequent_count(s)/64;
    while(lev>=MIN_WHITE_LVL) { if(s[lev]>mf) { highest = lev; break; } lev--; }
    /* [filtered_found] is never set in the reference, so the unfiltered pass always follows (continuing from [lev]). */
    while(lev>=MIN_WHITE_LVL) { if(s[lev]>0) { highest = lev; break; } lev--; }
    return highest;
}

static void find_stc007_bw(binarizer *b, stc_line *out, uint16_t *sprd)
{
    uint8_t pixel_val, brt_lev, stage;
    uint8_t br_mark_white, useful_low, useful_high, high_scan_limit, range_limit, bin_level, bin_low, bin_high;
    uint16_t pixel, pixel_limit, mark_ed_bit_start, mark_ed_bit_end, white_lvl_count, search_lim;
    uint32_t temp_calc;
    bool white_detected;
    search_lim = (uint16_t)(b->scan_start+b->estimated_ppb*10);
    for(pixel=b->scan_start;pixel<search_lim;pixel++) sprd[b->px[pixel]]++;
    search_lim = (uint16_t)(b->scan_end-b->estimated_ppb*20);
    for(pixel=search_lim;pixel<=b->scan_end;pixel++) sprd[b->px[pixel]]++;
    useful_low = usefull_low(sprd);
    useful_high = high_scan_limit = br_mark_white = usefull_high(sprd);
    range_limit = (uint8_t)(high_scan_limit-useful_low);
    high_scan_limit = (uint8_t)(high_scan_limit-(range_limit/4));
    bin_high = range_limit/8;
    brt_lev = useful_high;
    white_lvl_count = 0;
    white_detected = false;
    while(brt_lev>=high_scan_limit)
    {
        if(sprd[brt_lev]>white_lvl_count) { white_lvl_count = sprd[brt_lev]; br_mark_white = brt_lev; white_detected = true; }
        if(white_detected) if((br_mark_white-brt_lev)>=bin_high) break;
        brt_lev--;     /* uint8 wrap mirrors the reference */
    }
    pixel_limit = (uint16_t)(b->scan_end-b->scan_start);
    temp_calc = pixel_limit/8;
    pixel_limit = (uint16_t)(b->scan_start+(uint16_t)temp_calc);
    search_lim = (uint16_t)(b->scan_end-(uint16_t)temp_calc);
    memset(sprd, 0, 256*sizeof(uint16_t));
    for(pixel=pixel_limit;pixel<search_lim;pixel++) sprd[b->px[pixel]]++;
    stage = MARK_ED_START;
    mark_ed_bit_start = mark_ed_bit_end = 0;
    if(white_detected)
    {
        bin_level = pick_center_ref(useful_low, br_mark_white);
        bin_high = bin_low = bin_level;
        if(b->mark_end_min>(b->estimated_ppb*6)) pixel_limit = (uint16_t)(b->mark_end_min-b->estimated_ppb*6);
        else pixel_limit = 0;
        pixel = b->scan_end;
        while(pixel>pixel_limit)
        {
            pixel_val = b->px[pixel];
            if(stage==MARK_ED_START)
            {
                if(pixel<b->mark_end_min) break;
                if(pixel_val>=bin_low) { mark_ed_bit_end = (uint16_t)(pixel+1); stage = MARK_ED_TOP; }
            }
            else if(stage==MARK_ED_TOP)
            {
                if(pixel_val<bin_high)
                {
                    mark_ed_bit_start = (uint16_t)(pixel+1);
                    stage = MARK_ED_BOT;
                    if((mark_ed_bit_end-mark_ed_bit_start)>=(b->estimated_ppb*2)) { stage = MARK_ED_LEN_OK; break; }
                    else stage = MARK_ED_START;
                }
            }
            pixel--;
        }
        out->mark_ed_stage = stage;
        out->coords.stop = (int16_t)mark_ed_bit_start;
        out->marker_stop_ed = mark_ed_bit_end;
        if(line_has_stop(out))
        {
            uint16_t sprd_cnt = 0;
            search_lim = (uint16_t)(b->estimated_ppb*64);
            if(search_lim>mark_ed_bit_start) search_lim = b->mark_start_max;
            else search_lim = (uint16_t)(mark_ed_bit_start-search_lim);
            memset(sprd, 0, 256*sizeof(uint16_t));
            for(pixel=(uint16_t)(mark_ed_bit_start-1);pixel>search_lim;pixel--) { sprd[b->px[pixel]]++; sprd_cnt++; }
            if(sprd_cnt<32)
            {
                pixel_limit = (uint16_t)(b->scan_end-b->scan_start);
                pixel_limit = pixel_limit/8;
                search_lim = (uint16_t)(b->scan_end-pixel_limit);
                for(pixel=pixel_limit;pixel<search_lim;pixel++) sprd[b->px[pixel]]++;
            }
        }
    }
}

static bool find_black_white(binarizer *b, stc_line *out)
{
    uint8_t brt_lev, br_black = 0, br_white = 255, useful_low, useful_high, low_scan_limit, high_scan_limit, range_limit, bin_low, bin_high;
    uint16_t black_cnt, white_cnt, search_lim;
    uint32_t temp_calc;
    uint16_t sprd[256];
    bool black_det, white_det;
    memset(sprd, 0, sizeof(sprd));
    find_stc007_bw(b, out, sprd);
    useful_low = low_scan_limit = br_black = usefull_low(sprd);
    useful_high = high_scan_limit = br_white = usefull_high(sprd);
    range_limit = (uint8_t)(high_scan_limit-low_scan_limit);
    low_scan_limit = (uint8_t)(low_scan_limit+(range_limit/3));
    high_scan_limit = (uint8_t)(high_scan_limit-(range_limit/3));
    temp_calc = range_limit; temp_calc = temp_calc*10/100; bin_low = (uint8_t)temp_calc;
    temp_calc = range_limit; temp_calc = temp_calc*12/100; bin_high = (uint8_t)temp_calc;
    search_lim = most_frequent_count(sprd)/64;
    brt_lev = useful_low; black_cnt = 0; black_det = false;
    while(brt_lev<=low_scan_limit)
    {
        if(sprd[brt_lev]>black_cnt) { black_cnt = sprd[brt_lev]; if(black_cnt>search_lim) { br_black = brt_lev; black_det = true; } }
        if(black_det) if((brt_lev-br_black)>=bin_low) break;
        brt_lev++;      /* uint8 wrap mirrors the reference */
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
        else if(b->do_ref_lvl_sweep&&((br_white-br_black)<MIN_VALID_CRCS)) inv = true;
        else if(br_black>MAX_BLACK_LVL) inv = true;
        else if(br_white<MIN_WHITE_LVL) inv = true;
        if(inv) { black_det = white_det = false; br_black = useful_low; br_white = useful_high; }
    }
    b->was_BW_scanned = true;
    out->black_level = br_black;
    out->white_level = br_white;
    out->blk_wht_set = black_det&&white_det;
    return out->blk_wht_set;
}

/* ---- markers (binarizer.cpp:5275-5595, 6047-6113) */
static void search_markers(const binarizer *b, stc_line *l, uint8_t hyst_lvl)
{
    uint8_t stage = MARK_ST_START, pv, bin_level = l->ref_level, bin_low, bin_high;
    uint16_t pixel, pixel_limit, st1s = 0, st1e = 0, st3s = 0, st3e = 0, ed_s = 0, ed_e = 0;
    bin_low = get_low_level(bin_level, hyst_lvl);
    if(bin_low<MIN_REF_LVL) bin_low = MIN_REF_LVL;
    bin_high = bin_level;
    pixel_limit = (uint16_t)(b->mark_start_max+b->estimated_ppb*5);
    if(pixel_limit>b->line_length) pixel_limit = b->line_length;
    pixel = b->scan_start;
    while(pixel<pixel_limit)
    {
        pv = b->px[pixel];
        if(stage==MARK_ST_START)
        {
            if(pixel>b->mark_start_max) break;
            if(pv>=bin_low) { st1s = pixel; stage = MARK_ST_TOP_1; }
        }
        else if(stage==MARK_ST_TOP_1)
        {
            if(pv<bin_low) { st1e = pixel; stage = MARK_ST_BOT_1; }
        }
        else if(stage==MARK_ST_BOT_1)
        {
            if(pv>=bin_high)
            {
                st3s = pixel;
                if(((st3s-st1e)>(b->estimated_ppb*2))||((st3s-st1e)<(b->estimated_ppb/2))) stage = MARK_ST_START;
                else stage = MARK_ST_TOP_2;
            }
        }
        else if(stage==MARK_ST_TOP_2)
        {
            if(pv<bin_high)
            {
                st3e = pixel;
                if(((st3e-st3s)>(b->estimated_ppb*2))||((st3e-st3s)<(b->estimated_ppb/2))) stage = MARK_ST_START;
                else { stage = MARK_ST_BOT_2; break; }
            }
        }
        pixel++;
    }
    l->mark_st_stage = stage;
    l->marker_start_bg = st1s;
    l->marker_start_ed = st3e;
    stage = MARK_ED_START;
    if(line_has_start(l))
    {
        bin_low = bin_level;
        if(b->mark_end_min>(b->estimated_ppb*6)) pixel_limit = (uint16_t)(b->mark_end_min-b->estimated_ppb*6);
        else pixel_limit = 0;
        pixel = b->scan_end;
        while(pixel>pixel_limit)
        {
            pv = b->px[pixel];
            if(stage==MARK_ED_START)
            {
                if(pixel<b->mark_end_min) break;
                if(pv>=bin_low) { ed_e = (uint16_t)(pixel+1); stage = MARK_ED_TOP; }
            }
            else if(stage==MARK_ED_TOP)
            {
                if(pv<bin_high)
                {
                    ed_s = (uint16_t)(pixel+1);
                    stage = MARK_ED_BOT;
                    if(((ed_e-ed_s)>=(b->estimated_ppb*2))&&((ed_e-ed_s)<=(b->estimated_ppb*5))) { stage = MARK_ED_LEN_OK; break; }
                    else stage = MARK_ED_START;
                }
            }
            pixel--;
        }
        l->mark_ed_stage = stage;
    }
    coord_set(&l->coords, (int16_t)st1e, (int16_t)ed_s);
    l->marker_stop_ed = ed_e;
    l->coords_set = line_has_markers(l);
}

static void find_coordinates(const binarizer *b, stc_line *l)
{
    uint8_t best = 0;
    stc_line tmp = *l;
    coord_t bestc = coord_none(); bool have = false;
    for(uint8_t h=0;h<24;h++)
    {
        search_markers(b, &tmp, h);
        if(line_has_markers(&tmp))
        {   /* std::sort + first element == minimum under CoordinatePair::operator< (strict order, reference = h) */
            coord_t c = tmp.coords; c.reference = h;
            if(!have||coord_less(c, bestc)) { bestc = c; have = true; }
        }
    }
    if(have) best = bestc.reference;
    search_markers(b, l, best);
}

/* ---- bit extraction (binarizer.cpp:7322-7445, 7560-7691) */
static void fill_stc007(const binarizer *b, stc_line *l, uint8_t shift_stg)
{
    bool prev_high = false;
    uint8_t low_ref = l->ref_low, high_ref = l->ref_high, word_bit_pos = 13, word_index = 0, pcm_bit = 0, pv;
    uint16_t pcm_word = 0;
    while(pcm_bit<=(BITS_PCM_DATA-1))
    {
        pv = b->px[l->pix[shift_stg][pcm_bit]];
        if(!prev_high) { if(pv>low_ref) { pcm_word |= (uint16_t)(1<<word_bit_pos); prev_high = true; } }
        else { if(pv>=high_ref) pcm_word |= (uint16_t)(1<<word_bit_pos); else prev_high = false; }
        if(word_bit_pos==0)
        {
            if(pcm_bit>(BITS_PCM_DATA-16-1)) l->words[8] = pcm_word;
            else { l->words[word_index] = pcm_word&0x3FFF; l->word_crc[word_index] = l->word_valid[word_index] = false; }
            pcm_word = 0;
            word_index++;
            if(pcm_bit>(BITS_PCM_DATA-16-1)) break;
            else if(pcm_bit==(BITS_PCM_DATA-16-1)) word_bit_pos = 16;
            else word_bit_pos = 14;
        }
        word_bit_pos--;
        pcm_bit++;
    }
    line_calc_crc(l);
}

static uint8_t fill_data_words(const binarizer *b, stc_line *l, uint8_t ref_delta, uint8_t shift_stg)
{
    uint8_t low_ref, high_ref;
    if(ref_delta>HYST_DEPTH_MAX) return STG_NO_GOOD;
    if(shift_stg>SHIFT_MAX) return STG_NO_GOOD;
    low_ref = get_low_level(l->ref_level, ref_delta);
    high_ref = get_high_level(l->ref_level, ref_delta);
    l->ref_low = low_ref; l->ref_high = high_ref;
    if(low_ref<=l->black_level) { line_set_invalid_crc(l); return STG_NO_GOOD; }
    if(high_ref>=l->white_level) { line_set_invalid_crc(l); return STG_NO_GOOD; }
    l->hysteresis_depth = ref_delta;
    l->shift_stage = shift_stg;
    fill_stc007(b, l, shift_stg);
    return STG_DATA_OK;
}

/* binarizer.cpp:7695-8055 */
static void read_pcm_data(binarizer *b, stc_line *l)
{
    bool invalid_hyst;
    uint8_t hyst_cnt, shift_try, valid_hyst_n, valid_shift_n, hyst_good_cnt, valid_delta, valid_shift;
    line_calc_ppb(l, l->coords);
    if(b->hysteresis_depth_lim>HYST_DEPTH_MAX) b->hysteresis_depth_lim = HYST_DEPTH_MAX;
    if(b->shift_stages_lim>SHIFT_MAX) b->shift_stages_lim = SHIFT_MAX;
    if(!l->ref_level_sweeped)
    {
        hyst_cnt = (uint8_t)(b->hysteresis_depth_lim+1);
        while(hyst_cnt>0) { hyst_cnt--; b->hyst_crcs[hyst_cnt].result = REF_BAD_CRC; }
        valid_delta = hyst_good_cnt = 0;
        hyst_cnt = 0;
        do
        {
            invalid_hyst = false;
            reset_crc_stats(b->crc_stats, MAX_COLL_CRCS+1, &valid_shift_n);
            b->crc_stats[0].hyst_dph = 0; b->crc_stats[0].shift_stg = 0;
            shift_try = (uint8_t)(b->shift_stages_lim+1);
            while(shift_try>0) { shift_try--; b->shift_crcs[shift_try].result = REF_BAD_CRC; }
            shift_try = 0;
            do
            {
                b->shift_crcs[shift_try].hyst_dph = hyst_cnt;
                b->shift_crcs[shift_try].shift_stg = shift_try;
                if(fill_data_words(b, l, hyst_cnt, shift_try)!=STG_DATA_OK) { invalid_hyst = true; break; }
                else
                {
                    b->shift_crcs[shift_try].crc = l->calc_crc;
                    if(line_crc_ok(l))
                    {
                        b->shift_crcs[shift_try].result = REF_CRC_OK;
                        update_crc_stats(b->crc_stats, b->shift_crcs[shift_try], &valid_shift_n);
                        break;
                    }
                }
                shift_try++;
            }
            while(shift_try<=b->shift_stages_lim);
            if(valid_shift_n>0)
            {
                find_most_frequent_crc(b->crc_stats, &valid_shift_n, true);
                invalidate_non_frequent(b->shift_crcs, 0, b->shift_stages_lim, valid_shift_n, b->crc_stats[0].crc);
            }
            b->hyst_crcs[hyst_cnt].shift_stg = b->crc_stats[0].shift_stg;
            b->hyst_crcs[hyst_cnt].crc = b->crc_stats[0].crc;
            if(valid_shift_n>0)
            {
                b->hyst_crcs[hyst_cnt].hyst_dph = b->crc_stats[0].hyst_dph;
                b->hyst_crcs[hyst_cnt].result = REF_CRC_OK;
                hyst_good_cnt++;
                break;
            }
            else
            {
                b->hyst_crcs[hyst_cnt].hyst_dph = hyst_cnt;
                if(hyst_good_cnt>0) break;
            }
            if(invalid_hyst) break;
            hyst_cnt++;
        }
        while(hyst_cnt<=b->hysteresis_depth_lim);
        reset_crc_stats(b->crc_stats, MAX_COLL_CRCS, &valid_hyst_n);
        b->crc_stats[0].hyst_dph = 0; b->crc_stats[0].shift_stg = 0;
        if(hyst_good_cnt>0)
        {
            for(uint8_t i=0;i<=hyst_cnt;i++) if(b->hyst_crcs[i].result==REF_CRC_OK) update_crc_stats(b->crc_stats, b->hyst_crcs[i], &valid_hyst_n);
            if(valid_hyst_n>0)
            {
                find_most_frequent_crc(b->crc_stats, &valid_hyst_n, true);
                /* the reference scans hyst_crcs[0 .. hyst_cnt+1]; hyst_cnt <= HYST_DEPTH_MAX-... stays inside the array here */
                invalidate_non_frequent(b->hyst_crcs, 0, (uint8_t)(hyst_cnt+1 <= HYST_DEPTH_MAX ? hyst_cnt+1 : HYST_DEPTH_MAX), valid_hyst_n, b->crc_stats[0].crc);
            }
        }
        valid_delta = b->crc_stats[0].hyst_dph;
        valid_shift = b->crc_stats[0].shift_stg;
    }
    else
    {
        valid_delta = b->hysteresis_depth_lim;
        valid_shift = b->shift_stages_lim;
    }
    fill_data_words(b, l, valid_delta, valid_shift);
}

/* ---- reference level sweep (binarizer.cpp:3551-3817, 3821-4120) */
static void sweep_ref_level(binarizer *b, stc_line *pcm_line, crc_handler_t *res)
{
    bool skip_bin;
    uint8_t low_lvl = pcm_line->black_level, high_lvl = pcm_line->white_level;
    uint16_t ref_index;
    stc_line d;
    low_lvl = (uint8_t)(low_lvl+1);
    high_lvl = (uint8_t)(high_lvl-1);
    if(MIN_REF_LVL>low_lvl) low_lvl = MIN_REF_LVL;
    if(MAX_REF_LVL<high_lvl) high_lvl = MAX_REF_LVL;
    ref_index = high_lvl;
    line_clear(&d);     /* STC007Line temp_stc constructor */
    while(ref_index>=low_lvl)
    {
        /* dummy_line->clear() through the base pointer (binarizer.cpp:3615): the source CRC word of the previous
           level survives while calc_crc becomes 0, so a previous read of CRCC==0x0000 makes every lower level look
           "already valid" and skips its search (reference quirk, mirrored). */
        line_base_clear(&d);
        line_set_source_pixels(&d, 0, (uint16_t)(b->line_length-1));
        d.black_level = low_lvl; d.white_level = high_lvl;
        d.ref_level = (uint8_t)ref_index;
        if(coord_valid(b->in_def_coord))
        {
            skip_bin = false;
            find_coordinates(b, &d);
            if(!line_has_markers(&d)) skip_bin = true;
            if(skip_bin) { d.coords = b->in_def_coord; read_pcm_data(b, &d); }
        }
        if(!line_crc_ok(&d))
        {
            find_coordinates(b, &d);
            if(d.coords_set) read_pcm_data(b, &d);
        }
        if(d.hysteresis_depth>0x0F) d.hysteresis_depth = 0x0F;
        if(line_crc_ok(&d)&&coord_valid(d.coords))
        {
            res[ref_index].result = REF_CRC_OK;
            res[ref_index].data_start = d.coords.start; res[ref_index].data_stop = d.coords.stop;
            res[ref_index].hyst_dph = d.hysteresis_depth; res[ref_index].shift_stg = d.shift_stage; res[ref_index].crc = d.calc_crc;
        }
        else if(d.coords_set)
        {
            res[ref_index].result = REF_BAD_CRC;
            res[ref_index].data_start = d.coords.start; res[ref_index].data_stop = d.coords.stop;
            res[ref_index].hyst_dph = d.hysteresis_depth; res[ref_index].shift_stg = d.shift_stage; res[ref_index].crc = d.calc_crc;
        }
        ref_index--;
    }
}

static void calc_ref_by_sweep(binarizer *b, stc_line *l)
{
    uint8_t fast_ref, bin_level, valid_cnt, span_res;
    crc_handler_t sw[256];
    fast_ref = pick_center_ref(l->black_level, l->white_level);
    b->hysteresis_depth_lim = b->in_max_hysteresis_depth;
    b->shift_stages_lim = b->in_max_shift_stages;
    reset_crc_stats(sw, 256, NULL);
    sweep_ref_level(b, l, sw);
    span_res = SPAN_NOT_FOUND;
    reset_crc_stats(b->crc_stats, MAX_COLL_CRCS+1, &valid_cnt);
    b->crc_stats[0].hyst_dph = 0; b->crc_stats[0].shift_stg = 0;
    for(bin_level=(uint8_t)(l->white_level-1);bin_level>l->black_level;bin_level--)
        if(sw[bin_level].result==REF_CRC_OK) update_crc_stats(b->crc_stats, sw[bin_level], &valid_cnt);
    if(valid_cnt>0)
    {
        find_most_frequent_crc(b->crc_stats, &valid_cnt, true);
        invalidate_non_frequent(sw, (uint8_t)(l->black_level+1), (uint8_t)(l->white_level-1), valid_cnt, b->crc_stats[0].crc);
        if(valid_cnt>0)
        {
            if(b->crc_stats[0].result<MIN_VALID_CRCS) span_res = SPAN_TOO_NARROW;
            else span_res = pick_level_by_stats(sw, &l->ref_level, (uint8_t)(l->black_level+1), (uint8_t)(l->white_level-1), REF_CRC_OK, 0x0F, SHIFT_MAX);
        }
    }
    if(span_res==SPAN_OK)
    {
        crc_handler_t t = sw[l->ref_level];
        l->ref_level_sweeped = true;
        coord_set(&l->coords, t.data_start, t.data_stop);
        l->coords_set = true;
        find_coordinates(b, l);
        b->hysteresis_depth_lim = t.hyst_dph;
        if(b->hysteresis_depth_lim>HYST_DEPTH_MAX) b->hysteresis_depth_lim = HYST_DEPTH_MAX;
        b->shift_stages_lim = t.shift_stg;
    }
    else
    {
        if(span_res==SPAN_TOO_NARROW)
        {
            span_res = pick_level_by_stats_opt(sw, &l->ref_level, (uint8_t)(l->black_level+1), (uint8_t)(l->white_level-1), REF_CRC_OK, b->hysteresis_depth_lim, b->shift_stages_lim);
            l->forced_bad = true;
        }
        else
            span_res = pick_level_by_stats(sw, &l->ref_level, (uint8_t)(l->black_level+1), (uint8_t)(l->white_level-1), REF_BAD_CRC, 0xFF, 0xFF);
        if(span_res==SPAN_OK)
        {
            crc_handler_t t = sw[l->ref_level];
            coord_set(&l->coords, t.data_start, t.data_stop);
            l->coords_set = true;
            find_coordinates(b, l);
        }
        else if(bin_ref_preset(b))
        {
            l->ref_level = b->in_def_reference;
            if(coord_valid(b->in_def_coord)) l->coords = b->in_def_coord;
        }
        else
        {
            l->ref_level = fast_ref;
            if(!coord_valid(b->in_def_coord)) coord_set(&l->coords, (int16_t)(b->scan_start+b->estimated_ppb), (int16_t)(b->scan_end-(4*b->estimated_ppb)));
            else l->coords = b->in_def_coord;
        }
        b->hysteresis_depth_lim = HYST_DEPTH_MIN;
        b->shift_stages_lim = SHIFT_MIN;
    }
}

/* ---- Binarizer::processLine for a non-service, non-empty STC-007 line (binarizer.cpp:443-1724) */
static int process_line(binarizer *b, const uint8_t *px, uint16_t W, uint32_t frame, uint16_t line_no, stc_line *o)
{
    uint8_t stage_count, proc_state;
    uint32_t tmp;
    line_clear(o);
    o->frame_number = frame; o->line_number = line_no;
    b->px = px;
    b->line_length = W;
    b->scan_start = 0;
    b->scan_end = (uint16_t)(W-1);
    line_set_source_pixels(o, b->scan_start, b->scan_end);
    if(W<BITS_IN_LINE) return 3;    /* LB_RET_SHORT_LINE */
    b->mark_start_max = (uint16_t)(W*MARK_MAX_DIST);
    b->mark_start_max = b->mark_start_max/100;
    b->mark_end_min = (uint16_t)(b->scan_end-b->mark_start_max);
    b->mark_start_max = (uint16_t)(b->scan_start+b->mark_start_max);
    tmp = (uint32_t)W*INT_CALC_MULT;
    tmp = tmp/BITS_IN_LINE;
    b->estimated_ppb = (uint16_t)((tmp+(INT_CALC_MULT/2))/INT_CALC_MULT);
    coord_set(&o->coords, (int16_t)b->scan_start, (int16_t)b->scan_end);
    proc_state = STG_REF_FIND;
    b->was_BW_scanned = false;
    if(bin_bw_preset(b)) { o->black_level = b->in_def_black; o->white_level = b->in_def_white; o->blk_wht_set = true; }
    if(bin_ref_preset(b)) { if(coord_valid(b->in_def_coord)) proc_state = STG_INPUT_ALL; else proc_state = STG_INPUT_LEVEL; }
    b->hysteresis_depth_lim = b->in_max_hysteresis_depth;
    b->shift_stages_lim = b->in_max_shift_stages;
    stage_count = 0;
    do
    {
        stage_count++;
        if(proc_state==STG_INPUT_ALL)
        {
            if(!o->blk_wht_set) find_black_white(b, o);
            o->coords = b->in_def_coord;
            o->ref_level = b->in_def_reference;
            o->ref_level_sweeped = false;
            if(!o->blk_wht_set) proc_state = STG_NO_GOOD;
            else if((b->in_def_reference>=o->white_level)||(b->in_def_reference<=o->black_level)) proc_state = STG_REF_FIND;
            else
            {
                read_pcm_data(b, o);
                if(line_crc_ok(o)) { o->data_by_ext_tune = true; proc_state = STG_DATA_OK; }
                else proc_state = STG_INPUT_LEVEL;
            }
        }
        else if(proc_state==STG_INPUT_LEVEL)
        {
            if(!b->was_BW_scanned) find_black_white(b, o);
            coord_set(&o->coords, (int16_t)b->scan_start, (int16_t)b->scan_end);
            o->ref_level = b->in_def_reference;
            o->ref_level_sweeped = false;
            if(!o->blk_wht_set) proc_state = STG_NO_GOOD;
            else
            {
                proc_state = STG_REF_FIND;
                if((b->in_def_reference<o->white_level)&&(b->in_def_reference>o->black_level))
                {
                    if(!b->do_coord_search)
                    {
                        if(!coord_valid(b->in_def_coord)) coord_set(&o->coords, (int16_t)(b->scan_start+b->estimated_ppb), (int16_t)(b->scan_end-(4*b->estimated_ppb)));
                        else o->coords = b->in_def_coord;
                    }
                    else find_coordinates(b, o);
                    if(line_has_markers(o))
                    {
                        if((!coord_valid(b->in_def_coord))||coord_ne(o->coords, b->in_def_coord))
                        {
                            read_pcm_data(b, o);
                            if(line_crc_ok(o)) { o->data_by_ext_tune = true; proc_state = STG_DATA_OK; }
                        }
                    }
                }
            }
        }
        else if(proc_state==STG_REF_FIND)
        {
            if(!b->was_BW_scanned) find_black_white(b, o);
            if(!o->blk_wht_set) proc_state = STG_NO_GOOD;
            else
            {
                b->do_ref_lvl_sweep = false;
                if((b->bin_mode==SDVO_MODE_NORMAL)||(b->bin_mode==SDVO_MODE_INSANE)) b->do_ref_lvl_sweep = true;
                if(b->do_ref_lvl_sweep) proc_state = STG_REF_SWEEP_RUN;
                else
                {
                    b->hysteresis_depth_lim = HYST_DEPTH_SAFE;
                    b->shift_stages_lim = SHIFT_MIN;
                    proc_state = STG_READ_PCM;
                    o->ref_level = pick_center_ref(o->black_level, o->white_level);
                    if(!b->do_coord_search)
                    {
                        if(!coord_valid(b->in_def_coord)) coord_set(&o->coords, (int16_t)(b->scan_start+b->estimated_ppb), (int16_t)(b->scan_end-(4*b->estimated_ppb)));
                        else o->coords = b->in_def_coord;
                    }
                    else find_coordinates(b, o);
                    if(!line_has_markers(o)) { b->hysteresis_depth_lim = HYST_DEPTH_SAFE; b->shift_stages_lim = SHIFT_MIN; }
                    else { b->hysteresis_depth_lim = b->in_max_hysteresis_depth; b->shift_stages_lim = b->in_max_shift_stages; }
                }
            }
        }
        else if(proc_state==STG_REF_SWEEP_RUN)
        {
            calc_ref_by_sweep(b, o);
            proc_state = STG_READ_PCM;
        }
        else if(proc_state==STG_READ_PCM)
        {
            if(o->coords_set) read_pcm_data(b, o);
            if(line_crc_ok(o)) proc_state = STG_DATA_OK;
            if(proc_state!=STG_DATA_OK)
            {
                if(coord_valid(b->in_def_coord)&&(!b->do_ref_lvl_sweep)&&(!o->forced_bad)&&(!o->coords_set))
                {
                    if(coord_ne(o->coords, b->in_def_coord))
                    {
                        o->coords = b->in_def_coord;
                        o->marker_start_bg = 0; o->marker_start_ed = 0; o->marker_stop_ed = 0;
                        read_pcm_data(b, o);
                        if(line_crc_ok(o)) proc_state = STG_DATA_OK;
                    }
                }
                if(proc_state!=STG_DATA_OK) proc_state = STG_NO_GOOD;
            }
        }
        else if(proc_state==STG_DATA_OK)
        {
            if(o->forced_bad) proc_state = STG_NO_GOOD;
            else
            {
                line_apply_crc_per_word(o);
                if(line_has_control_block(o)) line_set_serv_ctrl_blk(o);
                break;
            }
        }
        else if(proc_state==STG_NO_GOOD)
        {
            if(line_crc_ok(o)) line_set_invalid_crc(o);
            line_apply_crc_per_word(o);
            break;
        }
        else break;
        if(stage_count>STG_MAX) break;
    }
    while(1);
    return 0;
}

/* ------------------------------------------------------------------------------------------- record export */
static void export_rec(const stc_line *l, sdvo_line_rec *r)
{
    memset(r, 0, sizeof(*r));
    r->frame = l->frame_number; r->line = l->line_number;
    memcpy(r->words, l->words, sizeof(r->words));
    r->data_start = l->coords.start; r->data_stop = l->coords.stop;
    r->black = l->black_level; r->white = l->white_level; r->ref_low = l->ref_low; r->ref = l->ref_level; r->ref_high = l->ref_high;
    r->hyst = l->hysteresis_depth; r->shift = l->shift_stage;
    r->service_type = l->service_type;
    uint16_t f = 0;
    if(line_crc_ok(l)) f |= 1<<0;
    if(line_crc_ok_ign(l)) f |= 1<<1;
    if(l->forced_bad) f |= 1<<2;
    if(l->blk_wht_set) f |= 1<<3;
    if(l->coords_set) f |= 1<<4;
    if(l->ref_level_sweeped) f |= 1<<5;
    if(l->data_by_ext_tune) f |= 1<<6;
    if(line_has_markers(l)) f |= 1<<8;
    if(line_has_start(l)) f |= 1<<9;
    if(line_has_stop(l)) f |= 1<<10;
    if(line_almost_silent(l)) f |= 1<<12;
    r->flags = f;
    r->mark_st_stage = l->mark_st_stage; r->mark_ed_stage = l->mark_ed_stage;
    r->marker_start_bg = l->marker_start_bg; r->marker_start_ed = l->marker_start_ed; r->marker_stop_ed = l->marker_stop_ed;
    for(int i=0;i<9;i++)
    {
        if((!l->forced_bad)&&l->word_crc[i]) r->word_crc_mask |= (uint16_t)(1<<i);
        if((!l->forced_bad)&&l->word_valid[i]) r->word_valid_mask |= (uint16_t)(1<<i);
    }
    r->pcm_type = SDVO_TYPE_STC007;
}

int sdvo_binarize_lines_stc007(int mode, const uint8_t *luma, int n, int W, int stride,
                               int preset_ref, int preset_black, int preset_white, int preset_start, int preset_stop,
                               sdvo_line_rec *out)
{
    binarizer b; stc_line l;
    bin_init(&b);
    for(int i=0;i<n;i++)
    {
        bin_reset_good(&b);
        bin_set_mode(&b, (uint8_t)mode);
        b.do_coord_search = true;
        if(preset_ref>0) b.in_def_reference = (uint8_t)preset_ref;
        if(preset_white>0) bin_set_bw(&b, (uint8_t)preset_black, (uint8_t)preset_white);
        if(preset_stop!=0) bin_set_coords2(&b, (int16_t)preset_start, (int16_t)preset_stop);
        process_line(&b, luma+(size_t)i*stride, (uint16_t)W, 1, (uint16_t)(i+1), &l);
        export_rec(&l, &out[i]);
    }
    return n;
}

/* ------------------------------------------------------------------------------------------- V2D chain */
enum { FIELD_INIT = 0, FIELD_NEW, FIELD_SAFE, FIELD_UNSAFE };
enum { COORD_HISTORY_DEPTH = 9, COORD_LONG_HISTORY = 16 };

typedef struct { coord_t *v; int n, cap; } coord_list;
static void cl_push(coord_list *l, coord_t c)
{
    if(l->n==l->cap) { l->cap = l->cap ? l->cap*2 : 64; l->v = (coord_t *)realloc(l->v, (size_t)l->cap*sizeof(coord_t)); }
    l->v[l->n++] = c;
}
static void cl_pop_front(coord_list *l) { memmove(l->v, l->v+1, (size_t)(l->n-1)*sizeof(coord_t)); l->n--; }
static int coord_cmp(const void *a, const void *b)
{
    coord_t x = *(const coord_t *)a, y = *(const coord_t *)b;
    if(coord_less(x, y)) return -1;
    if(coord_less(y, x)) return 1;
    return 0;
}
/* VideoToDigital::medianCoordinates (videotodigital.cpp:348-371): element size/2 of the sorted list. */
static coord_t cl_median(const coord_list *l)
{
    if(l->n==0) return coord_none();
    coord_t *t = (coord_t *)malloc((size_t)l->n*sizeof(coord_t));
    memcpy(t, l->v, (size_t)l->n*sizeof(coord_t));
    qsort(t, (size_t)l->n, sizeof(coord_t), coord_cmp);
    coord_t r = t[l->n/2];
    free(t);
    return r;
}
static bool delta_warning(coord_t d, uint8_t lim)
{
    return (d.start<=-(int)lim)||(d.start>=(int)lim)||(d.stop<=-(int)lim)||(d.stop>=(int)lim);
}

int sdvo_v2d_stc007(int mode, int line_dup, const uint8_t *luma, int n_frames, int H, int W, sdvo_line_rec *out)
{
    binarizer b; stc_line line, last_line;
    coord_list last_valid = {0}, frame_valid = {0}, frame_invalid = {0}, long_valid = {0};
    coord_t frame_avg = coord_none(), target = coord_none(), delta;
    uint8_t field_state;
    int n_out = 0;
    bin_init(&b);
    line_clear(&last_line);
    /* NEW_FILE service line: stats reset, good parameters reset (videotodigital.cpp:1011-1026).
       It travels inside the first frame, after the frame-start block below has run once with empty history. */
    for(int f=0;f<n_frames;f++)
    {
        field_state = FIELD_NEW;
        /* Frame start (videotodigital.cpp:774-823); prescan is disabled for STC-007 (videotodigital.cpp:176-199). */
        frame_avg = coord_none();
        frame_avg = cl_median(&long_valid);
        if(coord_valid(frame_avg)) bin_set_coords2(&b, frame_avg.start, frame_avg.stop);
        if(f==0)
        {
            last_valid.n = frame_valid.n = frame_invalid.n = long_valid.n = 0;
            target = coord_none();
            if(!coord_valid(frame_avg)) bin_reset_good(&b);
        }
        for(int field=0;field<2;field++)
        {
            for(int k=0;k<H/2;k++)
            {
                int row = 2*k+field;
                bool count_has_pcm, count_has_data;
                bin_set_mode(&b, (uint8_t)mode);
                b.do_coord_search = true;
                process_line(&b, luma+((size_t)f*H+row)*W, (uint16_t)W, (uint32_t)(f+1), (uint16_t)(row+1), &line);
                if(line.service_type!=SRV_NO)
                {
                    if(line.service_type==SRV_CTRL_BLOCK) { if(field_state==FIELD_NEW) field_state = FIELD_SAFE; }
                }
                else
                {
                    count_has_data = line_has_markers(&line);
                    count_has_pcm = line_crc_ok(&line)||count_has_data;
                    if(count_has_pcm) { if(field_state==FIELD_NEW) field_state = FIELD_UNSAFE; }
                    if(line_crc_ok(&line))
                    {
                        if(line_dup)
                        {
                            if(field_state==FIELD_UNSAFE)
                            {
                                bin_set_good(&b, &line);
                                line.forced_bad = true;     /* en_first_line_dup */
                            }
                            else
                            {
                                uint8_t diff = line_words_diff(&line, &last_line);
                                bool same = diff<=(BITS_PCM_DATA/32);
                                if((!line_almost_silent(&line))&&same) line.forced_bad = true;
                            }
                        }
                        if(line_crc_ok_ign(&line))
                        {
                            cl_push(&last_valid, line.coords);
                            cl_push(&frame_valid, line.coords);
                            while(last_valid.n>COORD_HISTORY_DEPTH) cl_pop_front(&last_valid);
                            if(last_valid.n>(COORD_HISTORY_DEPTH/2))
                            {
                                target = cl_median(&last_valid);
                                if(!coord_valid(target)) target = frame_avg;
                                if(coord_valid(target))
                                {
                                    delta = line.coords;
                                    delta.start = (int16_t)(delta.start-target.start); delta.stop = (int16_t)(delta.stop-target.stop);
                                    if(delta_warning(delta, (uint8_t)(line_get_ppb(&line)*3))) line.forced_bad = true;
                                }
                            }
                        }
                        if(line_crc_ok(&line)) bin_set_good(&b, &line);
                        field_state = FIELD_INIT;
                    }
                    else
                    {
                        if(coord_valid(line.coords)) cl_push(&frame_invalid, line.coords);
                        if(count_has_data)
                        {
                            coord_t preset = cl_median(&last_valid);
                            if(!coord_valid(preset)) preset = frame_avg;
                            field_state = FIELD_INIT;
                            bin_set_coords(&b, preset);
                            bin_set_bw(&b, 0, 0);
                        }
                        else bin_set_bw(&b, 0, 0);
                    }
                    if(count_has_pcm) last_line = line;
                }
                export_rec(&line, &out[n_out++]);
            }
            /* END_FIELD service line (videotodigital.cpp:1028-1048) */
            field_state = FIELD_NEW;
            line_clear(&last_line);
        }
        /* END_FRAME service line (videotodigital.cpp:1659-1723) */
        frame_avg = cl_median(&frame_valid);
        if(coord_valid(frame_avg))
        {
            cl_push(&long_valid, frame_avg);
            while(long_valid.n>COORD_LONG_HISTORY) cl_pop_front(&long_valid);
        }
        else
        {
            frame_avg = cl_median(&frame_invalid);
            if(!coord_valid(frame_avg)) frame_avg = cl_median(&long_valid);
        }
        frame_valid.n = 0; frame_invalid.n = 0;
    }
    free(last_valid.v); free(frame_valid.v); free(frame_invalid.v); free(long_valid.v);
    return n_out;
}
