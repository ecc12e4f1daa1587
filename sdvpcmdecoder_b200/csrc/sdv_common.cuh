// sdv_common.cuh -- shared definitions of the B200 STC-007 decode path.
//
// Everything marked SDV_HD is plain integer logic that compiles both as CUDA device code (the product) and as
// host code (tests/hostemu only: the same source run with a "CTA" of one thread, see struct Cta).  The product
// library never calls the host instantiation; there is no CPU fallback behind the C ABI.
#pragma once
#include <stdint.h>
#include <string.h>
#include "../../include/sdvpcm.h"

#if defined(__CUDACC__)
#define SDV_HD __host__ __device__ __forceinline__
#define SDV_HDN __host__ __device__ __noinline__
#else
#define SDV_HD inline
#define SDV_HDN inline
#endif

namespace sdv {

typedef uint8_t u8; typedef uint16_t u16; typedef uint32_t u32; typedef int16_t i16; typedef int32_t i32; typedef uint64_t u64;

// ------------------------------------------------------------------------------------------------ constants
// STC-007 line structure (stc007line.h:72-152), PCMLine fixed point (pcmline.h:46-48,60-71).
enum { NO_COORD_LEFT = -32768, NO_COORD_RIGHT = 32767, INT_CALC_MULT = 128 };
enum { BITS_PCM_DATA = 128, BITS_IN_LINE = 137, BITS_BETWEEN = 132, PS_STAGES = 5 };
// Binarizer limits (binarizer.h:230-250).
enum { HYST_DEPTH_MIN = 0, HYST_DEPTH_SAFE = 4, HYST_DEPTH_MAX = 10, SHIFT_MIN = 0, SHIFT_SAFE = 2, SHIFT_MAX = 4 };
enum { MAX_COLL_CRCS = 32 };
enum { REF_NO_PCM = 0, REF_BAD_CRC, REF_CRC_COLL, REF_CRC_OK };
enum { SPAN_NOT_FOUND = 0, SPAN_TOO_NARROW, SPAN_OK };
enum { STG_INPUT_ALL = 0, STG_INPUT_LEVEL, STG_REF_FIND, STG_REF_SWEEP_RUN, STG_READ_PCM, STG_DATA_OK, STG_NO_GOOD, STG_MAX };
enum { MARK_ST_START = 0, MARK_ST_TOP_1, MARK_ST_BOT_1, MARK_ST_TOP_2, MARK_ST_BOT_2 };
enum { MARK_ED_START = 0, MARK_ED_TOP, MARK_ED_BOT, MARK_ED_LEN_OK };
enum { MIN_VALID_CRCS = 5 };                    // Binarizer::MIN_VALID_CRCS (binarizer.h:229), the constant of pickLevelByCRCStats
// Fine settings: the numeric fields of bin_preset_t (binarizer.h:163-186; defaults bin_preset_t::reset, binarizer.cpp:48-65),
// set per decode call from the handle (sdv_bin_set_fine_settings).  Device code reads them from constant memory, the host
// build (tests/hostemu) from a plain object.
struct FineSet { u8 max_black_lvl, min_white_lvl, min_contrast, min_ref_lvl, max_ref_lvl, min_valid_crcs, mark_max_dist, left_bit_pick, right_bit_pick, en_coord_search, en_first_line_dup, pad[1]; };
#define SDV_FINE_DEFAULTS { 160, 28, 10, 7, 240, 5, 6, 4, 2, 1, 1, { 0 } }
#if defined(__CUDACC__)
__constant__ FineSet c_fine = SDV_FINE_DEFAULTS;
#endif
static FineSet h_fine = SDV_FINE_DEFAULTS;
#if defined(__CUDA_ARCH__)
#define SDV_FINE c_fine
#else
#define SDV_FINE h_fine
#endif
#define MAX_BLACK_LVL ((int)SDV_FINE.max_black_lvl)
#define MIN_WHITE_LVL ((int)SDV_FINE.min_white_lvl)
#define MIN_CONTRAST ((int)SDV_FINE.min_contrast)
#define MIN_REF_LVL ((int)SDV_FINE.min_ref_lvl)
#define MAX_REF_LVL ((int)SDV_FINE.max_ref_lvl)
#define FINE_MIN_VALID_CRCS ((int)SDV_FINE.min_valid_crcs)
#define MARK_MAX_DIST ((int)SDV_FINE.mark_max_dist)
#define P1_LEFT_BIT_PICK ((int)SDV_FINE.left_bit_pick)
#define P1_RIGHT_BIT_PICK ((int)SDV_FINE.right_bit_pick)
#define FINE_EN_COORD_SEARCH (SDV_FINE.en_coord_search!=0)     // bin_preset_t::en_coord_search (binarizer.cpp:1180,1264)
#define FINE_FIRST_LINE_DUP (SDV_FINE.en_first_line_dup!=0)     // bin_preset_t::en_first_line_dup (videotodigital.cpp:1199)
enum { MARK_TRIALS = 24 };                      // hysteresis trials of findSTC007Coordinates (binarizer.cpp:6047-6113)
enum { MAX_CAND = (HYST_DEPTH_MAX+1)*(SHIFT_MAX+1) };
// VideoToDigital chain (videotodigital.h, videotodigital.cpp:698-1815).
enum { FIELD_INIT = 0, FIELD_NEW, FIELD_SAFE, FIELD_UNSAFE };
enum { COORD_HISTORY_DEPTH = 9, COORD_LONG_HISTORY = 16 };
enum { SDV_MAX_W = 2048, SDV_MAX_H = 1250 };

// ------------------------------------------------------------------------------------------------ cooperative group of threads
// On the device a Cta is the thread block; in tests/hostemu it is a single thread (n = 1) and sync() is a no-op.
struct Cta
{
    int tid, n;
    SDV_HD void sync() const
    {
#if defined(__CUDA_ARCH__)
        __syncthreads();
#endif
    }
};

SDV_HD void hist_inc(u32 *bin)
{
#if defined(__CUDA_ARCH__)
    atomicAdd(bin, 1u);
#else
    (*bin)++;
#endif
}

// Next item of a shared work counter.
SDV_HD int grab_next(int *counter)
{
#if defined(__CUDA_ARCH__)
    return atomicAdd(counter, 1);
#else
    return (*counter)++;
#endif
}

// ------------------------------------------------------------------------------------------------ coordinates
struct Coord { i16 start, stop; };
SDV_HD Coord coord_none() { Coord c; c.start = NO_COORD_LEFT; c.stop = NO_COORD_RIGHT; return c; }
SDV_HD bool coord_valid(Coord c) { return (c.start!=NO_COORD_LEFT)&&(c.stop!=NO_COORD_RIGHT)&&(c.start<c.stop); }
SDV_HD bool coord_eq(Coord a, Coord b) { return (a.start==b.start)&&(a.stop==b.stop); }
// CoordinatePair::operator< (frametrimset.cpp:63-98); the third key (reference) is passed separately.
SDV_HD bool coord_less(Coord a, int ra, Coord b, int rb)
{
    if(a.start<b.start) return true;
    if(a.start==b.start)
    {
        if(a.stop>b.stop) return true;
        if(a.stop==b.stop) return ra<rb;
    }
    return false;
}

// ------------------------------------------------------------------------------------------------ CRC-16 (pcmline.cpp:461-487)
SDV_HD u16 crc16_update(u16 crc, u16 data, int bits)
{
    for(int i=bits-1;i>=0;i--)
    {
        u32 in = (data>>i)&1u;
        u32 msb = (crc>>15)&1u;
        crc = (u16)(crc<<1);
        if(in!=msb) crc ^= 0x1021;
    }
    return crc;
}
// The same CRC advanced over one message byte without a table (x^16+x^12+x^5+1: the byte folds in with two shifts).
SDV_HD u16 crc16_byte(u16 crc, u32 byte)
{
    u32 x = ((u32)(crc>>8)^byte)&0xFFu;
    x ^= x>>4;
    return (u16)(((u32)crc<<8)^(x<<12)^(x<<5)^x);
}
// STC-007 line CRCC: 8 x 14 bits = 14 message bytes.
SDV_HD u16 crc_stc007(const u16 *w8)
{
    const u64 a = ((u64)(w8[0]&0x3FFF)<<42)|((u64)(w8[1]&0x3FFF)<<28)|((u64)(w8[2]&0x3FFF)<<14)|(u64)(w8[3]&0x3FFF);
    const u64 b = ((u64)(w8[4]&0x3FFF)<<42)|((u64)(w8[5]&0x3FFF)<<28)|((u64)(w8[6]&0x3FFF)<<14)|(u64)(w8[7]&0x3FFF);
    u16 c = 0xFFFF;
    for(int k=6;k>=0;k--) c = crc16_byte(c, (u32)(a>>(8*k))&0xFFu);
    for(int k=6;k>=0;k--) c = crc16_byte(c, (u32)(b>>(8*k))&0xFFu);
    return c;
}

// ------------------------------------------------------------------------------------------------ line geometry
// Binarizer::processLine set-up (binarizer.cpp:574-649).
struct Geom
{
    int W;
    u16 scan_end, mark_start_max, mark_end_min, est_ppb;
};
SDV_HD Geom make_geom(int W)
{
    Geom g; g.W = W;
    g.scan_end = (u16)(W-1);
    u16 msm = (u16)(W*MARK_MAX_DIST); msm = msm/100;
    g.mark_end_min = (u16)(g.scan_end-msm);
    g.mark_start_max = msm;
    u32 t = (u32)W*INT_CALC_MULT; t = t/BITS_IN_LINE;
    g.est_ppb = (u16)((t+(INT_CALC_MULT/2))/INT_CALC_MULT);
    return g;
}

// PCMLine::setPPB / calcPPB (pcmline.cpp:223-234,506-519) for the 132 bit cells between the STC-007 data coordinates.
struct Ppb { u32 psm, half; i32 ofs; };
SDV_HD Ppb make_ppb(Coord c)
{
    Ppb p;
    p.psm = (u32)(c.stop-c.start);
    p.psm = (p.psm*INT_CALC_MULT+BITS_BETWEEN/2)/BITS_BETWEEN;
    p.ofs = c.start;
    p.half = (p.psm+1)/2;
    return p;
}
// PCMLine::getVideoPixeBylCalc (pcmline.cpp:249-311) with STC007Line's bit offset 3 (stc007line.cpp:1051-1057).
SDV_HD int pixel_of_bit(Ppb p, int pcm_bit, int shift_px, int pixel_stop)
{
    i32 vp = (i32)(((u32)(pcm_bit+3)*p.psm)+p.half);
    vp = vp/INT_CALC_MULT;
    vp = vp+p.ofs+shift_px;
    if(vp<0) vp = 0;
    else if(vp>=pixel_stop) vp = pixel_stop-1;
    return vp;
}
SDV_HD int pix_shift(int stage) { return (stage==0) ? 0 : ((stage==1) ? 1 : ((stage==2) ? -1 : ((stage==3) ? 2 : -2))); }

SDV_HD u8 get_low_level(u8 lvl, u8 diff) { if(lvl>diff) lvl = (u8)(lvl-diff); else lvl = 1; return lvl; }
SDV_HD u8 get_high_level(u8 lvl, u8 diff) { if(lvl<(255-diff)) lvl = (u8)(lvl+diff); else lvl = 254; return lvl; }
SDV_HD u8 pick_center_ref(u8 bl, u8 wh)
{   // binarizer.cpp:3504-3548
    u8 d = (u8)(wh-bl), r;
    if(d>=MIN_CONTRAST) { d = d/2; r = (u8)(d+bl); if(r<MIN_REF_LVL) r = MIN_REF_LVL; else if(r>MAX_REF_LVL) r = MAX_REF_LVL; }
    else { if(wh<MAX_REF_LVL) r = MAX_REF_LVL; else r = MIN_REF_LVL; }
    return r;
}

// ------------------------------------------------------------------------------------------------ record helpers
// Is the sample of a 14-bit word "almost silent" (stc007line.cpp:582-606, non-M2)?
// 14-bit word -> 16-bit sample: plain (<<2) or the M2 range/sign expansion (stc007line.cpp:286-323, stc007datablock.cpp:507-562).
SDV_HD i16 stc_sample(u16 w, bool m2)
{
    if(!m2) return (i16)(u16)(w<<2);
    if((w&0x2000)==0) return (i16)(u16)(w<<3);
    u16 d = (u16)(w&~0x2000);
    if(w&0x1000) d |= 0xE000;
    return (i16)d;
}
SDV_HD bool words_almost_silent(const u16 *w, bool m2 = false)
{
    int cnt = 0;
    for(int i=0;i<6;i++) { i16 s = stc_sample(w[i], m2); if(!(s>=16)&&!(s<-16)) cnt++; }
    return cnt>=2;
}
// STC007Line::getWordsDiffBitCount: the XOR is truncated to 8 bits (stc007line.cpp:329-356).
SDV_HD int words_diff8(const u16 *a, const u16 *b)
{
    int cnt = 0;
    for(int i=0;i<8;i++)
    {
        u32 d = (u32)((a[i]^b[i])&0xFF);
#if defined(__CUDA_ARCH__)
        cnt += __popc(d);
#else
        cnt += __builtin_popcount(d);
#endif
    }
    return cnt;
}
SDV_HD bool words_control_block(const u16 *w)
{   // stc007line.cpp:493-504
    return (w[0]==0x3333)&&(w[1]==0x0CCC)&&(w[2]==0x3333)&&(w[3]==0x0CCC)&&(w[4]==0x0000)&&((w[7]&0x0FF0)==0x0000);
}

}   // namespace sdv
