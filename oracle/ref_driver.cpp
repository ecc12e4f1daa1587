// oracle/ref_driver.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product library).
//
// C-ABI harness around the UNMODIFIED reference classes, compiled in place from /root/reference
// by oracle/Makefile into oracle/_ref/libsdvref.so (git-ignored).  It lets the parity tests and the
// `bench.py --impl reference` arm run the reference's own CPU path:
//   * VideoToDigital::doBinarize               (videotodigital.cpp:698)   -> per-line records
//   * STC007/PCM1/PCM16X0 DataStitcher         (stc007datastitcher.cpp:7250 ...) -> PCMSamplePair stream,
//     with the assembled lines / data blocks tapped from the newLineProcessed / newBlockProcessed signals
//   * STC007Deinterleaver::processBlock        (stc007deinterleaver.cpp:286) driven directly on records
//   * Binarizer::processLine                   (binarizer.cpp:443) driven directly with explicit presets
//   * the CRC routines of the three line classes (PCMTester KATs, pcmtester.cpp:9-99)
// The 32 Qt signals the reference declares have no moc here; they are defined below as empty bodies or taps.
#include <deque>
#include <vector>
#include <thread>
#include <mutex>
#include <atomic>
#include <cstring>
#include <chrono>
#include "videotodigital.h"
#define private public          /* test harness only: STC007DataStitcher::tryPadding is a private member */
#include "stc007datastitcher.h"
#undef private
#include "pcm1datastitcher.h"
#include "pcm16x0datastitcher.h"
#include "stc007deinterleaver.h"
#include "pcm1deinterleaver.h"
#include "pcm16x0deinterleaver.h"

extern "C" {

// One record per PCM line object the reference emits (64 bytes).
struct sdvref_line_rec
{
    uint32_t frame;
    uint16_t line;
    uint16_t words[9];          // STC-007: 8 words + CRCC; PCM-1: 6 + CRCC; PCM-16x0: 3 + CRCC
    int16_t  data_start, data_stop;
    uint8_t  black, white, ref_low, ref, ref_high, hyst, shift;
    uint8_t  service_type;      // PCMLine::SRVLINE_*
    uint16_t flags;             // bit0 isCRCValid, 1 isCRCValidIgnoreForced, 2 forced_bad, 3 bw_set, 4 coords_set,
                                // 5 ref_sweeped, 6 data_by_ext_tune, 7 coords_sweeped, 8 hasMarkers/hasStartMarker&&hasStopMarker,
                                // 9 start marker, 10 stop marker, 11 control bit (16x0), 12 almost silent
    uint8_t  mark_st_stage, mark_ed_stage;  // STC-007; PCM-1/16x0: picked_bits_left / picked_bits_right
    uint16_t marker_start_bg, marker_start_ed, marker_stop_ed;
    uint16_t word_crc_mask;     // STC-007 per-word isWordCRCOk bits
    uint16_t word_valid_mask;   // STC-007 per-word isWordValid bits
    uint8_t  line_part;         // PCM-16x0
    uint8_t  pcm_type;
    uint16_t queue_order;       // PCM-16x0
    uint8_t  pad[8];
};

// One record per PCMSamplePair (12 bytes).
struct sdvref_pair_rec
{
    int16_t  l, r;
    uint8_t  flags_l, flags_r;  // bit0 data_block_ok, 1 word_valid, 2 word_fixed, 3 word_masked
    uint8_t  service_type;      // PCMSamplePair::SRV_*
    uint8_t  emphasis;
    uint16_t sample_rate;
    uint16_t pad;
};

// One record per STC007DataBlock (48 bytes).
struct sdvref_block_rec
{
    uint16_t words[8];
    uint8_t  line_crc, word_valid, cwd_fixed;   // bit masks over the 8 words
    uint8_t  audio_state, resolution;
    uint8_t  flags;             // bit0 isBlockValid, 1 isDataBroken, 2 isDataFixedByP, 3 isDataFixedByQ, 4 isSilent, 5 isOnSeam
    uint16_t start_line, stop_line;
    uint32_t start_frame, stop_frame;
    int16_t  samples[6];
    uint8_t  pad[2];
};

}   // extern "C"

//------------------------------------------------------------------------------------------------
// Taps (filled from the signal bodies below).
static std::mutex g_tap_mtx;
static std::vector<sdvref_line_rec> *g_tap_asm_lines = NULL;
static std::vector<sdvref_block_rec> *g_tap_blocks = NULL;
static std::vector<FrameAsmSTC007> *g_tap_frasm_stc = NULL;

static void fill_base(sdvref_line_rec *r, PCMLine *l)
{
    memset(r, 0, sizeof(*r));
    r->frame = l->frame_number; r->line = l->line_number;
    r->data_start = l->coords.data_start; r->data_stop = l->coords.data_stop;
    r->black = l->black_level; r->white = l->white_level;
    r->ref_low = l->ref_low; r->ref = l->ref_level; r->ref_high = l->ref_high;
    r->hyst = l->hysteresis_depth; r->shift = l->shift_stage;
    r->pcm_type = l->getPCMType();
    uint8_t st = 0;
    if(l->isServNewFile()) st = PCMLine::SRVLINE_NEW_FILE;
    else if(l->isServEndFile()) st = PCMLine::SRVLINE_END_FILE;
    else if(l->isServFiller()) st = PCMLine::SRVLINE_FILLER;
    else if(l->isServEndField()) st = PCMLine::SRVLINE_END_FIELD;
    else if(l->isServEndFrame()) st = PCMLine::SRVLINE_END_FRAME;
    else if(l->isServiceLine()) st = 0xFF;      // refined by sub-class
    r->service_type = st;
    uint16_t f = 0;
    if(l->isCRCValid()) f |= 1<<0;
    if(l->isCRCValidIgnoreForced()) f |= 1<<1;
    if(l->isForcedBad()) f |= 1<<2;
    if(l->hasBWSet()) f |= 1<<3;
    if(l->hasDataCoordSet()) f |= 1<<4;
    if(l->isDataByRefSweep()) f |= 1<<5;
    if(l->isDataBySkip()) f |= 1<<6;
    if(l->isDataByCoordSweep()) f |= 1<<7;
    if(l->isAlmostSilent()) f |= 1<<12;
    r->flags = f;
}

static void fill_stc(sdvref_line_rec *r, STC007Line *l)
{
    fill_base(r, l);
    for(int i=0;i<9;i++) r->words[i] = l->getWord(i);
    if(l->isServCtrlBlk()) r->service_type = PCMLine::SRVLINE_CTRL_BLOCK;
    if(l->hasMarkers()) r->flags |= 1<<8;
    if(l->hasStartMarker()) r->flags |= 1<<9;
    if(l->hasStopMarker()) r->flags |= 1<<10;
    r->mark_st_stage = l->mark_st_stage; r->mark_ed_stage = l->mark_ed_stage;
    r->marker_start_bg = l->marker_start_bg_coord; r->marker_start_ed = l->marker_start_ed_coord;
    r->marker_stop_ed = l->marker_stop_ed_coord;
    for(int i=0;i<9;i++)
    {
        if(l->isWordCRCOk(i)) r->word_crc_mask |= (1<<i);
        if(l->isWordValid(i)) r->word_valid_mask |= (1<<i);
    }
}

static void fill_pcm1(sdvref_line_rec *r, PCM1Line *l)
{
    fill_base(r, l);
    for(int i=0;i<7;i++) r->words[i] = l->getWord(i);
    if(l->isServHeader()) r->service_type = PCMLine::SRVLINE_HEADER_LINE;
    r->mark_st_stage = l->picked_bits_left; r->mark_ed_stage = l->picked_bits_right;
}

static void fill_pcm16x0(sdvref_line_rec *r, PCM16X0SubLine *l)
{
    fill_base(r, l);
    for(int i=0;i<4;i++) r->words[i] = l->getWord(i);
    r->mark_st_stage = l->picked_bits_left; r->mark_ed_stage = l->picked_bits_right;
    r->line_part = l->line_part; r->queue_order = l->queue_order;
    if(l->control_bit) r->flags |= 1<<11;
}

static void fill_block(sdvref_block_rec *b, STC007DataBlock *d)
{
    memset(b, 0, sizeof(*b));
    for(int i=0;i<8;i++)
    {
        b->words[i] = d->getWord(i);
        if(d->isWordLineCRCOk(i)) b->line_crc |= (1<<i);
        if(d->isWordValid(i)) b->word_valid |= (1<<i);
        if(d->isWordCWDFixed(i)) b->cwd_fixed |= (1<<i);
    }
    b->audio_state = d->getAudioState(); b->resolution = d->getResolution();
    if(d->isBlockValid()) b->flags |= 1;
    if(d->isDataBroken()) b->flags |= 2;
    if(d->isDataFixedByP()) b->flags |= 4;
    if(d->isDataFixedByQ()) b->flags |= 8;
    if(d->isSilent()) b->flags |= 16;
    if(d->isOnSeam()) b->flags |= 32;
    b->start_line = d->getStartLine(); b->stop_line = d->getStopLine();
    b->start_frame = d->getStartFrame(); b->stop_frame = d->getStopFrame();
    for(int i=0;i<6;i++) b->samples[i] = d->getSample(i);
}

//------------------------------------------------------------------------------------------------
// moc substitute: the reference's signals.
void VideoToDigital::guiUpdFrameBin(FrameBinDescriptor) {}
void VideoToDigital::guiUpdFineSettings(bin_preset_t) {}
void VideoToDigital::newLine(STC007Line) {}
void VideoToDigital::newLine(PCM16X0SubLine) {}
void VideoToDigital::newLine(PCM1Line) {}
void VideoToDigital::finished() {}
void VideoToDigital::loopTime(quint64) {}
void PCM1DataStitcher::guiUpdFrameAsm(FrameAsmPCM1) {}
void PCM1DataStitcher::guiUpdFineUseECC(bool) {}
void PCM1DataStitcher::newLineProcessed(PCM1SubLine) {}
void PCM1DataStitcher::newBlockProcessed(PCM1DataBlock) {}
void PCM1DataStitcher::finished() {}
void PCM1DataStitcher::loopTime(quint64) {}
void STC007DataStitcher::guiUpdFrameAsm(FrameAsmSTC007 d)
{
    std::lock_guard<std::mutex> g(g_tap_mtx);
    if(g_tap_frasm_stc) g_tap_frasm_stc->push_back(d);
}
void STC007DataStitcher::guiUpdFineUseECC(bool) {}
void STC007DataStitcher::newLineProcessed(STC007Line l)
{
    std::lock_guard<std::mutex> g(g_tap_mtx);
    if(g_tap_asm_lines) { sdvref_line_rec r; fill_stc(&r, &l); g_tap_asm_lines->push_back(r); }
}
void STC007DataStitcher::newBlockProcessed(STC007DataBlock b)
{
    std::lock_guard<std::mutex> g(g_tap_mtx);
    if(g_tap_blocks) { sdvref_block_rec r; fill_block(&r, &b); g_tap_blocks->push_back(r); }
}
void STC007DataStitcher::guiUpdFineBrokeMask(uint8_t) {}
void STC007DataStitcher::guiUpdFineMaskSeams(bool) {}
void STC007DataStitcher::guiUpdFineMaxUnch14(uint8_t) {}
void STC007DataStitcher::guiUpdFineMaxUnch16(uint8_t) {}
void STC007DataStitcher::guiUpdFineTopLineFix(bool) {}
void STC007DataStitcher::finished() {}
void STC007DataStitcher::loopTime(quint64) {}
void PCM16X0DataStitcher::guiUpdFrameAsm(FrameAsmPCM16x0) {}
void PCM16X0DataStitcher::guiUpdFineUseECC(bool) {}
void PCM16X0DataStitcher::newLineProcessed(PCM16X0SubLine) {}
void PCM16X0DataStitcher::newBlockProcessed(PCM16X0DataBlock) {}
void PCM16X0DataStitcher::guiUpdFineBrokeMask(uint8_t) {}
void PCM16X0DataStitcher::guiUpdFineMaskSeams(bool) {}
void PCM16X0DataStitcher::finished() {}
void PCM16X0DataStitcher::loopTime(quint64) {}

//------------------------------------------------------------------------------------------------
// Input side: reproduce the VideoLine ordering of VideoInFFMPEG::spliceFrame / insertDummyFrame
// (vin_ffmpeg.cpp:213-364, 367-522): [NEW_FILE], per frame: odd rows as lines 1,3,5.., END_FIELD,
// even rows as 2,4,6.., END_FIELD, END_FRAME; at end of file one dummy frame (+END_FILE, END_FRAME).
static void push_frame(std::deque<VideoLine> &q, const uint8_t *luma, int H, int W, int stride, uint32_t frame_no)
{
    VideoLine vl;
    vl.setLength(W);
    uint16_t line_num = 0;
    for(int field=0; field<2; field++)
    {
        int row = field;
        line_num = row+1;
        while(1)
        {
            vl.line_number = line_num;
            vl.frame_number = frame_no;
            vl.setDoubleWidth(false);
            memcpy(vl.pixel_data.data(), luma + (size_t)row*stride, W);
            q.push_back(vl);
            if(row < (H-2)) row += 2;
            else { line_num += 2; VideoLine s; s.setServEndField(); s.frame_number = frame_no; s.line_number = line_num; q.push_back(s); break; }
            line_num += 2;
        }
    }
    line_num += 2;
    VideoLine s; s.setServEndFrame(); s.frame_number = frame_no; s.line_number = line_num; q.push_back(s);
}

static void push_new_file(std::deque<VideoLine> &q, uint32_t frame_no)
{
    VideoLine s; s.setServNewFile("synthetic.avi"); s.frame_number = frame_no; s.line_number = 0; q.push_back(s);
}

// eof_mode: 0 = filler lines (what the reference ingest does at EOF, insertDummyFrame(true,false)),
//           1 = empty lines (dropped-frame style dummy).
static void push_eof(std::deque<VideoLine> &q, int H, int W, uint32_t frame_no, int eof_mode)
{
    VideoLine d;
    d.setServNo();
    d.setLength(W);
    d.setDoubleWidth(false);
    uint16_t line_num = 0;
    for(int field=0; field<2; field++)
    {
        int row = field;
        while(1)
        {
            line_num = row+1;
            d.line_number = line_num;
            d.frame_number = frame_no;
            if(eof_mode==0) d.setServFiller(); else d.setEmpty(true);
            q.push_back(d);
            if(row < (H-2)) row += 2;
            else { line_num += 2; VideoLine s; s.setServEndField(); s.frame_number = frame_no; s.line_number = line_num; q.push_back(s); break; }
        }
    }
    line_num += 2;
    { VideoLine s; s.setServEndFile(); s.frame_number = frame_no; s.line_number = line_num; q.push_back(s); }
    line_num += 2;
    { VideoLine s; s.setServEndFrame(); s.frame_number = frame_no; s.line_number = line_num; q.push_back(s); }
}

struct v2d_cfg { int pcm_type; int mode; int line_dup; int eof_mode; };

// Fine binarization settings for the next runs (VideoToDigital::setFineSettings): set by sdvref_set_fine_settings, defaults otherwise.
static bin_preset_t g_fine_preset;
static void setup_v2d(VideoToDigital &v2d, const v2d_cfg &c)
{
    v2d.setLogLevel(0);
    v2d.setPCMType(c.pcm_type);
    v2d.setBinarizationMode(c.mode);
    v2d.setCheckLineDup(c.line_dup!=0);
    v2d.setFineSettings(g_fine_preset);
}

extern "C" {

// v[0..10] = max_black_lvl, min_white_lvl, min_contrast, min_ref_lvl, max_ref_lvl, min_valid_crcs, mark_max_dist, left_bit_pick,
// right_bit_pick, en_coord_search, en_first_line_dup; v == NULL: bin_preset_t::reset().  Applies to sdvref_v2d_run / sdvref_pipeline_run / sdvref_binarize_lines.
void sdvref_set_fine_settings(const int *v)
{
    g_fine_preset.reset();
    if(!v) return;
    g_fine_preset.max_black_lvl = (uint8_t)v[0]; g_fine_preset.min_white_lvl = (uint8_t)v[1]; g_fine_preset.min_contrast = (uint8_t)v[2];
    g_fine_preset.min_ref_lvl = (uint8_t)v[3]; g_fine_preset.max_ref_lvl = (uint8_t)v[4]; g_fine_preset.min_valid_crcs = (uint8_t)v[5];
    g_fine_preset.mark_max_dist = (uint8_t)v[6]; g_fine_preset.left_bit_pick = (uint8_t)v[7]; g_fine_preset.right_bit_pick = (uint8_t)v[8];
    g_fine_preset.en_coord_search = (v[9]!=0); g_fine_preset.en_first_line_dup = (v[10]!=0);
}

//------------------------------------------------------------------------------------------------
// CRC known-answer helpers (pcmtester.cpp:9-99).
uint16_t sdvref_crc_stc007(const uint16_t *w8)
{
    STC007Line l; for(int i=0;i<8;i++) l.setWord(i, w8[i]); l.calcCRC(); return l.getCalculatedCRC();
}
uint16_t sdvref_crc_pcm1(const uint16_t *w6)
{
    PCM1Line l; for(int i=0;i<6;i++) l.setWord(i, w6[i]); l.calcCRC(); return l.getCalculatedCRC();
}
uint16_t sdvref_crc_pcm16x0(const uint16_t *w3)
{
    PCM16X0SubLine l; for(int i=0;i<3;i++) l.setWord(i, w3[i]); l.calcCRC(); return l.getCalculatedCRC();
}

//------------------------------------------------------------------------------------------------
// Binarizer::processLine driven directly on n independent lines with explicit presets (0 = not preset).
// pcm_type: PCMLine::TYPE_*; mode: Binarizer::MODE_*; part: Binarizer::FULL_LINE.. (PCM-16x0 part).
int sdvref_binarize_lines(int pcm_type, int mode, int part, const uint8_t *luma, int n, int W, int stride,
                          int preset_ref, int preset_black, int preset_white, int preset_start, int preset_stop,
                          sdvref_line_rec *out)
{
    Binarizer bin;
    bin.setFineSettings(g_fine_preset);
    STC007Line stc; PCM1Line p1; PCM16X0SubLine p16;
    VideoLine vl; vl.setLength(W);
    for(int i=0;i<n;i++)
    {
        memcpy(vl.pixel_data.data(), luma+(size_t)i*stride, W);
        vl.frame_number = 1; vl.line_number = i+1; vl.scan_done = false;
        bin.setGoodParameters();
        bin.setMode(mode);
        bin.setLinePartMode(part);
        bin.setCoordinatesSearch(true);
        bin.setSource(&vl);
        if(preset_ref>0) bin.setReferenceLevel(preset_ref);
        if(preset_white>0) bin.setBWLevels(preset_black, preset_white);
        if(preset_stop!=0) bin.setDataCoordinates(preset_start, preset_stop);
        if(pcm_type==PCMLine::TYPE_STC007) { bin.setOutput(&stc); bin.processLine(); fill_stc(&out[i], &stc); }
        else if(pcm_type==PCMLine::TYPE_PCM1) { bin.setOutput(&p1); bin.processLine(); fill_pcm1(&out[i], &p1); }
        else { bin.setOutput(&p16); bin.processLine(); fill_pcm16x0(&out[i], &p16); }
    }
    return n;
}

//------------------------------------------------------------------------------------------------
// VideoToDigital over whole frames.  Returns number of records written (all emitted PCM line objects,
// service lines included), or -1 if max_out is too small.
int sdvref_v2d_run(int pcm_type, int mode, int line_dup, int eof_mode, const uint8_t *luma, int n_frames, int H, int W,
                   sdvref_line_rec *out, int max_out)
{
    std::deque<VideoLine> in_q; QMutex in_m;
    std::deque<STC007Line> q_stc; QMutex m_stc;
    std::deque<PCM1Line> q_p1; QMutex m_p1;
    std::deque<PCM16X0SubLine> q_p16; QMutex m_p16;
    VideoToDigital v2d;
    v2d_cfg c = {pcm_type, mode, line_dup, eof_mode};
    setup_v2d(v2d, c);
    v2d.setInputPointers(&in_q, &in_m);
    v2d.setOutSTC007Pointers(&q_stc, &m_stc);
    v2d.setOutPCM1Pointers(&q_p1, &m_p1);
    v2d.setOutPCM16X0Pointers(&q_p16, &m_p16);
    std::thread th([&](){ v2d.doBinarize(); });
    int n_out = 0; bool done = false, overflow = false;
    int next_frame = 0; bool eof_pushed = false;
    while(!done)
    {
        // Feed input with bounded look-ahead (the real ingest keeps <= 3 frames queued).
        in_m.lock();
        size_t qsz = in_q.size();
        if(qsz < (size_t)(2*(H+4)))
        {
            if(next_frame<n_frames)
            {
                if(next_frame==0) push_new_file(in_q, 1);
                push_frame(in_q, luma+(size_t)next_frame*H*W, H, W, W, next_frame+1);
                next_frame++;
            }
            else if(!eof_pushed)
            {
                if(n_frames==0) push_new_file(in_q, 1);
                push_eof(in_q, H, W, n_frames+1, eof_mode); eof_pushed = true;
            }
        }
        in_m.unlock();
        bool got = false;
        if((pcm_type==VideoToDigital::TYPE_STC007)||(pcm_type==VideoToDigital::TYPE_M2))
        {
            m_stc.lock();
            while(!q_stc.empty())
            {
                got = true;
                if(n_out<max_out) fill_stc(&out[n_out], &q_stc.front()); else overflow = true;
                if(q_stc.front().isServEndFile()) done = true;
                n_out++; q_stc.pop_front();
            }
            m_stc.unlock();
        }
        else if(pcm_type==VideoToDigital::TYPE_PCM1)
        {
            m_p1.lock();
            while(!q_p1.empty())
            {
                got = true;
                if(n_out<max_out) fill_pcm1(&out[n_out], &q_p1.front()); else overflow = true;
                if(q_p1.front().isServEndFile()) done = true;
                n_out++; q_p1.pop_front();
            }
            m_p1.unlock();
        }
        else
        {
            m_p16.lock();
            while(!q_p16.empty())
            {
                got = true;
                if(n_out<max_out) fill_pcm16x0(&out[n_out], &q_p16.front()); else overflow = true;
                if(q_p16.front().isServEndFile()) done = true;
                n_out++; q_p16.pop_front();
            }
            m_p16.unlock();
        }
        if(!got) std::this_thread::sleep_for(std::chrono::microseconds(20));
    }
    v2d.stop();
    th.join();
    return overflow ? -1 : n_out;
}

//------------------------------------------------------------------------------------------------
// Stitcher configuration for the full pipelines.
struct sdvref_stitch_cfg
{
    int video_std;      // FrameAsmDescriptor::VID_* (0 = unknown/auto)
    int field_order;    // FrameAsmDescriptor::ORDER_*
    int resolution;     // STC007DataStitcher::SAMPLE_RES_* preset (0 = auto)
    int p_corr, q_corr, cwd;
    int sample_rate;    // preset (0/1 = auto)
    int pcm16x0_format; // PCM16X0DataStitcher format
    int auto_line_offset;   // PCM-1
    int reserved[7];
};

// Full pipeline: VideoToDigital thread + stitcher thread, main thread feeds frames and drains PCMSamplePairs.
// Taps are optional (NULL/0 to skip).  Returns number of pairs written or -1 on overflow.
int sdvref_pipeline_run(int pcm_type, int mode, int line_dup, int eof_mode, const sdvref_stitch_cfg *sc,
                        const uint8_t *luma, int n_frames, int H, int W,
                        sdvref_pair_rec *pairs, int max_pairs,
                        sdvref_line_rec *asm_lines, int max_asm, int *n_asm,
                        sdvref_block_rec *blocks, int max_blocks, int *n_blocks)
{
    std::deque<VideoLine> in_q; QMutex in_m;
    std::deque<STC007Line> q_stc; QMutex m_stc;
    std::deque<PCM1Line> q_p1; QMutex m_p1;
    std::deque<PCM16X0SubLine> q_p16; QMutex m_p16;
    std::deque<PCMSamplePair> q_out; QMutex m_out;
    std::vector<sdvref_line_rec> tap_lines; std::vector<sdvref_block_rec> tap_blocks;
    {
        std::lock_guard<std::mutex> g(g_tap_mtx);
        g_tap_asm_lines = asm_lines ? &tap_lines : NULL;
        g_tap_blocks = blocks ? &tap_blocks : NULL;
    }
    VideoToDigital v2d;
    v2d_cfg c = {pcm_type, mode, line_dup, eof_mode};
    setup_v2d(v2d, c);
    v2d.setInputPointers(&in_q, &in_m);
    v2d.setOutSTC007Pointers(&q_stc, &m_stc);
    v2d.setOutPCM1Pointers(&q_p1, &m_p1);
    v2d.setOutPCM16X0Pointers(&q_p16, &m_p16);
    STC007DataStitcher st_stc; PCM1DataStitcher st_p1; PCM16X0DataStitcher st_p16;
    std::thread th_st;
    if((pcm_type==VideoToDigital::TYPE_STC007)||(pcm_type==VideoToDigital::TYPE_M2))
    {
        st_stc.setM2SampleFormat(pcm_type==VideoToDigital::TYPE_M2);
        st_stc.setInputPointers(&q_stc, &m_stc);
        st_stc.setOutputPointers(&q_out, &m_out);
        st_stc.setVideoStandard(sc->video_std);
        st_stc.setFieldOrder(sc->field_order);
        st_stc.setResolutionPreset(sc->resolution);
        st_stc.setPCorrection(sc->p_corr!=0);
        st_stc.setQCorrection(sc->q_corr!=0);
        st_stc.setCWDCorrection(sc->cwd!=0);
        if(sc->sample_rate>1) st_stc.setSampleRatePreset(sc->sample_rate);
        th_st = std::thread([&](){ st_stc.doFrameReassemble(); });
    }
    else if(pcm_type==VideoToDigital::TYPE_PCM1)
    {
        st_p1.setInputPointers(&q_p1, &m_p1);
        st_p1.setOutputPointers(&q_out, &m_out);
        st_p1.setFieldOrder(sc->field_order);
        st_p1.setAutoLineOffset(sc->auto_line_offset!=0);
        st_p1.setOddLineOffset((int8_t)sc->reserved[0]);
        st_p1.setEvenLineOffset((int8_t)sc->reserved[1]);
        th_st = std::thread([&](){ st_p1.doFrameReassemble(); });
    }
    else
    {
        st_p16.setInputPointers(&q_p16, &m_p16);
        st_p16.setOutputPointers(&q_out, &m_out);
        st_p16.setFormat(sc->pcm16x0_format);
        st_p16.setFieldOrder(sc->field_order);
        st_p16.setPCorrection(sc->p_corr!=0);
        if(sc->sample_rate>1) st_p16.setSampleRatePreset(sc->sample_rate);
        th_st = std::thread([&](){ st_p16.doFrameReassemble(); });
    }
    std::thread th_v2d([&](){ v2d.doBinarize(); });
    int n_out = 0; bool done = false, overflow = false;
    int next_frame = 0; bool eof_pushed = false;
    while(!done)
    {
        in_m.lock();
        if(in_q.size() < (size_t)(2*(H+4)))
        {
            if(next_frame<n_frames)
            {
                if(next_frame==0) push_new_file(in_q, 1);
                push_frame(in_q, luma+(size_t)next_frame*H*W, H, W, W, next_frame+1);
                next_frame++;
            }
            else if(!eof_pushed)
            {
                if(n_frames==0) push_new_file(in_q, 1);
                push_eof(in_q, H, W, n_frames+1, eof_mode); eof_pushed = true;
            }
        }
        in_m.unlock();
        bool got = false;
        m_out.lock();
        while(!q_out.empty())
        {
            got = true;
            PCMSamplePair &p = q_out.front();
            if(n_out<max_pairs)
            {
                sdvref_pair_rec *r = &pairs[n_out];
                memset(r, 0, sizeof(*r));
                r->l = p.samples[0].audio_word; r->r = p.samples[1].audio_word;
                for(int ch=0;ch<2;ch++)
                {
                    uint8_t f = 0;
                    if(p.samples[ch].data_block_ok) f |= 1;
                    if(p.samples[ch].word_valid) f |= 2;
                    if(p.samples[ch].word_fixed) f |= 4;
                    if(p.samples[ch].word_masked) f |= 8;
                    if(ch==0) r->flags_l = f; else r->flags_r = f;
                }
                r->service_type = p.service_type; r->emphasis = p.emphasis; r->sample_rate = p.sample_rate;
            }
            else overflow = true;
            if(p.isServEndFile()) done = true;
            n_out++; q_out.pop_front();
        }
        m_out.unlock();
        if(!got) std::this_thread::sleep_for(std::chrono::microseconds(20));
    }
    v2d.stop(); st_stc.stop(); st_p1.stop(); st_p16.stop();
    th_v2d.join(); th_st.join();
    {
        std::lock_guard<std::mutex> g(g_tap_mtx);
        g_tap_asm_lines = NULL; g_tap_blocks = NULL;
    }
    if(asm_lines && n_asm)
    {
        *n_asm = (int)tap_lines.size();
        for(size_t i=0;i<tap_lines.size() && (int)i<max_asm;i++) asm_lines[i] = tap_lines[i];
    }
    if(blocks && n_blocks)
    {
        *n_blocks = (int)tap_blocks.size();
        for(size_t i=0;i<tap_blocks.size() && (int)i<max_blocks;i++) blocks[i] = tap_blocks[i];
    }
    return overflow ? -1 : n_out;
}

//------------------------------------------------------------------------------------------------
// STC007Deinterleaver::processBlock driven directly: n line records (words + per-line CRC-ok flag),
// one block per start line s in [0, n-112).  crc_ok[i] bit0 -> line i carries a valid CRC, bit1 -> line has data
// (valid coordinates and B&W levels; this is what replaces the CRC flag when ignore_crc is set).
int sdvref_deint_stc007(const uint16_t *words /*[n][8]*/, const uint8_t *crc_ok, int n, int res_mode,
                        int ignore_crc, int force_check, int p_corr, int q_corr, sdvref_block_rec *out)
{
    std::deque<STC007Line> q;
    for(int i=0;i<n;i++)
    {
        STC007Line l;
        l.frame_number = 1 + i/588; l.line_number = 1 + i%588;
        for(int w=0;w<8;w++) l.setWord(w, words[i*8+w]);
        l.calcCRC();
        l.setSourceCRC(l.getCalculatedCRC());
        if(!(crc_ok[i]&1)) l.setInvalidCRC();
        if(crc_ok[i]&2) { l.coords.setCoordinates(0, 719); l.setBWLevelsState(true); }   // "line has data" (used when CRC is ignored)
        l.applyCRCStatePerWord();
        q.push_back(l);
    }
    STC007Deinterleaver di; STC007DataBlock blk;
    di.setInput(&q); di.setOutput(&blk);
    di.setResMode(res_mode); di.setIgnoreCRC(ignore_crc!=0); di.setForcedErrorCheck(force_check!=0);
    di.setPCorrection(p_corr!=0); di.setQCorrection(q_corr!=0); di.setCWDCorrection(false);
    int nb = 0;
    for(int s=0; s+STC007DataBlock::MIN_DEINT_DATA<n; s++)
    {
        blk.clear();
        di.processBlock(s);
        fill_block(&out[nb++], &blk);
    }
    return nb;
}

// PCM1Deinterleaver::processBlock (pcm1deinterleaver.cpp:69-278) over whole fields of 735 sub-lines.
// lr: [n_fields*735][2] 13-bit words (left, right); flags per sub-line: bit0 CRC valid, bit1 black/white set,
// bit2 picked bits left, bit3 picked bits right.  Per field 8 interleave blocks of 184 (last: 182) words = 1470 words.
// out_samples [n_fields*1470]; out_flags: bit0 isBlockValid, bit1 isWordValid (as PCM1DataStitcher::outputDataBlock
// passes them to PCMSamplePair::setSample, pcm1datastitcher.cpp:1284-1301).
int sdvref_deint_pcm1(const uint16_t *lr, const uint8_t *flags, int n_fields, int ignore_crc, int16_t *out_samples, uint8_t *out_flags)
{
    PCM1Deinterleaver di; PCM1DataBlock blk;
    di.setIgnoreCRC(ignore_crc!=0);
    int o = 0;
    for(int f=0;f<n_fields;f++)
    {
        std::vector<PCM1SubLine> v;
        for(int i=0;i<PCM1DataBlock::MIN_DEINT_DATA;i++)
        {
            PCM1SubLine s;
            size_t k = (size_t)f*PCM1DataBlock::MIN_DEINT_DATA+i;
            s.frame_number = 1+f/2; s.line_number = 1+i/3;
            s.setLeft(lr[2*k]); s.setRight(lr[2*k+1]);
            s.setLinePart(i%3);
            s.setCRCValid((flags[k]&1)!=0);
            s.setBWLevels((flags[k]&2)!=0);
            s.picked_bits_left = (flags[k]&4) ? 1 : 0;
            s.picked_bits_right = (flags[k]&8) ? 1 : 0;
            v.push_back(s);
        }
        di.setInput(&v); di.setOutput(&blk);
        for(int n=0;n<PCM1DataBlock::INT_BLK_PER_FIELD;n++)
        {
            blk.clear();
            if(di.processBlock(n, 0)!=PCM1Deinterleaver::DI_RET_OK) return -1;
            for(uint8_t w=0;w<blk.getWordCount();w++)
            {
                out_samples[o] = blk.getSample(w);
                out_flags[o] = (uint8_t)((blk.isBlockValid() ? 1 : 0)|(blk.isWordValid(w) ? 2 : 0));
                o++;
            }
        }
    }
    return o;
}

// PCM16X0Deinterleaver::processBlock (pcm16x0deinterleaver.cpp:128-708), SI format, driven the way
// PCM16X0DataStitcher::performDeinterleave does (pcm16x0datastitcher.cpp:5204-5346): per interleave block of 105 sub-lines,
// data block i = 0..34 at line_sh = i with even_order = (i odd).
// words [n_itl*105][3]; flags per sub-line: bit0 CRC valid, bit1 "has data" (coordinates valid and black/white set),
// bit3 picked bits right; picked_left [n] = number of picked bits at the left.
// out_samples [n_itl*35][6] = (L,R) of sub-blocks 1..3; out_flags [..][6]: bit0 block state, bit1 word valid, bit2 word
// "fixed" flag exactly as PCM16X0DataStitcher::outputDataBlock computes them (4998-5085); out_state [..][3] audio state.
static int deint_pcm16x0_fmt(const uint16_t *words, const uint8_t *flags, const uint8_t *picked_left, int n_itl, int ignore_crc,
                             int force_check, int p_corr, int ei, int16_t *out_samples, uint8_t *out_flags, uint8_t *out_state)
{
    PCM16X0Deinterleaver di; PCM16X0DataBlock blk;
    di.setIgnoreCRC(ignore_crc!=0); di.setForcedErrorCheck(force_check!=0); di.setPCorrection(p_corr!=0);
    if(ei) di.setEIFormat(); else di.setSIFormat();
    const int unit = ei ? 1470 : 105, nblk = ei ? 490 : 35;      // sub-lines / data blocks per interleave unit
    int o = 0;
    for(int m=0;m<n_itl;m++)
    {
        std::vector<PCM16X0SubLine> v;
        for(int i=0;i<unit;i++)
        {
            size_t k = (size_t)m*unit+i;
            PCM16X0SubLine s;
            s.frame_number = 1; s.line_number = (uint16_t)(1+k/3); s.line_part = (uint8_t)(k%3); s.queue_order = (uint16_t)i;
            for(int w=0;w<3;w++) s.setWord(w, words[3*k+w]);
            s.calcCRC();
            s.setSourceCRC(s.getCalculatedCRC());
            if(!(flags[k]&1)) s.setInvalidCRC();
            if(flags[k]&2) { s.coords.setCoordinates(8, 712); s.setBWLevelsState(true); }
            s.picked_bits_left = picked_left[k];
            s.picked_bits_right = (flags[k]&8) ? 1 : 0;
            v.push_back(s);
        }
        di.setInput(&v); di.setOutput(&blk);
        bool even_order = false;
        for(int i=0;i<nblk;i++)
        {
            blk.clear();
            if(di.processBlock((uint16_t)i, even_order)!=PCM16X0Deinterleaver::DI_RET_OK) return -1;
            for(uint8_t sb=0;sb<3;sb++)
            {
                bool bstate, lv, rv, lf, rf;
                if(!blk.isDataBroken(sb))
                {
                    bstate = blk.isBlockValid();
                    if(!bstate) lf = rf = false;
                    else { lf = blk.isWordCRCOk(sb, PCM16X0DataBlock::WORD_L); rf = blk.isWordCRCOk(sb, PCM16X0DataBlock::WORD_R); }
                    lv = blk.isWordValid(sb, PCM16X0DataBlock::WORD_L); rv = blk.isWordValid(sb, PCM16X0DataBlock::WORD_R);
                }
                else { bstate = lv = rv = lf = rf = false; }
                out_samples[o*6+2*sb] = blk.getSample(sb, PCM16X0DataBlock::WORD_L);
                out_samples[o*6+2*sb+1] = blk.getSample(sb, PCM16X0DataBlock::WORD_R);
                out_flags[o*6+2*sb] = (uint8_t)((bstate ? 1 : 0)|(lv ? 2 : 0)|(lf ? 4 : 0));
                out_flags[o*6+2*sb+1] = (uint8_t)((bstate ? 1 : 0)|(rv ? 2 : 0)|(rf ? 4 : 0));
                out_state[o*3+sb] = blk.getAudioState(sb);
            }
            o++;
            even_order = !even_order;
        }
    }
    return o;
}

int sdvref_deint_pcm16x0(const uint16_t *words, const uint8_t *flags, const uint8_t *picked_left, int n_itl, int ignore_crc,
                         int force_check, int p_corr, int16_t *out_samples, uint8_t *out_flags, uint8_t *out_state)
{
    return deint_pcm16x0_fmt(words, flags, picked_left, n_itl, ignore_crc, force_check, p_corr, 0, out_samples, out_flags, out_state);
}
// EI format: units of 1470 sub-lines (one frame), data block i from sub-lines i, i+490, i+980 (pcm16x0datablock.h:41,70-72).
int sdvref_deint_pcm16x0_ei(const uint16_t *words, const uint8_t *flags, const uint8_t *picked_left, int n_units, int ignore_crc,
                            int force_check, int p_corr, int16_t *out_samples, uint8_t *out_flags, uint8_t *out_state)
{
    return deint_pcm16x0_fmt(words, flags, picked_left, n_units, ignore_crc, force_check, p_corr, 1, out_samples, out_flags, out_state);
}

// STC007DataStitcher::tryPadding (stc007datastitcher.cpp:1417-1740) for paddings 0..n_pad-1 on one field seam.
// Field line arrays as in sdvref_deint_stc007 (words [n][8], crc_ok [n] bit0).  out [n_pad][6] = index, valid, silent,
// unchecked, broken (FieldStitchStats), return code (DS_RET_*).
int sdvref_try_padding(const uint16_t *w1, const uint8_t *ok1, int n1, const uint16_t *w2, const uint8_t *ok2, int n2,
                       int n_pad, int p_corr, int q_corr, uint16_t *out)
{
    STC007DataStitcher st;
    st.setPCorrection(p_corr!=0); st.setQCorrection(q_corr!=0); st.setCWDCorrection(false);
    std::vector<STC007Line> f1, f2;
    for(int pass=0;pass<2;pass++)
    {
        const uint16_t *w = pass ? w2 : w1; const uint8_t *ok = pass ? ok2 : ok1; int n = pass ? n2 : n1;
        for(int i=0;i<n;i++)
        {
            STC007Line l;
            l.frame_number = 1; l.line_number = (uint16_t)(1+2*i+pass);
            for(int k=0;k<8;k++) l.setWord(k, w[i*8+k]);
            l.calcCRC();
            l.setSourceCRC(l.getCalculatedCRC());
            if(!(ok[i]&1)) l.setInvalidCRC();
            l.applyCRCStatePerWord();
            (pass ? f2 : f1).push_back(l);
        }
    }
    for(int pad=0;pad<n_pad;pad++)
    {
        FieldStitchStats fs;
        uint8_t rc = st.tryPadding(&f1, (uint16_t)n1, &f2, (uint16_t)n2, (uint16_t)pad, &fs);
        out[pad*6+0] = fs.index; out[pad*6+1] = fs.valid; out[pad*6+2] = fs.silent; out[pad*6+3] = fs.unchecked; out[pad*6+4] = fs.broken;
        out[pad*6+5] = rc;
    }
    return n_pad;
}

// STC007DataStitcher::findPadding (stc007datastitcher.cpp:1743-2054) on one field seam.  out[0..2] = padding, DS_RET_* code,
// last_pad_counter.  video_std: FrameAsmDescriptor::VID_PAL (1) / VID_NTSC (2); resolution: STC007DataBlock::RES_14BIT / RES_16BIT.
int sdvref_find_padding(const uint16_t *w1, const uint8_t *ok1, int n1, const uint16_t *w2, const uint8_t *ok2, int n2,
                        int video_std, int resolution_16bit, int p_corr, int q_corr, uint16_t *out)
{
    STC007DataStitcher st;
    st.setPCorrection(p_corr!=0); st.setQCorrection(q_corr!=0); st.setCWDCorrection(false);
    std::vector<STC007Line> f1, f2;
    for(int pass=0;pass<2;pass++)
    {
        const uint16_t *w = pass ? w2 : w1; const uint8_t *ok = pass ? ok2 : ok1; int n = pass ? n2 : n1;
        for(int i=0;i<n;i++)
        {
            STC007Line l;
            l.frame_number = 1; l.line_number = (uint16_t)(1+2*i+pass);
            for(int k=0;k<8;k++) l.setWord(k, w[i*8+k]);
            l.calcCRC();
            l.setSourceCRC(l.getCalculatedCRC());
            if(!(ok[i]&1)) l.setInvalidCRC();
            l.applyCRCStatePerWord();
            (pass ? f2 : f1).push_back(l);
        }
    }
    uint16_t padding = 0xFFFF;
    uint8_t rc = st.findPadding(&f1, (uint16_t)n1, &f2, (uint16_t)n2, (uint8_t)video_std,
                                (uint8_t)(resolution_16bit ? STC007DataBlock::RES_16BIT : STC007DataBlock::RES_14BIT), &padding);
    out[0] = padding; out[1] = rc; out[2] = st.last_pad_counter;
    return rc;
}

int sdvref_sizeof_line_rec() { return (int)sizeof(sdvref_line_rec); }
int sdvref_sizeof_pair_rec() { return (int)sizeof(sdvref_pair_rec); }
int sdvref_sizeof_block_rec() { return (int)sizeof(sdvref_block_rec); }

}   // extern "C"
