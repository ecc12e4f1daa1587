/* oracle/sdv_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C CPU restatement of the reference's per-frame decode hot path.  It is the
 * checker for the CUDA path: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may link or
 * call it; the product library (sdvpcmdecoder_b200/csrc) never does.
 *
 * Parity pinning: every function here is differentially tested against the UNMODIFIED reference compiled from
 * /root/reference into oracle/_ref/libsdvref.so (tests/test_oracle_vs_ref.py), against the PCMTester
 * known-answer vectors (pcmtester.cpp:14-82) and against fixtures generated from that build (tests/golden/).
 */
#ifndef SDV_ORACLE_H
#define SDV_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Same layout as sdvref_line_rec (oracle/ref_driver.cpp) so that tests can compare field by field. */
typedef struct
{
    uint32_t frame;
    uint16_t line;
    uint16_t words[9];
    int16_t  data_start, data_stop;
    uint8_t  black, white, ref_low, ref, ref_high, hyst, shift;
    uint8_t  service_type;
    uint16_t flags;
    uint8_t  mark_st_stage, mark_ed_stage;
    uint16_t marker_start_bg, marker_start_ed, marker_stop_ed;
    uint16_t word_crc_mask;
    uint16_t word_valid_mask;
    uint8_t  line_part;
    uint8_t  pcm_type;
    uint16_t queue_order;
    uint8_t  pad[8];
} sdvo_line_rec;

/* Same layout as sdvref_block_rec. */
typedef struct
{
    uint16_t words[8];
    uint8_t  line_crc, word_valid, cwd_fixed;
    uint8_t  audio_state, resolution;
    uint8_t  flags;
    uint16_t start_line, stop_line;
    uint32_t start_frame, stop_frame;
    int16_t  samples[6];
    uint8_t  pad[2];
} sdvo_block_rec;

enum { SDVO_TYPE_PCM1 = 0, SDVO_TYPE_PCM16X0 = 1, SDVO_TYPE_STC007 = 2 };
enum { SDVO_MODE_DRAFT = 0, SDVO_MODE_FAST = 1, SDVO_MODE_NORMAL = 2, SDVO_MODE_INSANE = 3 };

/* CRC-16 (poly 0x1021, init 0xFFFF, MSB first) of the three line formats (pcmline.cpp:461-487). */
uint16_t sdvo_crc_stc007(const uint16_t *w8);
uint16_t sdvo_crc_pcm1(const uint16_t *w6);
uint16_t sdvo_crc_pcm16x0(const uint16_t *w3);

/* Binarizer::processLine (binarizer.cpp:443) on n independent STC-007 lines with explicit presets (0 = none). */
int sdvo_binarize_lines_stc007(int mode, const uint8_t *luma, int n, int W, int stride,
                               int preset_ref, int preset_black, int preset_white, int preset_start, int preset_stop,
                               sdvo_line_rec *out);

/* VideoToDigital::doBinarize (videotodigital.cpp:698) for STC-007 over whole frames u8[F][H][W]:
 * emits one record per video line (stream order: odd field rows, then even field rows), no service lines. */
int sdvo_v2d_stc007(int mode, int line_dup, const uint8_t *luma, int n_frames, int H, int W, sdvo_line_rec *out);

/* STC007Deinterleaver::processBlock (stc007deinterleaver.cpp:286) for each start line s in [0, n-112). */
int sdvo_deint_stc007(const uint16_t *words, const uint8_t *crc_ok, int n, int res_mode,
                      int ignore_crc, int force_check, int p_corr, int q_corr, sdvo_block_rec *out);

/* PCM1Deinterleaver::processBlock (pcm1deinterleaver.cpp:69) over n_fields x 735 sub-lines: lr [n][2] 13-bit words,
 * flags bit0 CRC valid, bit1 black/white set; out: 1470 samples per field + flags (bit0 block valid, bit1 word valid). */
int sdvo_deint_pcm1(const uint16_t *lr, const uint8_t *flags, int n_fields, int ignore_crc, int16_t *out_samples, uint8_t *out_flags);

/* PCM16X0Deinterleaver::processBlock (pcm16x0deinterleaver.cpp:128), SI format, for the 35 data blocks of each of n_itl
 * interleave blocks of 105 sub-lines.  words [n][3]; flags bit0 CRC valid, bit1 has data, bit3 picked right; picked_left [n]
 * = picked bit counts.  out: 6 samples + 6 flags per data block (bit0 block state, bit1 word valid, bit2 fixed flag), 3 audio states. */
int sdvo_deint_pcm16x0_ei(const uint16_t *words, const uint8_t *flags, const uint8_t *picked_left, int n_units, int ignore_crc,
                          int force_check, int p_corr, int16_t *out_samples, uint8_t *out_flags, uint8_t *out_state);
int sdvo_deint_pcm16x0(const uint16_t *words, const uint8_t *flags, const uint8_t *picked_left, int n_itl, int ignore_crc,
                       int force_check, int p_corr, int16_t *out_samples, uint8_t *out_flags, uint8_t *out_state);

/* STC007DataStitcher::tryPadding (stc007datastitcher.cpp:1417) for paddings 0..n_pad-1; out [n_pad][6] = index, valid,
 * silent, unchecked, broken, DS_RET_* code. */
int sdvo_try_padding(const uint16_t *w1, const uint8_t *ok1, int n1, const uint16_t *w2, const uint8_t *ok2, int n2,
                     int n_pad, int res_mode, int ignore_crc, int p_corr, int q_corr, int lim14, int lim16, uint16_t *out);

/* STC007DataStitcher::findPadding (stc007datastitcher.cpp:1743-2054); out[0..2] = padding, DS_RET_* code, last_pad_counter. */
int sdvo_find_padding(const uint16_t *w1, const uint8_t *ok1, int n1, const uint16_t *w2, const uint8_t *ok2, int n2,
                      int video_std, int resolution_16bit, int res_mode, int ignore_crc, int p_corr, int q_corr,
                      int lim14, int lim16, uint16_t *out);

#ifdef __cplusplus
}
#endif
#endif
