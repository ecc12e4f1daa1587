// stc007_cwd.cuh -- Cross-Word Decoding (CWD) of the STC-007 stitcher.
//
// STC007DataStitcher::prescanFrame (stc007datastitcher.cpp:6401-6452) runs performCWD (5905-6398) over the frame that
// was just queued until a pass repairs nothing more: every data block of the queue is deinterleaved (P / Q correction,
// parity check forced); a block that comes out valid hands its corrected words back to the LINES they came from, and a
// line whose words have all been confirmed that way becomes a valid line for the blocks that read its other words --
// blocks with more than two erasures become correctable.  The deinterleaver then runs with its CWD stage enabled
// (STC007Deinterleaver::processBlock, stc007deinterleaver.cpp:286-1123: STG_CWD_CORR 638-712).
//
// Dependencies.  Block b reads and patches only lines b + 16 k: the blocks of one residue class modulo 16 form a
// sequential chain, the 16 classes never touch the same line -- one pass is 16 independent sequential walks, and the passes
// are joined by the reference's termination rule (another pass while ANY class repaired a line).  Frames are chained through
// the 112 lines the queue keeps for the next frame (they stay patched).  A frame without a single line that CWD may patch
// (CRC wrong, data coordinates valid, not forced bad) is untouched by it and hands its lines on as they are, so the tape
// falls into independent chains of consecutive "dirty" frames; each chain is walked by one thread block.
#pragma once
#include "stc007_stitch.cuh"

namespace sdv {

// ------------------------------------------------------------------------------------------------ block logic with the CWD stage
struct BlockX { Block b; u8 cwd_fixed; u8 cwd_applied; };     // STC007DataBlock::cwd_fixed[] as a mask, cwd_applied
SDV_HD bool blkx_fixed_by_cwd(const BlockX *x) { return x->cwd_applied&&(x->cwd_fixed!=0); }      // isDataFixedByCWD
SDV_HD void blkx_mark_broken(BlockX *x) { blk_mark_broken(&x->b); x->cwd_fixed &= (u8)~blk_word_limit_mask(&x->b); x->cwd_applied = 0; }
SDV_HD void blkx_mark_unsafe(BlockX *x)
{
    if(x->b.audio_state==SDV_AUD_BROKEN) return;
    blk_mark_unsafe(&x->b); x->cwd_fixed &= (u8)~blk_word_limit_mask(&x->b); x->cwd_applied = 0;
}
SDV_HD u8 bit_of(u8 i) { return (i<W_CNT) ? (u8)(1u<<i) : (u8)0; }

// STC007Deinterleaver::processBlock with en_cwd (cwd_lines: bit k = line s+16k isFixedByCWD()).
SDV_HD void deint_block_cwd(BlockX *x, const BlockIn *in, u8 cwd_lines, DeintCfg cfg, bool en_cwd)
{
    Block *blk = &x->b;
    u8 run_res, stage_count = 0, fill_passes, all_errs = 0, aud_errs = 0, first_bad = NO_ERR_INDEX, second_bad = NO_ERR_INDEX, fix_result, st;
    if(cfg.res_mode==SDV_RES_MODE_14BIT) { run_res = RES_14BIT; fill_passes = DI_MAX_PASSES; }
    else if(cfg.res_mode==SDV_RES_MODE_14BIT_AUTO) { run_res = RES_14BIT; fill_passes = 0; }
    else if(cfg.res_mode==SDV_RES_MODE_16BIT_AUTO) { run_res = RES_16BIT; fill_passes = 0; }
    else { run_res = RES_16BIT; fill_passes = DI_MAX_PASSES; }
    st = DSTG_DATA_FILL;
    x->cwd_fixed = 0; x->cwd_applied = 0;
    for(;;)
    {
        stage_count++;
        if(st==DSTG_DATA_FILL)
        {
            blk_fill(blk, in, run_res);
            x->cwd_fixed = (run_res==RES_14BIT) ? cwd_lines : (u8)(cwd_lines&0x7F);
            x->cwd_applied = 0;
            fill_passes++;
            st = DSTG_ERROR_CHECK;
        }
        else if(st==DSTG_ERROR_CHECK)
        {
            first_bad = second_bad = NO_ERR_INDEX;
            {
                u32 bad = (u32)(~blk->line_crc)&0x3Fu;
                if(bad) { first_bad = (u8)lowest_bit(bad); bad &= bad-1; if(bad) second_bad = (u8)lowest_bit(bad); }
            }
            aud_errs = (u8)popc8((u32)(~blk->line_crc)&0x3Fu);
            all_errs = (u8)popc8((u32)(~blk->line_crc)&blk_word_limit_mask(blk));
            st = DSTG_TASK_SELECTION;
        }
        else if(st==DSTG_TASK_SELECTION)
        {
            st = DSTG_BAD_BLOCK;
            if(all_errs<=2)
            {
                if(aud_errs==0)
                {
                    if(!cfg.force_check) st = DSTG_DATA_OK;
                    else if(cfg.p_corr) st = DSTG_P_CORR;
                    else st = DSTG_NO_CHECK;
                }
                else if(aud_errs==1) { if(cfg.p_corr) st = DSTG_P_CORR; }
                else if(aud_errs==2)
                {
                    if(run_res==RES_14BIT) { if(cfg.q_corr) st = DSTG_Q_CORR; }
                    else if(en_cwd&&!blkx_fixed_by_cwd(x)) st = DSTG_CWD_CORR;
                }
            }
            else if(en_cwd&&!x->cwd_applied) st = DSTG_CWD_CORR;
        }
        else if(st==DSTG_CWD_CORR)
        {   // words of lines that an earlier CWD pass made valid count as valid; then the ordinary correction once more
            st = DSTG_BAD_BLOCK;
            if(x->cwd_fixed)
            {
                blk->word_valid |= x->cwd_fixed;
                x->cwd_applied = 1;
                first_bad = second_bad = NO_ERR_INDEX;
                const u32 bad_all = (u32)(~blk->word_valid)&0xFFu;
                u32 bad = bad_all&0x3Fu;
                all_errs = (u8)popc8(bad_all); aud_errs = (u8)popc8(bad);
                if(bad) { first_bad = (u8)lowest_bit(bad); bad &= bad-1; if(bad) second_bad = (u8)lowest_bit(bad); }
                st = DSTG_TASK_SELECTION;
            }
        }
        else if(st==DSTG_P_CORR)
        {
            st = DSTG_BAD_BLOCK;
            if(blk_valid(blk, W_P0))
            {
                fix_result = blk_fix_by_p(blk, first_bad);
                if(fix_result==FIX_BROKEN) blkx_mark_broken(x);
                else
                {
                    st = DSTG_DATA_OK;
                    x->cwd_fixed &= (u8)~bit_of(first_bad);
                    if(fix_result==FIX_DONE) blk->audio_state = SDV_AUD_FIX_P;
                    else if(fix_result==FIX_NOT_NEED) { if(first_bad<W_P0) blk->audio_state = SDV_AUD_FIX_P; }
                    if((run_res==RES_14BIT)&&cfg.q_corr)
                    {
                        if(blk_valid(blk, W_Q0))
                        {
                            if(cfg.force_check) { if(blk_synd_q(blk)!=0) { st = DSTG_BAD_BLOCK; blkx_mark_broken(x); } }
                        }
                        else
                        {
                            u16 q = blk_calc_q(blk);
                            if(blk->words[W_Q0]!=q) blk_set_word(blk, W_Q0, q, blk_crc(blk, W_Q0));
                            blk_set_valid(blk, W_Q0);
                            x->cwd_fixed &= (u8)~bit_of(W_Q0);
                        }
                    }
                }
            }
            else
            {
                if(run_res==RES_14BIT)
                {
                    if(cfg.q_corr) st = DSTG_Q_CORR;
                    else if(aud_errs==0) st = DSTG_NO_CHECK;
                }
                else if(aud_errs==0) st = DSTG_NO_CHECK;
            }
        }
        else if(st==DSTG_Q_CORR)
        {
            st = DSTG_BAD_BLOCK;
            if(blk_valid(blk, W_Q0))
            {
                fix_result = blk_fix_by_q(blk, first_bad, second_bad);
                if(!blk_crc(blk, W_P0)) second_bad = W_P0;
                if((fix_result==FIX_DONE)||(fix_result==FIX_NOT_NEED))
                {
                    st = DSTG_DATA_OK;
                    x->cwd_fixed &= (u8)~(bit_of(first_bad)|bit_of(second_bad));
                    if((fix_result==FIX_DONE)||(first_bad<W_P0)) blk->audio_state = SDV_AUD_FIX_Q;
                }
                else if(fix_result==FIX_SWITCH_P) st = DSTG_P_CORR;
                else if(fix_result==FIX_BROKEN) blkx_mark_broken(x);
            }
            else if(first_bad==NO_ERR_INDEX)
            {
                st = DSTG_NO_CHECK;
                blk_set_word(blk, W_P0, blk_calc_p(blk), false); blk_set_valid(blk, W_P0);
                blk_set_word(blk, W_Q0, blk_calc_q(blk), false); blk_set_valid(blk, W_Q0);
            }
        }
        else if(st==DSTG_BAD_BLOCK)
        {
            x->cwd_applied = 0;
            if(fill_passes>=DI_MAX_PASSES) break;
            run_res = (run_res==RES_16BIT) ? RES_14BIT : RES_16BIT;
            st = DSTG_DATA_FILL;
        }
        else break;
        if(stage_count>(DSTG_CONVERT_MAX*DI_MAX_PASSES)) break;
    }
    blk->m2 = cfg.m2;
}

// ------------------------------------------------------------------------------------------------ a queued line
// STC007Line as the stitcher's conv_queue holds it: the nine words, the per-word flags CWD changes, where it came from.
enum { CL_FORCED_BAD = 1, CL_COORDS = 2, CL_CRC_OK = 4 /* isCRCValidIgnoreForced */, CL_BW = 8 /* hasBWSet */ };
struct CwdLine
{
    u16 w[9];
    u16 crc_mask, valid_mask;       // word_crc[], word_valid[] (bits 0..8)
    u8  flags, pad;
    u16 line;                       // line number
    u16 pad2;                       // (explicit: the speculative walk compares lines byte by byte)
    i32 frame;                      // frame number
};
SDV_HD bool cl_crc_valid(const CwdLine &l) { return ((l.flags&CL_FORCED_BAD)==0)&&((l.flags&CL_CRC_OK)!=0); }                 // isCRCValid
SDV_HD bool cl_word_crc(const CwdLine &l, int i) { return ((l.flags&CL_FORCED_BAD)==0)&&(((l.crc_mask>>i)&1)!=0); }           // isWordCRCOk
SDV_HD bool cl_word_valid(const CwdLine &l, int i) { return ((l.flags&CL_FORCED_BAD)==0)&&(((l.valid_mask>>i)&1)!=0); }       // isWordValid
SDV_HD bool cl_fixed_by_cwd(const CwdLine &l) { return cl_crc_valid(l)&&(((~l.crc_mask)&l.valid_mask&0xFF)!=0); }             // isFixedByCWD
SDV_HD void cl_calc_crc(CwdLine *l) { if(crc_stc007(l->w)==l->w[8]) l->flags |= CL_CRC_OK; else l->flags &= (u8)~CL_CRC_OK; }  // calcCRC + the comparison that follows it
SDV_HD void cl_set_word(CwdLine *l, int i, u16 w, bool ok)
{   // STC007Line::setWord
    l->w[i] = (u16)(w&((i==8) ? 0xFFFF : 0x3FFF));
    const u16 m = (u16)(1u<<i);
    if(ok) { l->crc_mask |= m; l->valid_mask |= m; } else { l->crc_mask &= (u16)~m; l->valid_mask &= (u16)~m; }
}
SDV_HD CwdLine cl_empty(i32 frame, int line)
{   // an empty line of addFieldPadding: silent words, inverted CRC, coordinates zeroed
    CwdLine l; for(int i=0;i<9;i++) l.w[i] = 0;
    l.crc_mask = l.valid_mask = 0; l.flags = 0; l.pad = 0; l.pad2 = 0; l.line = (u16)line; l.frame = frame;
    return l;
}
SDV_HD CwdLine cl_from_rec(const sdv_line_rec *r, i32 frame, int line)
{
    if(!r) return cl_empty(frame, line);
    CwdLine l;
    for(int i=0;i<9;i++) l.w[i] = r->words[i];
    l.flags = (u8)(((r->flags&SDV_LF_FORCED_BAD) ? CL_FORCED_BAD : 0)|((r->flags&SDV_LF_CRC_OK_IGN) ? CL_CRC_OK : 0)|((r->flags&SDV_LF_BW_SET) ? CL_BW : 0));
    Coord c; c.start = r->data_start; c.stop = r->data_stop;
    if(coord_valid(c)) l.flags |= CL_COORDS;
    l.crc_mask = l.valid_mask = (r->flags&SDV_LF_CRC_OK) ? 0x1FF : 0;      // applyCRCStatePerWord
    l.pad = 0; l.pad2 = 0; l.line = (u16)line; l.frame = frame;
    return l;
}
// May CWD write into this line?  (stc007datastitcher.cpp:5969-5972, without the frame test)
SDV_HD bool rec_cwd_patchable(const sdv_line_rec *r)
{
    if(r->service_type!=SDV_SRV_NO) return false;
    Coord c; c.start = r->data_start; c.stop = r->data_stop;
    return ((r->flags&(SDV_LF_CRC_OK_IGN|SDV_LF_FORCED_BAD))==0)&&coord_valid(c);
}

// Block s of a queue of CwdLines.
SDV_HD void cwd_block_in(const CwdLine *q, int s, bool ignore_crc, BlockIn *in, u8 *cwd_lines)
{
    in->ok = 0; *cwd_lines = 0;
    for(int k=0;k<8;k++)
    {
        const CwdLine &l = q[s+16*k];
        in->w[k] = l.w[k]; in->sw[k] = l.w[7];
        const bool ok = ignore_crc ? (((l.flags&CL_COORDS)!=0)&&((l.flags&CL_BW)!=0)) : cl_word_crc(l, k);
        if(ok) in->ok |= (u8)(1u<<k);
        if(cl_fixed_by_cwd(l)) *cwd_lines |= (u8)(1u<<k);
    }
}

// performCWD's treatment of one block (stc007datastitcher.cpp:5943-6380): returns the number of lines it made valid.
SDV_HD int cwd_patch_block(CwdLine *q, int s, DeintCfg cfg, i32 next_frame)
{
    BlockIn in; u8 cwd_lines;
    cwd_block_in(q, s, cfg.ignore_crc!=0, &in, &cwd_lines);
    BlockX x;
    deint_block_cwd(&x, &in, cwd_lines, cfg, true);
    const Block &b = x.b;
    if(!blk_block_valid(&b)) return 0;
    if((((u32)(~b.line_crc))&b.word_valid&0xFFu)==0) return 0;                  // isDataFixed
    const int max_fixable = ((!cfg.q_corr)||(b.resolution==RES_16BIT)) ? W_P0 : W_Q0;
    int fixed_lines = 0;
    for(int wi=0;wi<=max_fixable;wi++)
    {
        if(blk_crc(&b, wi)) continue;
        CwdLine *l = &q[s+16*wi];
        const u16 bw = blk_get_word(&b, wi);
        if(((l->flags&CL_CRC_OK)==0)&&((l->flags&CL_COORDS)!=0)&&((l->flags&CL_FORCED_BAD)==0)&&(l->frame!=next_frame))
        {
            if(b.resolution==RES_14BIT)
            {
                if(l->w[wi]!=bw)
                {
                    cl_set_word(l, wi, bw, cl_word_crc(*l, wi));
                    cl_calc_crc(l);
                    l->valid_mask |= (u16)(1u<<wi);
                    if(l->flags&CL_CRC_OK) { l->valid_mask |= 0x1FF; fixed_lines++; }
                }
                else l->valid_mask |= (u16)(1u<<wi);
                if((l->flags&CL_CRC_OK)==0)
                {   // every word confirmed by its block: the line is what was recorded, its CRCC was the damaged part
                    if((l->valid_mask&0xFF)==0xFF)
                    {
                        l->w[8] = crc_stc007(l->w); l->flags |= CL_CRC_OK;
                        l->valid_mask |= 0x100;
                        fixed_lines++;
                    }
                }
            }
            else
            {   // 16-bit word: 14 bits in the word, 2 in the line's S word
                const int sh = 12-2*wi;
                const u16 new_word = (u16)(bw>>2), new_bits = (u16)((bw&3u)<<sh);
                const u16 old_bits = (u16)(l->w[7]&(3u<<sh));
                if(l->w[wi]!=new_word)
                {
                    cl_set_word(l, wi, new_word, cl_word_crc(*l, wi));
                    cl_calc_crc(l);
                    l->valid_mask |= (u16)(1u<<wi);
                    if(l->flags&CL_CRC_OK) { l->valid_mask |= 0x1FF; fixed_lines++; }
                }
                if(((l->flags&CL_CRC_OK)==0)&&(old_bits!=new_bits))
                {
                    cl_set_word(l, 7, (u16)((l->w[7]&~(3u<<sh))|new_bits), cl_word_crc(*l, 7));
                    cl_calc_crc(l);
                    if(l->flags&CL_CRC_OK) { l->valid_mask |= 0x1FF; fixed_lines++; }
                }
            }
        }
        else if(cl_crc_valid(*l)&&(b.resolution==RES_14BIT)&&(l->w[wi]!=bw)) l->flags |= CL_FORCED_BAD;     // a "valid" line contradicted by a valid block
    }
    return fixed_lines;
}

// ------------------------------------------------------------------------------------------------ one chain of frames
// What the host chain knows of a frame when the reference runs prescanFrame on it.
struct CwdStep
{
    i32 begin, end;                 // stream lines the frame adds to the queue (the 112 closing lines of a file included)
    u32 nf_first; u16 nf_cnt, nf_hole;      // fillNextFieldForCWD (stc007datastitcher.cpp:5390-5437): the first lines of the next frame's
    u16 pad;                                // first field, appended for the pass and removed after it (nf_cnt = 0: none)
};
enum { CWD_QMAX = 112+2*ST_BUF_SIZE_FIELD+16+112+112, CWD_THREADS = 128, CWD_CLASSES = 16 };
enum { SDV_BF_CWD = 64 };           // sdv_block_rec.flags: isDataFixedByCWD
struct CwdParams
{
    StitchMap map; const CwdStep *steps; const int *chains;    // chains[2c] = first step, chains[2c+1] = number of steps
    DeintCfg cfg; long long n_blocks;
    sdv_block_rec *blocks; i16 *samples; u8 *sflags;
    u32 *broken_bits; u8 *broken_sum;           // candidate bits of the countdown walk (NULL: none kept)
    u8 *masked_bits;                            // host build only: seam-masked blocks (one byte per block)
    const CwdLine *carry_in; CwdLine *carry_out; int carry_out_step;    // patched lines handed from call to call (step -1: none)
    int *status;                                // [0] != 0: a queue did not fit (never with frames the stitcher can build)
    // Speculative walk of a long chain (every frame at once instead of one after the other): the chains listed are single frames;
    // a frame takes the lines its predecessor leaves in the queue from step_out[S-1] (step_mode[S] = 1) or, as a first guess, as they
    // are on the tape (0), notes what it took in step_used[S] and what it leaves in step_out[S] (112 CwdLines each, counts in
    // step_n[2S], step_n[2S+1]).  cwd_verify_kernel then compares step_used[S] with step_out[S-1]: by induction from the chain's first
    // frame, whose input is exact, every frame that passes is what the sequential walk produces; the others are walked again.
    CwdLine *step_out, *step_used; u16 *step_n; const u8 *step_mode; const u8 *step_first;     // NULL: sequential chains
};
struct CwdShared { CwdLine q[CWD_QMAX]; int fixes; int n_old; };

SDV_HD void cwd_bits_update(u32 *bits, u8 *sum, long long b, bool set)
{
    if(!bits) return;
    const u32 m = 1u<<(u32)(b&31);
#if defined(__CUDA_ARCH__)
    if(set) { atomicOr(&bits[b>>5], m); sum[b>>10] = 1; } else atomicAnd(&bits[b>>5], ~m);
#else
    if(set) { bits[b>>5] |= m; sum[b>>10] = 1; } else bits[b>>5] &= ~m;
#endif
}

// One chain of consecutive frames: per frame the queue (lines kept from the frame before, the frame, the preview of the next
// field), performCWD until nothing is repaired, performDeinterleave of the frame's blocks from the patched queue.
SDV_HD void cwd_chain_cta(const Cta &c, const CwdParams &p, int chain, CwdShared *sh)
{
    const StitchMap &m = p.map;
    const int s_first = p.chains[2*chain], s_cnt = p.chains[2*chain+1];
    int hint = 0;
    long long qa = 0;               // stream line of q[0]
    for(int S=s_first;S<s_first+s_cnt;S++)
    {
        const CwdStep st = p.steps[S];
        // ---- the queue
        int n_old;
        if(S==s_first)
        {   // lines the frames before left in the queue: as they are on the tape (those frames were clean), or as the previous call left them
            qa = ((long long)st.begin>112) ? ((long long)st.begin-112) : 0;
            n_old = (int)(st.begin-qa);
            if(p.step_out&&p.step_mode[S]&&(S>0))
            {   // speculative walk: what the frame before left, as far as it is known
                n_old = p.step_n[2*(S-1)+1];
                qa = (long long)st.begin-n_old;
                for(int i=c.tid;i<n_old;i+=c.n) sh->q[i] = p.step_out[(size_t)(S-1)*112+i];
            }
            else for(int i=c.tid;i<n_old;i+=c.n)
            {
                if((S==0)&&p.carry_in&&(qa+i<m.n_carry)) sh->q[i] = p.carry_in[qa+i];
                else { int h2 = (int)((qa+i-m.n_carry-m.lead)/m.frame_len); const AsmLine l = stitch_line(m, qa+i, &h2); sh->q[i] = cl_from_rec(l.rec, l.frame, l.line); }
            }
            if(p.step_out)
            {
                c.sync();
                for(int i=c.tid;i<n_old;i+=c.n) p.step_used[(size_t)S*112+i] = sh->q[i];
                if(c.tid==0) p.step_n[2*S] = (u16)n_old;
            }
        }
        else n_old = sh->n_old;
        c.sync();
        const int n_new = st.end-st.begin;
        const int n_q = n_old+n_new+st.nf_cnt;
        if(n_q>CWD_QMAX) { if(c.tid==0) p.status[0] = 1; return; }
        for(int i=c.tid;i<n_new;i+=c.n)
        {
            int h2 = (int)(((long long)st.begin+i-m.n_carry-m.lead)/m.frame_len);
            const AsmLine l = stitch_line(m, (long long)st.begin+i, &h2);
            sh->q[n_old+i] = cl_from_rec(l.rec, l.frame, l.line);
        }
        const i32 next_frame = m.frame_base+S+2;
        for(int i=c.tid;i<(int)st.nf_cnt;i+=c.n)
        {
            const sdv_line_rec *r = m.recs+st.nf_first+i+((i>=(int)st.nf_hole) ? 1 : 0);
            sh->q[n_old+n_new+i] = cl_from_rec(r, next_frame, 0);
        }
        c.sync();
        // ---- performCWD, pass after pass
        if(n_q>112)
        {
            DeintCfg pc = p.cfg; pc.force_check = pc.ignore_crc ? 0 : 1;
            if(m.step_res) pc.res_mode = seam_res_mode(stitch_line_res(m, S, sh->q[0].frame, sh->q[0].line), stitch_line_res(m, S, sh->q[112].frame, sh->q[112].line));
            if(pc.m2) pc.res_mode = SDV_RES_MODE_14BIT;
            for(int pass=0;pass<4096;pass++)
            {
                if(c.tid==0) sh->fixes = 0;
                c.sync();
                int mine = 0;
                for(int r=c.tid;r<CWD_CLASSES;r+=c.n)
                    for(int b=r;b<n_q-112;b+=CWD_CLASSES) mine += cwd_patch_block(sh->q, b, pc, next_frame);
#if defined(__CUDA_ARCH__)
                if(mine) atomicAdd(&sh->fixes, mine);
#else
                sh->fixes += mine;
#endif
                c.sync();
                const int total = sh->fixes;
                c.sync();
                if(total==0) break;
            }
        }
        // ---- performDeinterleave: the blocks that start in front of the last 112 lines of the frame
        const int n_keep = n_old+n_new;
        const int n_out = (n_keep>112) ? (n_keep-112) : 0;
        for(int i=c.tid;i<n_out;i+=c.n)
        {
            const long long b = qa+i;
            if(b>=p.n_blocks) continue;
            const CwdLine &first = sh->q[i], &last = sh->q[i+112];
            DeintCfg bc = p.cfg;
            if(m.step_res) bc.res_mode = seam_res_mode(stitch_line_res(m, S, first.frame, first.line), stitch_line_res(m, S, last.frame, last.line));
            if(bc.m2) bc.res_mode = SDV_RES_MODE_14BIT;
            bool masked = false;
            const int fi = last.frame-m.frame_base-1;
            if((fi>=0)&&(fi<m.n_frames))
            {
                const u8 mk = m.fa[fi].mask;
                if((mk&1)&&(first.frame==last.frame)&&(first.line>last.line)) masked = true;
                if((mk&2)&&(first.frame!=last.frame)&&(first.frame==last.frame-1)) masked = true;
            }
            BlockIn in; u8 cwd_lines;
            cwd_block_in(sh->q, i, bc.ignore_crc!=0, &in, &cwd_lines);
            BlockX x;
            deint_block_cwd(&x, &in, cwd_lines, bc, true);
            const bool by_cwd = blkx_fixed_by_cwd(&x);
            const bool silent = blk_silent(&x.b);
            bool unsafe = false;
            if(masked&&!silent) { unsafe = (x.b.audio_state!=SDV_AUD_BROKEN); blkx_mark_unsafe(&x); }
            const bool broken_ns = (x.b.audio_state==SDV_AUD_BROKEN)&&(!silent)&&(!masked);
            cwd_bits_update(p.broken_bits, p.broken_sum, b, broken_ns);
            if(p.masked_bits) p.masked_bits[b] = masked ? 1 : 0;
            if(p.samples&&p.sflags) blk_output(&x.b, p.samples+b*6, p.sflags+b*6);
            if(p.blocks) { blk_export(&x.b, unsafe, p.blocks+b); if(by_cwd&&!unsafe&&(x.b.audio_state!=SDV_AUD_BROKEN)) p.blocks[b].flags |= SDV_BF_CWD; }
        }
        c.sync();
        // ---- the last 112 lines stay for the next frame
        const int keep = (n_keep>112) ? 112 : n_keep;
        const int drop = n_keep-keep;
        if(drop>0)
        {
            for(int base=0;base<keep;base+=c.n)
            {   // (moved towards the front in chunks: a chunk is read by all threads before any of them writes)
                const int i = base+c.tid;
                CwdLine t;
                if(i<keep) t = sh->q[drop+i];
                c.sync();
                if(i<keep) sh->q[i] = t;
                c.sync();
            }
        }
        if(c.tid==0) sh->n_old = keep;
        qa += drop;
        c.sync();
        if((S==p.carry_out_step)&&p.carry_out) for(int i=c.tid;i<keep;i+=c.n) p.carry_out[i] = sh->q[i];
        if(p.step_out)
        {
            for(int i=c.tid;i<keep;i+=c.n) p.step_out[(size_t)S*112+i] = sh->q[i];
            if(c.tid==0) p.step_n[2*S+1] = (u16)keep;
        }
        (void)hint;
    }
}

}   // namespace sdv
