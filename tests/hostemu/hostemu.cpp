// tests/hostemu/hostemu.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Compiles the integer logic of the CUDA path (sdvpcmdecoder_b200/csrc/*.cuh, the SDV_HD functions) as host code and
// runs it with a "thread block" of one thread, so that the decode logic can be checked against the oracle on a
// machine without a GPU.  It is built by tests/conftest.py into tests/hostemu/_build/ and is never linked into, or
// loaded by, the product library: the product has no CPU path.
//
//   emu_v2d_chain  : every line through process_line_cta() + chain_line()        (what stc007_chain_kernel does on
//                    lines that fail the preset decode)
//   emu_v2d_hybrid : the host loop of sdv_bin_decode_frames() with scalar stand-ins for the bulk kernel and for the
//                    look-ahead of the chain kernel (checks the hand-off rules between the two)
//   emu_deint      : deint_block() + sample output + broken-block windows
#include <cstdio>
#include <cstdlib>
#define SDV_EMU_COUNTERS 1
static long long g_emu_counters[8];     // [0] bit-sliced PCM-1 searches, [1] grid points read one by one (PCM-1), [2]/[3] the same for PCM-16x0
#include <vector>
#include <cstring>
#include "../../sdvpcmdecoder_b200/csrc/stc007_chain.cuh"
#include "../../sdvpcmdecoder_b200/csrc/stc007_deint.cuh"

using namespace sdv;

static Work g_w;

// scalar stand-in of warp_fast_decode(): the (hysteresis 0, shift 0) candidate with the preset reference/coordinates
static bool fast_decode(const u8 *px, int W, const BinState *b, u16 *words9)
{
    Line l; line_clear(&l);
    l.ref = b->def_ref; l.black = 0; l.white = 255; l.coords = b->def_coord; l.ppb = make_ppb(b->def_coord);
    Cand cd; eval_cand(px, W-1, &l, 0, 0, &cd);
    for(int i=0;i<9;i++) words9[i] = cd.words[i];
    return cd.ok&&(cd.calc_crc==cd.words[8]);
}

extern "C" int emu_v2d_chain(int mode, int line_dup, const u8 *luma, int n_frames, int H, int W, sdv_line_rec *recs, sdv_line_aux *aux)
{
    static ChainCtx x;
    Cta c = { 0, 1 };
    Geom g = make_geom(W);
    chain_reset(&x, mode, line_dup);
    int hf = H/2;
    for(int f=0;f<n_frames;f++)
    {
        chain_frame_start(&x, f==0);
        for(int fld=0;fld<2;fld++)
        {
            for(int k=0;k<hf;k++)
            {
                BinState b = x.bin;
                process_line_cta(c, &g_w, &b, luma+((size_t)f*H+2*k+fld)*W, g);
                chain_line(&x, &g_w.o);
                size_t ridx = (size_t)f*H+(size_t)fld*hf+k;
                export_line(&g_w.o, recs+ridx, aux ? aux+ridx : 0);
            }
            chain_field_end(&x);
        }
        Coord mv, mi;
        { int scr = 0; median_cta(c, x.frame_valid, x.n_fv, &mv, &scr); }
        { int scr = 0; median_cta(c, x.frame_invalid, x.n_fi, &mi, &scr); }
        chain_frame_end(&x, mv, mi);
    }
    return 0;
}

// ---- scalar stand-in of stc007_bulk_kernel for one frame
static bool bulk_frame(const u8 *frame, int H, int W, const BinState *b, int line_dup, sdv_line_rec *recs, sdv_line_aux *aux)
{
    bool clean = true;
    int hf = H/2;
    for(int fld=0;fld<2;fld++)
    {
        u16 prev[8] = {0};
        for(int k=0;k<hf;k++)
        {
            u16 w[9];
            bool ok = fast_decode(frame+(size_t)(2*k+fld)*W, W, b, w);
            bool is_cb = ok&&words_control_block(w);
            if(!ok) clean = false;
            if(is_cb&&(k!=0)) clean = false;
            bool forced_bad = false;
            if((line_dup&1)&&!is_cb)
            {
                if(k==0) forced_bad = FINE_FIRST_LINE_DUP;
                else forced_bad = (words_diff8(w, prev)<=(BITS_PCM_DATA/32))&&!words_almost_silent(w, (line_dup&2)!=0);
            }
            Line l;
            line_from_fast(&l, b, w);
            if(!is_cb) { l.forced_bad = forced_bad; l.m2 = (u8)((line_dup>>1)&1); }
            size_t ridx = (size_t)fld*hf+k;
            export_line(&l, recs+ridx, aux ? aux+ridx : 0);
            if(!is_cb) memcpy(prev, w, 16);
        }
    }
    return clean;
}

extern "C" int emu_v2d_hybrid(int mode, int line_dup, const u8 *luma, int n_frames, int H, int W, sdv_line_rec *recs, sdv_line_aux *aux,
                              long long *stats /*[4]: chain lines, chain fast lines, bulk frames, bulk launches*/)
{
    static ChainCtx x;
    Cta c = { 0, 1 };
    Geom g = make_geom(W);
    chain_reset(&x, mode, line_dup);
    int hf = H/2;
    std::vector<u8> clean(n_frames+1, 0);
    bool have_spec = false; u8 spec_ref = 0, spec_black = 0, spec_white = 0; Coord spec_c = coord_none();
    int f = 0;
    stats[0] = stats[1] = stats[2] = stats[3] = 0;
    while(f<n_frames)
    {
        // ---- stc007_chain_kernel
        int nproc = 0, stable = 0;
        for(;;)
        {
            chain_frame_start(&x, f==0);
            const u8 *frame = luma+(size_t)f*H*W;
            for(int fld=0;fld<2;fld++)
            {
                int k = 0;
                while(k<hf)
                {
                    bool ready = bin_fast_ready(&x.bin);
                    bool slow = !ready;
                    if(ready)
                    {
                        int nb = (hf-k<32) ? (hf-k) : 32, taken = 0;
                        BinState b0 = x.bin;
                        FastRes fr[32]; FastPlan plan[32];
                        for(int i=0;i<nb;i++) fr[i].ok = fast_decode(frame+(size_t)(2*(k+i)+fld)*W, W, &b0, fr[i].words) ? 1 : 0;
                        size_t ridx = (size_t)f*H+(size_t)fld*hf+k;
                        int n = chain_fast_batch(c, &x, fr, nb, plan, &taken, recs+ridx, aux ? aux+ridx : 0);
                        stats[0] += n; stats[1] += n;
                        k += n; slow = n<nb;
                    }
                    if(slow&&(k<hf))
                    {
                        BinState b = x.bin;
                        process_line_cta(c, &g_w, &b, frame+(size_t)(2*k+fld)*W, g);
                        chain_line(&x, &g_w.o);
                        size_t ridx = (size_t)f*H+(size_t)fld*hf+k;
                        export_line(&g_w.o, recs+ridx, aux ? aux+ridx : 0);
                        stats[0]++;
                        k++;
                    }
                }
                chain_field_end(&x);
            }
            Coord mv, mi;
            { int scr = 0; median_cta(c, x.frame_valid, x.n_fv, &mv, &scr); }
            { int scr = 0; median_cta(c, x.frame_invalid, x.n_fi, &mi, &scr); }
            chain_frame_end(&x, mv, mi);
            int stop = ((f+1>=n_frames)||(nproc+1>=64)) ? 1 : 0, st = 0;
            if((f+1<n_frames)&&chain_is_stable(&x))
            {
                bool match = have_spec&&(spec_ref==x.bin.def_ref)&&coord_eq(spec_c, x.bin.def_coord);
                if(match) { if(clean[f+1]) { stop = 1; st = 1; } }
                else { stop = 1; st = 1; }
            }
            f++; nproc++;
            if(stop) { stable = st; break; }
        }
        if(f>=n_frames) break;
        if(!stable) continue;
        BinState b = x.bin;
        if(!have_spec||(spec_ref!=b.def_ref)||!coord_eq(spec_c, b.def_coord))
        {   // ---- stc007_bulk_kernel over all remaining frames
            for(int q=f;q<n_frames;q++)
                clean[q] = bulk_frame(luma+(size_t)q*H*W, H, W, &b, line_dup, recs+(size_t)q*H, aux ? aux+(size_t)q*H : 0) ? 1 : 0;
            have_spec = true; spec_ref = b.def_ref; spec_c = b.def_coord; spec_black = b.def_black; spec_white = b.def_white;
            stats[3]++;
        }
        int fb = f;
        while((fb<n_frames)&&clean[fb]) fb++;
        if(fb>f)
        {
            if((b.def_black!=spec_black)||(b.def_white!=spec_white))
                for(size_t i=(size_t)f*H;i<(size_t)fb*H;i++) if(recs[i].service_type==SDV_SRV_NO) { recs[i].black = b.def_black; recs[i].white = b.def_white; }
            chain_skip_clean_frames(&x, fb-f);
            stats[2] += fb-f;
            f = fb;
        }
    }
    return 0;
}

static int emu_deint_impl(const sdv_line_rec *lines, int n_lines, int res_mode, int ignore_crc, int force_check, int p_corr, int q_corr,
                          int broken_mask_dur, sdv_block_rec *blocks, i16 *samples, u8 *sflags, int countdown_in, int *countdown_out);
extern "C" int emu_deint(const sdv_line_rec *lines, int n_lines, int res_mode, int ignore_crc, int force_check, int p_corr, int q_corr,
                         int broken_mask_dur, sdv_block_rec *blocks, i16 *samples, u8 *sflags)
{
    return emu_deint_impl(lines, n_lines, res_mode, ignore_crc, force_check, p_corr, q_corr, broken_mask_dur, blocks, samples, sflags, 0, 0);
}
// the same with the countdown the blocks before this array left open (countdown_in) and what this array leaves (countdown_out)
extern "C" int emu_deint_carry(const sdv_line_rec *lines, int n_lines, int res_mode, int ignore_crc, int force_check, int p_corr, int q_corr,
                               int broken_mask_dur, sdv_block_rec *blocks, i16 *samples, u8 *sflags, int countdown_in, int *countdown_out)
{
    return emu_deint_impl(lines, n_lines, res_mode, ignore_crc, force_check, p_corr, q_corr, broken_mask_dur, blocks, samples, sflags, countdown_in, countdown_out);
}
static int emu_deint_impl(const sdv_line_rec *lines, int n_lines, int res_mode, int ignore_crc, int force_check, int p_corr, int q_corr,
                          int broken_mask_dur, sdv_block_rec *blocks, i16 *samples, u8 *sflags, int countdown_in, int *countdown_out)
{
    int nb = n_lines-112;
    if(nb<=0) return 0;
    DeintCfg cfg; cfg.m2 = (u8)((res_mode>>8)&1); cfg.res_mode = (u8)res_mode; cfg.ignore_crc = (u8)ignore_crc; cfg.force_check = (u8)force_check; cfg.p_corr = (u8)p_corr; cfg.q_corr = (u8)q_corr;
    std::vector<u8> unsafe(nb, 0);
    for(int pass=0;pass<2;pass++)
    {
        long long open_until = (pass==0) ? countdown_in : -1;
        bool any = false;
        if((pass==0)&&(countdown_in>0)) { any = true; for(int q=0;(q<countdown_in)&&(q<nb);q++) unsafe[q] = 1; }
        for(int b=0;b<nb;b++)
        {
            BlockIn in; in.ok = 0;
            for(int k=0;k<8;k++)
            {
                const sdv_line_rec *r = lines+b+16*k;
                in.w[k] = r->words[k]; in.sw[k] = r->words[7];
                if(line_rec_ok(r, ignore_crc!=0)) in.ok |= (u8)(1<<k);
            }
            Block blk;
            deint_dispatch(&blk, &in, cfg);
            bool silent = blk_silent(&blk);
            bool broken_ns = (blk.audio_state==SDV_AUD_BROKEN)&&!silent;
            if(pass==0)
            {
                if(broken_ns&&(broken_mask_dur>0)&&(b>=open_until)) { open_until = (long long)b+broken_mask_dur; any = true; for(int q=b;(q<b+broken_mask_dur)&&(q<nb);q++) unsafe[q] = 1; }
            }
            bool uns = false;
            if((pass==1)&&unsafe[b]&&!silent) { uns = blk.audio_state!=SDV_AUD_BROKEN; blk_mark_unsafe(&blk); }
            if(samples&&sflags) blk_output(&blk, samples+(size_t)b*6, sflags+(size_t)b*6);
            if(blocks) blk_export(&blk, uns, blocks+b);
        }
        if((pass==0)&&countdown_out) *countdown_out = (int)((open_until>nb) ? (open_until-nb) : 0);
        if(!any) break;
    }
    return nb;
}

extern "C" int emu_sizes(int *out) { out[0] = (int)sizeof(sdv_line_rec); out[1] = (int)sizeof(sdv_line_aux); out[2] = (int)sizeof(sdv_block_rec); out[3] = (int)sizeof(ChainCtx); return 0; }

// ---- PCM-16x0 deinterleave: the same x0_process_block()/x0_output() the kernel runs, one data block at a time
#include "../../sdvpcmdecoder_b200/csrc/pcm16x0_deint.cuh"
extern "C" int emu_deint_pcm16x0(const sdv_pcm16x0_subline *sub, int n_itl, int ignore_crc, int force_check, int p_corr, int ei,
                                 i16 *samples, u8 *sflags, u8 *states)
{
    X0Cfg cfg; cfg.ignore_crc = (u8)ignore_crc; cfg.force_check = (u8)force_check; cfg.p_corr = (u8)p_corr;
    const int per = ei ? X0_BLOCKS_EI : X0_BLOCKS_ITL, unit = ei ? X0_SUBLINES_EI : X0_SUBLINES_ITL, ofs = ei ? X0_OFS_EI : X0_OFS;
    long long nb = (long long)n_itl*per;
    for(long long b=0;b<nb;b++)
    {
        long long m = b/per; int i = (int)(b-m*per);
        const sdv_pcm16x0_subline *base = sub+m*unit+i;
        X0Block blk;
        x0_process_block(&blk, base, base+ofs, base+2*ofs, (i&1)!=0, cfg);
        x0_output(&blk, samples+b*6, sflags+b*6, states+b*3);
    }
    return (int)nb;
}

// ---- PCM-1 line decode + chain: prescan, frame presets, every line through p1_process_line_cta() + p1_chain_line()
#include "../../sdvpcmdecoder_b200/csrc/pcm1_chain.cuh"
static P1Work g_p1w;
extern "C" int emu_p1_v2d_chain(int mode, int line_dup, const u8 *luma, int n_frames, int H, int W, sdv_line_rec *recs, sdv_line_aux *aux,
                                P1Preset *presets /*[n_frames], may be NULL*/)
{
    static P1ChainCtx x;
    Cta c = { 0, 1 };
    Geom g = make_geom(W);
    p1_chain_reset(&x, mode, line_dup);
    int hf = H/2;
    for(int f=0;f<n_frames;f++)
    {
        const bool first = (f==0);
        const bool ran = p1_prescan_runs(H, first, mode);
        P1Preset ps; ps.valid = 0; ps.ref = 0; ps.coords = coord_none(); ps.pad[0] = ps.pad[1] = 0;
        if(ran)
        {
            P1Preset r[P1_COORD_CHECK_LINES];
            for(int idx=0;idx<P1_COORD_CHECK_LINES;idx++)
            {
                r[idx] = ps;
                int row = p1_prescan_row(H, first, idx);
                if(row<0) continue;
                BinState b = x.bin; bin_reset_good(&b);
                p1_process_line_cta(c, &g_p1w, &b, true, luma+((size_t)f*H+row)*W, g);
                if(p1_crc_ok(&g_p1w.o)) { r[idx].valid = 1; r[idx].coords = g_p1w.o.coords; r[idx].ref = g_p1w.o.ref; }
            }
            ps = p1_prescan_reduce(r);
        }
        if(presets) presets[f] = ps;
        p1_chain_frame_start(&x, ran, ps);
        for(int fld=0;fld<2;fld++)
        {
            for(int k=0;k<hf;k++)
            {
                BinState b = x.bin;
                p1_process_line_cta(c, &g_p1w, &b, p1_chain_coord_search(&x), luma+((size_t)f*H+2*k+fld)*W, g);
                p1_chain_line(&x, &g_p1w.o);
                size_t ridx = (size_t)f*H+(size_t)fld*hf+k;
                p1_export_line(&g_p1w.o, recs+ridx, aux ? aux+ridx : 0);
            }
            p1_chain_field_end(&x);
        }
        p1_chain_frame_end(&x, median_small(x.frame_valid, x.n_fv), median_small(x.frame_invalid, x.n_fi));
    }
    return 0;
}

// ---- PCM-1 frame assembly (PCM1DataStitcher): p1_assemble_frame_cta() per frame
#include "../../sdvpcmdecoder_b200/csrc/pcm1_stitch.cuh"
extern "C" int emu_p1_assemble(const sdv_line_rec *recs, int n_frames, int H, int bff, int file_start, int manual, int ofs_odd, int ofs_even,
                               sdv_pcm1_subline *sub, sdv_pcm1_frame_info *info)
{
    static P1AsmScratch s;
    Cta c = { 0, 1 };
    for(int f=0;f<n_frames;f++)
        p1_assemble_frame_cta(c, recs+(size_t)f*H, H, bff!=0, (file_start!=0)&&(f==0), manual!=0, ofs_odd, ofs_even, sub+(size_t)f*2*P1S_SUBLINES_PF, &s, info ? info+f : 0);
    return 0;
}

// ---- PCM-16x0 line decode + chain: prescan (right part), frame presets, three sub-lines per video line
#include "../../sdvpcmdecoder_b200/csrc/pcm16x0_chain.cuh"
static X0Work g_x0w;
extern "C" int emu_x0_v2d_chain(int mode, int line_dup, const u8 *luma, int n_frames, int H, int W, sdv_line_rec *recs, sdv_line_aux *aux,
                                P1Preset *presets /*[n_frames], may be NULL*/)
{
    static X0ChainCtx x;
    Cta c = { 0, 1 };
    Geom g = make_geom(W);
    x0_chain_reset(&x, mode, line_dup);
    int hf = H/2;
    std::vector<u8> scanned(H);
    for(int f=0;f<n_frames;f++)
    {
        const bool first = (f==0);
        const bool ran = p1_prescan_runs(H, first, mode);
        P1Preset ps; ps.valid = 0; ps.ref = 0; ps.coords = coord_none(); ps.pad[0] = ps.pad[1] = 0;
        std::fill(scanned.begin(), scanned.end(), 0);
        if(ran)
        {
            P1Preset r[P1_COORD_CHECK_LINES];
            for(int idx=0;idx<P1_COORD_CHECK_LINES;idx++)
            {
                r[idx] = ps;
                int row = p1_prescan_row(H, first, idx);
                if(row<0) continue;
                BinState b = x.bin; bin_reset_good(&b);
                g_x0w.scan_done = 0;
                x0_process_line_cta(c, &g_x0w, &b, X0L_RIGHT, true, luma+((size_t)f*H+row)*W, g);
                scanned[row] = g_x0w.scan_done;
                if(x0_crc_ok(&g_x0w.o)) { r[idx].valid = 1; r[idx].coords = g_x0w.o.coords; r[idx].ref = g_x0w.o.ref; }
            }
            ps = p1_prescan_reduce(r);
        }
        if(presets) presets[f] = ps;
        x0_chain_frame_start(&x, ran, ps);
        for(int fld=0;fld<2;fld++)
        {
            for(int k=0;k<hf;k++)
            {
                const int row = 2*k+fld;
                g_x0w.scan_done = scanned[row];
                x0_chain_line_start(&x);
                for(int part=0;part<3;part++)
                {
                    BinState b = x.bin;
                    x0_process_line_cta(c, &g_x0w, &b, part, x0_chain_coord_search(&x), luma+((size_t)f*H+row)*W, g);
                    x0_chain_subline(&x, &g_x0w.o, g_x0w.scan_done!=0);
                    size_t ridx = ((size_t)f*H+(size_t)fld*hf+k)*3+part;
                    x0_export_line(&g_x0w.o, recs+ridx, aux ? aux+ridx : 0);
                }
            }
            x0_chain_field_end(&x);
        }
        x0_chain_frame_end(&x, median_small(x.frame_valid, x.n_fv), median_small(x.frame_invalid, x.n_fi));
    }
    return 0;
}

// ---- PCM-16x0 frame assembly + deinterleave with preset alignment (PCM16X0DataStitcher, SI format)
#include "../../sdvpcmdecoder_b200/csrc/pcm16x0_stitch.cuh"
extern "C" int emu_x0_stitch_info(const sdv_line_rec *recs, int n_frames, int H, int bff, int top_odd, int top_even, int ignore_crc, int p_corr,
                                  int broken_mask_dur, const u8 *mask_seams, i16 *samples, u8 *sflags, sdv_pcm16x0_frame_info *info)
{
    static X0AsmScratch s;
    Cta c = { 0, 1 };
    X0Cfg cfg; cfg.ignore_crc = (u8)ignore_crc; cfg.force_check = (u8)!ignore_crc; cfg.p_corr = (u8)p_corr;
    for(int f=0;f<n_frames;f++)
        x0_stitch_frame_cta(c, recs+(size_t)f*H*3, H, bff!=0, top_odd, top_even, cfg, broken_mask_dur, mask_seams ? (mask_seams[f]!=0) : false, &s,
                            samples+(size_t)f*X0S_BLOCKS_FRAME*6, sflags+(size_t)f*X0S_BLOCKS_FRAME*6, info+f);
    for(int f=0;f<n_frames;f++) x0_ctrl_effective(info, f);
    return 0;
}
extern "C" int emu_x0_stitch(const sdv_line_rec *recs, int n_frames, int H, int bff, int top_odd, int top_even, int ignore_crc, int p_corr,
                             int broken_mask_dur, const u8 *mask_seams, i16 *samples, u8 *sflags)
{
    static X0AsmScratch s;
    Cta c = { 0, 1 };
    X0Cfg cfg; cfg.ignore_crc = (u8)ignore_crc; cfg.force_check = (u8)!ignore_crc; cfg.p_corr = (u8)p_corr;
    for(int f=0;f<n_frames;f++)
        x0_stitch_frame_cta(c, recs+(size_t)f*H*3, H, bff!=0, top_odd, top_even, cfg, broken_mask_dur, mask_seams ? (mask_seams[f]!=0) : false, &s,
                            samples+(size_t)f*X0S_BLOCKS_FRAME*6, sflags+(size_t)f*X0S_BLOCKS_FRAME*6);
    return 0;
}

// ---- STC-007 stitcher: trims, seam statistics and the assembled-stream deinterleave through the same SDV_HD code the
// kernels run (stc007_stitch.cuh), the decision chain through the library's own host code (stc007_stitch_host.h).
#include "../../sdvpcmdecoder_b200/csrc/stc007_stitch_host.h"
namespace {
struct HostSeams : SeamOracle
{
    const sdv_line_rec *recs; const FrameTrim *trims; int n_frames, H; DeintCfg cfg; int lim14, lim16;
    const u8 *step_res;
    std::vector<sdv_stitch_stats> tmp;
    long long evaluations;
    SeamField field(int frame, int even) const
    {
        SeamField f; f.first = 0; f.size = 0; f.hole = ST_NO_HOLE;
        if(frame>=n_frames) return f;
        const FieldTrim &t = even ? trims[frame].even : trims[frame].odd;
        f.first = (u32)((size_t)frame*H+(even ? H/2 : 0)+t.first); f.size = t.data_lines; f.hole = t.hole;
        return f;
    }
    sdv_stitch_stats eval(int frame, int kind, int pad)
    {
        static const int tab[SEAM_KINDS][3] = { {0, 0, 1}, {1, 0, 0}, {1, 1, 0}, {0, 1, 1}, {1, 1, 1}, {0, 1, 0} };
        SeamTask t; t.f1 = field(frame, tab[kind][0]); t.f2 = field(frame+tab[kind][1], tab[kind][2]); t.pad0 = (u16)pad; t.n_pad = 1; t.out = 0;
        t.res1 = t.res2 = RES_ANY;
        if(step_res) { const u8 *r = step_res+4*(size_t)frame; t.res1 = r[tab[kind][0]]; t.res2 = r[2*tab[kind][1]+tab[kind][2]]; }
        const SeamGeom g = seam_geom(t.f1.size, t.f2.size, pad);
        const int lim = cfg.q_corr ? lim14 : lim16;
        DeintCfg qc = cfg; qc.res_mode = seam_queue_res_mode(t, g, cfg.res_mode);
        SeamCount c; seam_count_init(&c);
        for(int s=0;s<g.nblk;s++) seam_count_step(&c, seam_block_flags(recs, t, g, s, qc), lim);
        evaluations++;
        return seam_count_finish(&c, g, lim);
    }
    bool try_padding(int frame, int kind, int padding, uint8_t *result) override { *result = eval(frame, kind, padding).result; return true; }
    bool sweep(int frame, int kind, const sdv_stitch_stats **stats32) override
    {
        tmp.resize(32);
        for(int p=0;p<32;p++) tmp[p] = eval(frame, kind, p);
        *stats32 = tmp.data();
        return true;
    }
};
}
// settings: [video_std, field_order, res16, mask_seams, fix_cut_above, max_unch14, max_unch16, file_start, file_end, cwd]
extern "C" int emu_stc007_stitch(const sdv_line_rec *recs, int n_frames, int H, const int *settings, int res_mode, int ignore_crc, int p_corr, int q_corr,
                                 int broken_mask_dur, int m2, sdv_block_rec *blocks, i16 *samples, u8 *sflags, sdv_stc007_frame_info *info)
{
    Cta c = { 0, 1 };
    std::vector<FrameTrim> trims((size_t)n_frames+1);
    memset(&trims[n_frames], 0, sizeof(FrameTrim)); trims[n_frames].odd.hole = trims[n_frames].even.hole = ST_NO_HOLE;
    int scr[8];
    for(int f=0;f<n_frames;f++)
    {
        trim_field_cta(c, recs+(size_t)f*H, H/2, 0, scr, &trims[f].odd);
        trim_field_cta(c, recs+(size_t)f*H+H/2, H/2, 1, scr, &trims[f].even);
    }
    DeintCfg cfg; cfg.m2 = (u8)m2; cfg.res_mode = (u8)res_mode; cfg.ignore_crc = (u8)ignore_crc; cfg.force_check = (u8)!ignore_crc;
    cfg.q_corr = (u8)(q_corr ? 1 : 0); cfg.p_corr = (u8)((p_corr||q_corr) ? 1 : 0);
    HostSeams seams; seams.recs = recs; seams.trims = trims.data(); seams.n_frames = n_frames; seams.H = H; seams.cfg = cfg; seams.cfg.force_check = 1;
    seams.lim14 = settings[5]; seams.lim16 = settings[6]; seams.evaluations = 0;
    Stitcher sx;
    sx.set.video_std = (u8)settings[0]; sx.set.field_order = (u8)settings[1]; sx.set.res16 = (u8)settings[2];
    sx.set.mask_seams = (u8)settings[3]; sx.set.fix_cut_above = (u8)settings[4]; sx.set.max_unch14 = (u8)settings[5]; sx.set.max_unch16 = (u8)settings[6];
    sx.set.p_corr = cfg.p_corr; sx.set.q_corr = cfg.q_corr;
    sx.st.reset(); sx.seams = &seams;
    const int file_end = settings[8];
    const int n_done = file_end ? n_frames : ((n_frames>0) ? n_frames-1 : 0);
    const bool res_auto = (settings[2]==2)&&!m2;
    std::vector<u8> step_res(4*(size_t)n_done+4);
    if(res_auto)
    {   // getFieldResolution of every field, then detectAudioResolution frame by frame
        std::vector<u8> fr(2*(size_t)n_frames+2, (u8)ST_RES_UNKNOWN);
        for(int f=0;f<n_frames;f++) for(int even=0;even<2;even++)
        {
            const SeamField fld = seams.field(f, even);
            const int n = (fld.size>112) ? (fld.size-112) : 0;
            int c14 = 0, c16 = 0;
            for(int i=0;i<n;i++) field_res_step(&c14, &c16, field_res_flags(recs, fld, i, false));
            fr[2*(size_t)f+even] = n ? field_res_decide(c14, c16) : (u8)ST_RES_UNKNOWN;
        }
        ResChain rc; rc.reset();
        for(int f=0;f<n_done;f++) rc.step(fr[2*(size_t)f], fr[2*(size_t)f+1], fr[2*(size_t)f+2], fr[2*(size_t)f+3], &step_res[4*(size_t)f]);
    }
    seams.step_res = res_auto ? step_res.data() : 0;
    std::vector<FrameAsm> fa((size_t)n_done+1);
    const bool cwd = settings[9]!=0;
    std::vector<CwdStep> cwd_steps((size_t)n_done);
    long long pos = ST_LEAD_IN; int frame_len = 2*ST_LINES_PF_NTSC, lead_line0 = 0;
    for(int f=0;f<n_done;f++)
    {
        if(!sx.step(f, trims[f], trims[f+1], &fa[f], res_auto ? &step_res[4*(size_t)f] : 0)) return -1;
        fa[f].start = (i32)pos; pos += fa[f].total;
        const FrameSt &r = sx.st.f0;
        if(cwd)
        {
            CwdStep cs; cs.begin = fa[f].start; cs.end = cs.begin+fa[f].total+(((f==n_done-1)&&file_end) ? ST_TAIL : 0);
            cwd_next_field(r, trims[f+1], (size_t)(f+1)*H, H, &cs);
            cwd_steps[f] = cs;
        }
        if(f==0) { const int T = (r.video_std==ST_VID_PAL) ? ST_LINES_PF_PAL : ST_LINES_PF_NTSC; frame_len = 2*T; lead_line0 = 2*T-2*ST_LEAD_IN; }
        if(info)
        {
            sdv_stc007_frame_info o; memset(&o, 0, sizeof(o));
            o.start = fa[f].start; o.pre = fa[f].pre; o.n1 = fa[f].n1; o.inner = fa[f].inner; o.n2 = fa[f].n2; o.outer = fa[f].outer;
            o.skip1 = fa[f].skip1; o.skip2 = fa[f].skip2;
            o.odd_top = trims[f].odd.top; o.odd_bottom = trims[f].odd.bottom; o.even_top = trims[f].even.top; o.even_bottom = trims[f].even.bottom;
            o.odd_data_lines = trims[f].odd.data_lines; o.even_data_lines = trims[f].even.data_lines;
            o.odd_valid_lines = trims[f].odd.valid_lines; o.even_valid_lines = trims[f].even.valid_lines;
            o.inner_padding = r.inner_pad; o.outer_padding = r.outer_pad; o.field_order = r.order; o.video_std = r.video_std;
            o.flags = (uint8_t)((r.inner_ok ? SDV_FA_INNER_OK : 0)|(r.outer_ok ? SDV_FA_OUTER_OK : 0)|(r.inner_silence ? SDV_FA_INNER_SILENCE : 0)
                      |(r.outer_silence ? SDV_FA_OUTER_SILENCE : 0)|(r.order_guessed ? SDV_FA_ORDER_GUESSED : 0)
                      |((fa[f].mask&1) ? SDV_FA_MASK_INNER : 0)|((fa[f].mask&2) ? SDV_FA_MASK_PREV_OUTER : 0));
            o.odd_res_mode = res_auto ? r.odd_res : (u8)res_mode; o.even_res_mode = res_auto ? r.even_res : (u8)res_mode;
            info[f] = o;
        }
    }
    StitchMap m; memset(&m, 0, sizeof(m));
    if(res_auto&&(n_done>0)) { m.step_res = step_res.data(); m.f0_res[0] = m.f0_res[1] = step_res[fa[0].first_even ? 1 : 0]; }
    m.recs = recs; m.fa = fa.data(); m.n_frames = n_done; m.H = H; m.lead = ST_LEAD_IN; m.lead_line0 = lead_line0; m.tail = file_end ? ST_TAIL : 0;
    m.frame_base = 0; m.frame_len = frame_len; m.n_lines = pos+m.tail;
    const long long nb = (m.n_lines>ST_TAIL) ? (m.n_lines-ST_TAIL) : 0;
    int countdown = 0, hint = 0;
    if(cwd)
    {   // as the library does it: every block without CWD, the chains of frames CWD can touch once more from their patched queues,
        // then the countdown over the stored blocks
        std::vector<sdv_block_rec> own; if(!blocks) { own.resize((size_t)nb+1); blocks = own.data(); }
        std::vector<u8> masked_v((size_t)nb+1, 0);
        for(long long b=0;b<nb;b++)
        {
            BlockIn in; DeintCfg bc = cfg;
            const bool masked = stitch_block_in(m, b, ignore_crc!=0, &in, &hint, &bc.res_mode);
            Block blk; deint_dispatch(&blk, &in, bc);
            bool unsafe = false;
            if(masked&&!blk_silent(&blk)) { unsafe = blk.audio_state!=SDV_AUD_BROKEN; blk_mark_unsafe(&blk); }
            masked_v[b] = masked ? 1 : 0;
            blk_export(&blk, unsafe, blocks+b);
        }
        std::vector<u8> patch((size_t)n_frames+1, 0);
        for(int f=0;f<n_frames;f++) for(int j=0;j<H;j++) if(rec_cwd_patchable(recs+(size_t)f*H+j)) { patch[f] = 1; break; }
        std::vector<int> chains; std::vector<u8> dirty;
        cwd_plan_chains(patch.data(), fa.data(), n_done, false, &chains, &dirty);
        int status = 0;
        CwdParams cp; memset(&cp, 0, sizeof(cp));
        cp.map = m; cp.steps = cwd_steps.data(); cp.chains = chains.data(); cp.cfg = cfg; cp.n_blocks = nb;
        cp.blocks = blocks; cp.masked_bits = masked_v.data(); cp.carry_out_step = -1; cp.status = &status;
        static CwdShared sh;
        for(size_t ch=0;ch<chains.size()/2;ch++) cwd_chain_cta(c, cp, (int)ch, &sh);
        if(status) return -2;
        for(long long b=0;b<nb;b++)
        {
            const sdv_block_rec r = blocks[b];
            Block blk; for(int k=0;k<8;k++) blk.words[k] = r.words[k];
            blk.line_crc = r.line_crc; blk.word_valid = r.word_valid; blk.resolution = r.resolution; blk.audio_state = r.audio_state; blk.m2 = cfg.m2;
            bool unsafe = (r.flags&SDV_BF_UNSAFE)!=0;
            if(!(r.flags&SDV_BF_SILENT)&&!masked_v[b])
            {
                if((broken_mask_dur>0)&&(countdown==0)&&(blk.audio_state==SDV_AUD_BROKEN)) countdown = broken_mask_dur;
                if(countdown!=0) { unsafe = blk.audio_state!=SDV_AUD_BROKEN; blk_mark_unsafe(&blk); }
            }
            if(countdown>0) countdown--;
            if(samples&&sflags) blk_output(&blk, samples+(size_t)b*6, sflags+(size_t)b*6);
            const u8 keep = (u8)(r.flags&SDV_BF_CWD);
            blk_export(&blk, unsafe, blocks+b);
            if(!unsafe) blocks[b].flags |= keep;
        }
        return (int)nb;
    }
    for(long long b=0;b<nb;b++)
    {   // performDeinterleave, block by block as the reference does it
        BlockIn in;
        DeintCfg bc = cfg;
        const bool masked = stitch_block_in(m, b, ignore_crc!=0, &in, &hint, &bc.res_mode);
        Block blk;
        deint_dispatch(&blk, &in, bc);
        bool unsafe = false;
        if(!blk_silent(&blk))
        {
            if(masked) { unsafe = blk.audio_state!=SDV_AUD_BROKEN; blk_mark_unsafe(&blk); }
            else
            {
                if((broken_mask_dur>0)&&(countdown==0)&&(blk.audio_state==SDV_AUD_BROKEN)) countdown = broken_mask_dur;
                if(countdown!=0) { unsafe = blk.audio_state!=SDV_AUD_BROKEN; blk_mark_unsafe(&blk); }
            }
        }
        if(countdown>0) countdown--;
        if(samples&&sflags) blk_output(&blk, samples+(size_t)b*6, sflags+(size_t)b*6);
        if(blocks) blk_export(&blk, unsafe, blocks+b);
    }
    return (int)nb;
}

// ---- PCM-16x0 (SI) with the reference's own vertical alignment: the padding scan per field (x0_sipad_scan_cta), the library's
// decision chain (pcm16x0_stitch_host.h), the frame stitcher with the alignment found
#include "../../sdvpcmdecoder_b200/csrc/pcm16x0_stitch_host.h"
extern "C" int emu_x0_stitch_auto(const sdv_line_rec *recs, int n_frames, int H, int bff, int ignore_crc, int p_corr, int broken_mask_dur,
                                  int mask_seams, i16 *samples, u8 *sflags, sdv_pcm16x0_alignment *align)
{
    static X0AsmScratch s; static X0PadScratch ps;
    Cta c = { 0, 1 };
    X0Cfg cfg; cfg.ignore_crc = (u8)ignore_crc; cfg.force_check = (u8)!ignore_crc; cfg.p_corr = (u8)p_corr;
    X0PadChain chain; chain.reset(); chain.p_corr = p_corr!=0;
    for(int f=0;f<n_frames;f++)
    {
        X0PadScan sc[2]; X0FieldGeo geo[2]; uint8_t res[2];
        for(int k=0;k<2;k++) x0_sipad_scan_cta(c, recs+(size_t)f*H*3, H, k, cfg, &ps, &sc[k]);
        const bool m = chain.frame(sc[0], sc[1], geo, res);
        if(align)
        {
            sdv_pcm16x0_alignment a; memset(&a, 0, sizeof(a));
            for(int k=0;k<2;k++) { a.top_padding[k] = geo[k].top_pad; a.cut_lines[k] = geo[k].cut; a.lines[k] = geo[k].lines; a.result[k] = res[k]; }
            a.mask_seams = (m&&mask_seams) ? 1 : 0;
            align[f] = a;
        }
        x0_stitch_frame_cta(c, recs+(size_t)f*H*3, H, bff!=0, 0, 0, cfg, broken_mask_dur, m&&mask_seams, &s,
                            samples+(size_t)f*X0S_BLOCKS_FRAME*6, sflags+(size_t)f*X0S_BLOCKS_FRAME*6, 0, geo);
    }
    return 0;
}

// ---- PCM-16x0 (EI): the padding scan per frame (x0_eipad_scan_cta), findEIFrameStitching's decisions (X0PadChain::frame_ei),
// the frame stitcher in EI order
extern "C" int emu_x0_stitch_auto_ei(const sdv_line_rec *recs, int n_frames, int H, int bff, int ignore_crc, int p_corr, int broken_mask_dur,
                                     int mask_seams, i16 *samples, u8 *sflags, sdv_pcm16x0_alignment *align)
{
    static X0AsmScratch s; static X0EIScratch es;
    Cta c = { 0, 1 };
    X0Cfg cfg; cfg.ignore_crc = (u8)ignore_crc; cfg.force_check = (u8)!ignore_crc; cfg.p_corr = (u8)p_corr;
    X0PadChain chain; chain.reset(); chain.p_corr = p_corr!=0;
    for(int f=0;f<n_frames;f++)
    {
        X0EIScan sc; X0FieldGeo geo[2]; uint8_t res[2];
        x0_eipad_scan_cta(c, recs+(size_t)f*H*3, H, bff!=0, cfg, &es, &sc);
        const bool m = chain.frame_ei(sc, bff!=0, geo, res);
        if(align)
        {
            sdv_pcm16x0_alignment a; memset(&a, 0, sizeof(a));
            for(int k=0;k<2;k++) { a.top_padding[k] = geo[k].top_pad; a.cut_lines[k] = geo[k].cut; a.lines[k] = geo[k].lines; a.result[k] = res[k]; }
            a.mask_seams = (m&&mask_seams) ? 1 : 0;
            align[f] = a;
        }
        x0_stitch_frame_cta(c, recs+(size_t)f*H*3, H, bff!=0, 0, 0, cfg, broken_mask_dur, m&&mask_seams, &s,
                            samples+(size_t)f*X0S_BLOCKS_FRAME*6, sflags+(size_t)f*X0S_BLOCKS_FRAME*6, 0, geo, true);
    }
    return 0;
}

// The EI scan of one frame (diagnostics / tests): tryEIPadding statistics [81][6] + {zero_ofs, iblk, n_sub, top} per field
extern "C" int emu_x0_ei_scan(const sdv_line_rec *fr, int H, int bff, int ignore_crc, int p_corr, int *stats, int *fields)
{
    static X0EIScratch es;
    Cta c = { 0, 1 };
    X0Cfg cfg; cfg.ignore_crc = (u8)ignore_crc; cfg.force_check = (u8)!ignore_crc; cfg.p_corr = (u8)p_corr;
    X0EIScan sc;
    x0_eipad_scan_cta(c, fr, H, bff!=0, cfg, &es, &sc);
    for(int p=0;p<X0S_MAX_PAD_EI;p++)
    {
        const sdv_stitch_stats &t = sc.st[p];
        int *o = stats+6*p;
        o[0] = t.index; o[1] = t.valid; o[2] = t.silent; o[3] = t.unchecked; o[4] = t.broken; o[5] = t.result;
    }
    for(int k=0;k<2;k++) { fields[4*k] = sc.zero_ofs[k]; fields[4*k+1] = sc.iblk[k]; fields[4*k+2] = sc.n_sub[k]; fields[4*k+3] = sc.top[k]; }
    return 0;
}

// ---- Binarizer fine settings of the host build (the numeric fields of bin_preset_t): v == NULL restores the defaults
extern "C" void emu_set_fine(const int *v)
{
    const FineSet d = SDV_FINE_DEFAULTS;
    h_fine = d;
    if(!v) return;
    h_fine.max_black_lvl = (u8)v[0]; h_fine.min_white_lvl = (u8)v[1]; h_fine.min_contrast = (u8)v[2]; h_fine.min_ref_lvl = (u8)v[3];
    h_fine.max_ref_lvl = (u8)v[4]; h_fine.min_valid_crcs = (u8)v[5]; h_fine.mark_max_dist = (u8)v[6]; h_fine.left_bit_pick = (u8)v[7];
    h_fine.right_bit_pick = (u8)v[8]; h_fine.en_coord_search = (u8)(v[9] ? 1 : 0); h_fine.en_first_line_dup = (u8)(v[10] ? 1 : 0);
}

extern "C" void emu_counters(long long *out, int reset) { for(int i=0;i<8;i++) { out[i] = g_emu_counters[i]; if(reset) g_emu_counters[i] = 0; } }

// ---- diagnostic for the speculative CWD walk: for every frame of one chain, does the walk from the raw lines leave the same lines
// in the queue as the exact sequential walk?  out[S] = 1 same, 0 different, 2 frame not dirty.  (Not a test of the product.)
extern "C" int emu_cwd_speculation_probe(const sdv_line_rec *recs, int n_frames, int H, u8 *out)
{
    Cta c = { 0, 1 };
    std::vector<FrameTrim> trims((size_t)n_frames+1);
    memset(&trims[n_frames], 0, sizeof(FrameTrim)); trims[n_frames].odd.hole = trims[n_frames].even.hole = ST_NO_HOLE;
    int scr[8];
    for(int f=0;f<n_frames;f++) { trim_field_cta(c, recs+(size_t)f*H, H/2, 0, scr, &trims[f].odd); trim_field_cta(c, recs+(size_t)f*H+H/2, H/2, 1, scr, &trims[f].even); }
    DeintCfg cfg; cfg.m2 = 0; cfg.res_mode = 0; cfg.ignore_crc = 0; cfg.force_check = 1; cfg.q_corr = 1; cfg.p_corr = 1;
    HostSeams seams; seams.recs = recs; seams.trims = trims.data(); seams.n_frames = n_frames; seams.H = H; seams.cfg = cfg; seams.lim14 = 0x40; seams.lim16 = 0x20; seams.evaluations = 0; seams.step_res = 0;
    Stitcher sx; sx.set.video_std = 1; sx.set.field_order = 1; sx.set.res16 = 0; sx.set.mask_seams = 1; sx.set.fix_cut_above = 0; sx.set.max_unch14 = 0x40; sx.set.max_unch16 = 0x20; sx.set.p_corr = 1; sx.set.q_corr = 1;
    sx.st.reset(); sx.seams = &seams;
    std::vector<FrameAsm> fa((size_t)n_frames+1); std::vector<CwdStep> steps((size_t)n_frames);
    long long pos = ST_LEAD_IN;
    for(int f=0;f<n_frames;f++)
    {
        if(!sx.step(f, trims[f], trims[f+1], &fa[f], 0)) return -1;
        fa[f].start = (i32)pos; pos += fa[f].total;
        CwdStep cs; cs.begin = fa[f].start; cs.end = cs.begin+fa[f].total+((f==n_frames-1) ? ST_TAIL : 0);
        cwd_next_field(sx.st.f0, trims[f+1], (size_t)(f+1)*H, H, &cs);
        steps[f] = cs;
    }
    StitchMap m; memset(&m, 0, sizeof(m));
    m.recs = recs; m.fa = fa.data(); m.n_frames = n_frames; m.H = H; m.lead = ST_LEAD_IN; m.lead_line0 = 2*ST_LINES_PF_PAL-2*ST_LEAD_IN; m.tail = ST_TAIL;
    m.frame_len = 2*ST_LINES_PF_PAL; m.n_lines = pos+ST_TAIL;
    std::vector<u8> patch((size_t)n_frames+1, 0);
    for(int f=0;f<n_frames;f++) for(int j=0;j<H;j++) if(rec_cwd_patchable(recs+(size_t)f*H+j)) { patch[f] = 1; break; }
    std::vector<int> chains; std::vector<u8> dirty;
    cwd_plan_chains(patch.data(), fa.data(), n_frames, false, &chains, &dirty);
    int status = 0;
    std::vector<CwdLine> out_a((size_t)n_frames*112), used_a((size_t)n_frames*112), out_b((size_t)n_frames*112), used_b((size_t)n_frames*112);
    std::vector<u16> n_a(2*(size_t)n_frames, 0), n_b(2*(size_t)n_frames, 0); std::vector<u8> mode((size_t)n_frames, 0);
    CwdParams cp; memset(&cp, 0, sizeof(cp));
    cp.map = m; cp.steps = steps.data(); cp.cfg = cfg; cp.n_blocks = m.n_lines-ST_TAIL; cp.carry_out_step = -1; cp.status = &status; cp.step_mode = mode.data();
    static CwdShared sh;
    cp.chains = chains.data(); cp.step_out = out_a.data(); cp.step_used = used_a.data(); cp.step_n = n_a.data();
    for(size_t ch=0;ch<chains.size()/2;ch++) cwd_chain_cta(c, cp, (int)ch, &sh);           // exact
    std::vector<int> singles;
    for(int f=0;f<n_frames;f++) if(dirty[f]) { singles.push_back(f); singles.push_back(1); }
    cp.chains = singles.data(); cp.step_out = out_b.data(); cp.step_used = used_b.data(); cp.step_n = n_b.data();
    for(size_t ch=0;ch<singles.size()/2;ch++) cwd_chain_cta(c, cp, (int)ch, &sh);          // every frame from the raw lines
    for(int f=0;f<n_frames;f++)
    {
        if(!dirty[f]) { out[f] = 2; continue; }
        out[f] = ((n_a[2*f+1]==n_b[2*f+1])&&(memcmp(&out_a[(size_t)f*112], &out_b[(size_t)f*112], (size_t)n_a[2*f+1]*sizeof(CwdLine))==0)) ? 1 : 0;
        if(!out[f]&&(f<2)&&getenv("SDV_PROBE_VERBOSE"))
        {
            fprintf(stderr, "frame %d: kept %d / %d\n", f, n_a[2*f+1], n_b[2*f+1]);
            for(int i=0;i<n_a[2*f+1];i++)
            {
                const CwdLine &a = out_a[(size_t)f*112+i], &b = out_b[(size_t)f*112+i];
                if(memcmp(&a, &b, sizeof(CwdLine))) fprintf(stderr, "  line %d: frame %d/%d line %d/%d flags %02x/%02x crc %03x/%03x valid %03x/%03x w0 %04x/%04x w8 %04x/%04x pad %d/%d\n", i, a.frame, b.frame, a.line, b.line, a.flags, b.flags, a.crc_mask, b.crc_mask, a.valid_mask, b.valid_mask, a.w[0], b.w[0], a.w[8], b.w[8], a.pad, b.pad);
            }
        }
    }
    return 0;
}
