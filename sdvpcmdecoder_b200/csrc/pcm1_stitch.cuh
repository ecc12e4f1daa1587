// pcm1_stitch.cuh -- PCM1DataStitcher frame assembly: decoded PCM-1 lines of one frame -> two fields of 735 sub-lines,
// the input of the PCM-1 deinterleaver (pcm1_deint.cuh).
//
// Follows PCM1DataStitcher with automatic line offset (the default): findFrameTrim (pcm1datastitcher.cpp:202-568) picks
// the first/last data line of each field (lines with black/white levels, or -- once more than 4/5 of the field has a
// valid CRC -- lines with a valid CRC), notes a header line ahead of the data (header_present) or behind it
// (emphasis_set); splitFrameToFields (609-806) cuts the lines in that range into three sub-lines each, skipping service
// lines; findFramePadding (809-923) pads the field to 245 lines at the top, or at the bottom when the header was seen;
// fillFirst/SecondFieldForOutput (1076-1218) emit the fields in the preset order.  One thread block per frame.
#pragma once
#include "sdv_common.cuh"

namespace sdv {

enum { P1S_LINES_PF = 245, P1S_SUBLINES_PF = 735, P1S_MIN_GOOD = P1S_LINES_PF*4/5 };        // pcm1datastitcher.h:118-127

struct P1AsmScratch
{
    int good[2], first_valid[2], last_valid[2], hdr_first[2], hdr_last[2], top[2], bottom[2], count[2];
    u16 data_k[2][P1S_LINES_PF];
};

SDV_HD void p1s_atomic_add(int *p, int v)
{
#if defined(__CUDA_ARCH__)
    atomicAdd(p, v);
#else
    *p += v;
#endif
}
SDV_HD void p1s_atomic_min(int *p, int v)
{
#if defined(__CUDA_ARCH__)
    atomicMin(p, v);
#else
    if(v<*p) *p = v;
#endif
}
SDV_HD void p1s_atomic_max(int *p, int v)
{
#if defined(__CUDA_ARCH__)
    atomicMax(p, v);
#else
    if(v>*p) *p = v;
#endif
}

// fr: the H line records of the frame in stream order (odd field, then even field); out: [2][735] in output field order.
// file_start: the frame carries the NEW_FILE line -- doFrameReassemble resets header_present / emphasis_set after the
// trim search of that frame (resetState, pcm1datastitcher.cpp:63-69,1671-1675), so it is always padded at the top.
// manual: preset line offsets instead of the automatic alignment (setAutoLineOffset(false), setOdd/EvenLineOffset:
// findFrameTrim 362-383, findFramePadding 840-888): a positive offset skips lines at the top of the field, a negative one
// pads it.
SDV_HD void p1_assemble_frame_cta(const Cta &c, const sdv_line_rec *fr, int H, bool bff, bool file_start, bool manual, int ofs_odd,
                                  int ofs_even, sdv_pcm1_subline *out, P1AsmScratch *s, sdv_pcm1_frame_info *info)
{
    const int hf = H/2;
    const int BIG = 1<<30;
    c.sync();
    if(c.tid==0)
        for(int f=0;f<2;f++)
        {
            s->good[f] = 0; s->first_valid[f] = BIG; s->last_valid[f] = -1; s->hdr_first[f] = BIG; s->hdr_last[f] = -1;
            s->top[f] = BIG; s->bottom[f] = -1; s->count[f] = 0;
        }
    c.sync();
    for(int i=c.tid;i<2*hf;i+=c.n)
    {
        const int f = i/hf, k = i-f*hf;
        const u16 fl = fr[i].flags; const u8 sv = fr[i].service_type;
        if((sv==SDV_SRV_NO)&&(fl&SDV_LF_CRC_OK)) { p1s_atomic_add(&s->good[f], 1); p1s_atomic_min(&s->first_valid[f], k); p1s_atomic_max(&s->last_valid[f], k); }
        if(sv==SDV_SRV_HEADER_LINE) { p1s_atomic_min(&s->hdr_first[f], k); p1s_atomic_max(&s->hdr_last[f], k); }
    }
    c.sync();
    for(int i=c.tid;i<2*hf;i+=c.n)
    {
        const int f = i/hf, k = i-f*hf;
        if(fr[i].service_type!=SDV_SRV_NO) continue;
        const bool skip_bad = s->good[f]>P1S_MIN_GOOD;
        const u16 fl = fr[i].flags;
        if(skip_bad ? ((fl&SDV_LF_CRC_OK_IGN)!=0) : ((fl&SDV_LF_BW_SET)!=0)) { p1s_atomic_min(&s->top[f], k); p1s_atomic_max(&s->bottom[f], k); }
    }
    c.sync();
    for(int f=c.tid;f<2;f+=c.n)
    {   // the data lines of the field: non-service lines of [top, bottom], at most 245
        int n = 0;
        if(manual)
        {   // the top is preset; the bottom is the last line with data (0 = line number zero when there is none)
            const int ofs = f ? ofs_even : ofs_odd;
            s->top[f] = (ofs>0) ? ofs : 0;
        }
        if(s->bottom[f]>=0)
            for(int k=s->top[f];(k<=s->bottom[f])&&(n<P1S_LINES_PF);k++)
                if(fr[f*hf+k].service_type==SDV_SRV_NO) s->data_k[f][n++] = (u16)k;
        if(manual)
        {
            const int ofs = f ? ofs_even : ofs_odd;
            const int top_pad = (ofs>0) ? 0 : -ofs;
            const int span = (s->bottom[f]>=s->top[f]) ? (s->bottom[f]-s->top[f]+1) : 0;
            if(span+top_pad>P1S_LINES_PF)
            {   // too many lines: the bottom is trimmed, the count becomes the span of line numbers (findFramePadding 858-866)
                int m = P1S_LINES_PF-top_pad;
                if(m<0) m = 0;
                if(m<n) n = m;
            }
        }
        s->count[f] = n;
    }
    c.sync();
    const bool header_present = (!file_start)&&((s->hdr_first[0]<s->first_valid[0])||(s->hdr_first[1]<s->first_valid[1]));
    const bool emphasis_set = (!file_start)&&((s->hdr_last[0]>s->last_valid[0])||(s->hdr_last[1]>s->last_valid[1]));
    for(int i=c.tid;i<2*P1S_SUBLINES_PF;i+=c.n)
    {
        const int slot = i/P1S_SUBLINES_PF, sl = i-slot*P1S_SUBLINES_PF;
        const int f = bff ? (1-slot) : slot;                    // which captured field goes out in this slot
        const int line = sl/3, part = sl-3*line;
        const int n = s->count[f];
        const int pad = P1S_LINES_PF-n;
        const int ofs = f ? ofs_even : ofs_odd;
        const int top_pad = manual ? ((ofs>0) ? 0 : ((-ofs<P1S_LINES_PF) ? -ofs : P1S_LINES_PF)) : (header_present ? 0 : pad);
        sdv_pcm1_subline o;
        o.left = o.right = 0x1000; o.flags = 0; o.reserved[0] = o.reserved[1] = o.reserved[2] = 0;
        const int j = line-top_pad;
        if((j>=0)&&(j<n))
        {
            const sdv_line_rec *r = fr+(size_t)f*hf+s->data_k[f][j];
            o.left = r->words[2*part]; o.right = r->words[2*part+1];
            u8 fl = 0;
            if(r->flags&SDV_LF_CRC_OK) fl |= SDV_P1F_CRC_OK;
            if(r->flags&SDV_LF_BW_SET) fl |= SDV_P1F_BW_SET;
            if((part==0)&&(r->mark_stages&0x0F)) fl |= SDV_P1F_PICKED_LEFT;
            if(r->mark_stages&0xF0) fl |= SDV_P1F_PICKED_RIGHT;
            o.flags = fl;
        }
        out[i] = o;
    }
    if(info&&(c.tid==0))
    {
        sdv_pcm1_frame_info t;
        t.odd_top = (u16)((s->bottom[0]>=0) ? s->top[0] : 0); t.odd_bottom = (u16)((s->bottom[0]>=0) ? s->bottom[0] : 0);
        t.even_top = (u16)((s->bottom[1]>=0) ? s->top[1] : 0); t.even_bottom = (u16)((s->bottom[1]>=0) ? s->bottom[1] : 0);
        t.odd_data_lines = (u16)s->count[0]; t.even_data_lines = (u16)s->count[1];
        t.header_present = header_present ? 1 : 0; t.emphasis_set = emphasis_set ? 1 : 0;
        t.reserved[0] = t.reserved[1] = 0;
        *info = t;
    }
    c.sync();
}

#if defined(__CUDACC__)
__global__ void __launch_bounds__(256) pcm1_assemble_kernel(const sdv_line_rec *recs, int n_frames, int H, int bff, int file_start, int manual, int ofs_odd,
                                                            int ofs_even, sdv_pcm1_subline *sub,
                                                            sdv_pcm1_frame_info *info)
{
    __shared__ P1AsmScratch s;
    const int f = blockIdx.x;
    if(f>=n_frames) return;
    Cta c = { (int)threadIdx.x, (int)blockDim.x };
    p1_assemble_frame_cta(c, recs+(size_t)f*H, H, bff!=0, (file_start!=0)&&(f==0), manual!=0, ofs_odd, ofs_even, sub+(size_t)f*2*P1S_SUBLINES_PF, &s, info ? info+f : (sdv_pcm1_frame_info *)0);
}
#endif

}   // namespace sdv
