// pcm1_deint.cuh -- PCM-1 deinterleave (device only).
//
// PCM1Deinterleaver::processBlock / setWordData (pcm1deinterleaver.cpp:69-278): a field is 735 sub-lines (one L/R word
// pair + the CRC state of its video line); 8 interleave blocks of 92 sub-lines (last: 91), each made of two 46-pair
// stripes that swap places with the block's parity; no error correction.  Samples: 13 -> 16 bit range/sign expansion
// (pcm1datablock.cpp:309-345).  Flags as PCM1DataStitcher::outputDataBlock gives them to PCMSamplePair
// (pcm1datastitcher.cpp:1284-1301): block valid (every word of the interleave block), word valid.
// One thread per output word, one thread block per (field, interleave block): a pure permutation, HBM bound.
#pragma once
#include "sdv_common.cuh"

namespace sdv {

enum { P1_SUBLINES = 735, P1_BLOCKS = 8, P1_STRIPE = 46, P1_WORDS = 184, P1_WORDS_FIELD = 1470, P1_THREADS = 192 };

__device__ __forceinline__ i16 pcm1_sample(u32 w)
{
    if((w&0x1000u)==0) return (i16)(u16)(w<<4);
    u32 v = (w&0x0FFFu)<<2;
    if(w&0x0800u) v |= 0xC000u;
    return (i16)(u16)v;
}

__global__ void __launch_bounds__(P1_THREADS) pcm1_deint_kernel(const sdv_pcm1_subline *sub, int n_fields, int ignore_crc,
                                                                i16 *samples, u8 *sflags)
{
    const int fld = blockIdx.x/P1_BLOCKS, n = blockIdx.x%P1_BLOCKS, w = threadIdx.x;
    const int count = (n!=(P1_BLOCKS-1)) ? P1_WORDS : (P1_WORDS-2);
    const bool in_block = w<count;
    bool ok = true; u32 word = 0;
    if(in_block)
    {
        const bool even_stripe = (w&2)!=0, even_block = (n&1)==0;
        const int j = w>>2;
        const int s = n*2*P1_STRIPE+((even_block==even_stripe) ? 0 : P1_STRIPE)+j;
        const sdv_pcm1_subline r = sub[(size_t)fld*P1_SUBLINES+s];
        word = (w&1) ? r.right : r.left;
        ok = ignore_crc ? ((r.flags&SDV_P1F_BW_SET)!=0) : ((r.flags&SDV_P1F_CRC_OK)!=0);
    }
    const int block_ok = __syncthreads_and(ok ? 1 : 0);
    if(in_block)
    {
        const size_t o = (size_t)fld*P1_WORDS_FIELD+(size_t)n*P1_WORDS+w;
        samples[o] = pcm1_sample(word);
        if(sflags) sflags[o] = (u8)((block_ok ? SDV_SF_BLOCK_OK : 0)|(ok ? SDV_SF_WORD_VALID : 0));
    }
}

}   // namespace sdv
