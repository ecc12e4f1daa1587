/* oracle/sdv_oracle_pcm1.c -- TEST INFRASTRUCTURE ONLY (see sdv_oracle.h).
 *
 * Sequential C restatement of the PCM-1 deinterleaver: PCM1Deinterleaver::processBlock / setWordData
 * (pcm1deinterleaver.cpp:69-278), PCM1DataBlock word storage, validity and 13 -> 16 bit sample expansion
 * (pcm1datablock.cpp:69-78, 166-199, 309-345) and the flags PCM1DataStitcher::outputDataBlock hands to PCMSamplePair
 * (pcm1datastitcher.cpp:1284-1301).
 */
#include <string.h>
#include <stdbool.h>
#include "sdv_oracle.h"

enum { P1_SUBLINES = 735, P1_BLOCKS = 8, P1_STRIPE = 46, P1_STRIPE_SHORT = 45, P1_WORDS = 184 };

static int16_t pcm1_sample(uint16_t w)
{
    if((w&0x1000)==0) return (int16_t)(uint16_t)(w<<4);
    bool positive = (w&0x0800)==0;
    w = (uint16_t)(w&~0x1000);
    w = (uint16_t)(w<<2);
    if(!positive) w |= 0xC000;
    return (int16_t)w;
}

int sdvo_deint_pcm1(const uint16_t *lr, const uint8_t *flags, int n_fields, int ignore_crc, int16_t *out_samples, uint8_t *out_flags)
{
    int o = 0;
    for(int f=0;f<n_fields;f++)
    {
        const uint16_t *flr = lr+(size_t)f*P1_SUBLINES*2;
        const uint8_t *ffl = flags+(size_t)f*P1_SUBLINES;
        for(int n=0;n<P1_BLOCKS;n++)
        {
            uint16_t words[P1_WORDS]; bool ok[P1_WORDS];
            int count = (n!=(P1_BLOCKS-1)) ? P1_WORDS : (P1_WORDS-2);
            memset(words, 0, sizeof(words)); memset(ok, 0, sizeof(ok));
            bool even_block = (n%2)==0;
            int base = n*2*P1_STRIPE;
            for(int even_stripe=1;even_stripe>=0;even_stripe--)
            {
                int len = P1_STRIPE;
                if((n==(P1_BLOCKS-1))&&even_stripe) len = P1_STRIPE_SHORT;
                int word_ofs = even_stripe ? 2 : 0;
                int sub0 = ((even_block==(even_stripe!=0)) ? base : (base+P1_STRIPE));
                for(int j=0;j<len;j++)
                {
                    int s = sub0+j;
                    bool v = ignore_crc ? ((ffl[s]&2)!=0) : ((ffl[s]&1)!=0);
                    if(word_ofs<count) { words[word_ofs] = flr[2*s]; ok[word_ofs] = v; }
                    if((word_ofs+1)<count) { words[word_ofs+1] = flr[2*s+1]; ok[word_ofs+1] = v; }
                    word_ofs += 4;
                }
            }
            bool block_ok = true;
            for(int w=0;w<count;w++) if(!ok[w]) block_ok = false;
            for(int w=0;w<count;w++)
            {
                out_samples[o] = pcm1_sample(words[w]);
                out_flags[o] = (uint8_t)((block_ok ? 1 : 0)|(ok[w] ? 2 : 0));
                o++;
            }
        }
    }
    return o;
}
