#include "../../oracle/mp3stego_oracle.c"
#include <stdio.h>
/* CPU model for a NEXT optimisation of k_enc_probe: certified bounds that decide a probe of the step search without
 * running it.  TEST INFRASTRUCTURE (tests/test_rate_variants_model.py builds and runs it); it includes the oracle.
 *
 *   lower bound  bits >= nnz + 2 * (non-zero pairs) + (pairs holding a value > 1)
 *                (a sign per non-zero value; a non-zero pair costs >= 3 code bits in every reachable book and >= 2 as half of a
 *                 count1 quad; a pair with a value > 1 must be a big-values pair)
 *   upper bound  (largest quantised value < 15, big_values != 0 or count1 == 0)
 *                bits <= (pairs holding a value > 1) * (longest code of books 13 / 15 for values <= max, + 2 signs)
 *                        + (other non-zero pairs) * (4 + 2) + 3 * big_values + 4 * count1
 *                (a zero pair costs <= 3 bits in book 15; table B codes every count1 quad in 4 bits; with the largest value below 15 the
 *                 table search only reaches books 13 and 15, and so does the stego swap: pair targets of 13 / 15 are 15 / 13)
 * Every quantity is a per-step count that one pass over the granule can tabulate for all 121 steps.  The model walks the oracle's own
 * binary search, evaluates the bounds next to the true bit count of every probe and reports how many probes they decide -- and
 * "violations": probes where a bound would have decided differently from the true count (must be 0).
 *
 * With a 4th argument the oracle hides a random payload and every granule starts at a random payload offset (so every swap
 * variant -- no swap, swap by a '0', swap by a '1', payload exhausted mid-granule -- meets every probe): the bounds must hold
 * for the swapped tables too (pair targets stay inside {13, 15} / the linbits books, MP3_Encoder.py:419-449).
 *
 * usage: probe_bounds_model pcm.raw n_frames bitrate_kbps [payload_seed] */
static int maxlen13_le[16]={3,4,7,9,10,10,11,12,12,13,14,14,15,16,17,19};   /* books 13 and 15, values <= index */
int main(int argc,char**argv){
    int nfr=atoi(argv[2]); int br=atoi(argv[3]);
    FILE*f=fopen(argv[1],"rb"); int16_t*pcm=malloc((size_t)nfr*1152*2*2); if(fread(pcm,2,(size_t)nfr*1152*2,f)){} fclose(f);
    enc_t*e=calloc(1,sizeof*e);
    e->nch=2;e->samplerate=44100;e->bitrate=br;e->sr_index=0;
    e->buffer=pcm;e->buffer_len=(int64_t)nfr*1152*2;e->buffer_pos[0]=0;e->buffer_pos[1]=1;
    e->hide_str="";e->hide_len=0;e->cache=0;e->cache_bits=32;
    static char payload[4096];
    unsigned rs = argc > 4 ? (unsigned)atoi(argv[4]) * 2654435761u + 12345u : 0u;
    if (argc > 4) {
        for (int i = 0; i < 4095; i++) { rs = rs * 1664525u + 1013904223u; payload[i] = (rs >> 16) & 1 ? '1' : '0'; }
        e->hide_str = payload; e->hide_len = 4095;
    }
    for(int i=0;i<10000;i++) e->int2idx[i]=(int32_t)(sqrt(sqrt((double)i)*(double)i)-0.0946+0.5);
    double avg=(2.0*576/44100.0)*(1000*(double)br/8.0);
    e->whole_slots_per_frame=(int)avg;e->frac_slots_per_frame=avg-(double)e->whole_slots_per_frame;e->slot_lag=-e->frac_slots_per_frame;
    e->side_info_len=288;
    long probes=0, dec_lb=0, dec_ub=0, grans=0, viol=0; long hist[10]={0};
    for(int fr=0;fr<nfr;fr++){
        if(e->frac_slots_per_frame!=0){e->padding=e->slot_lag<=(e->frac_slots_per_frame-1.0)?1:0;e->slot_lag+=e->padding-e->frac_slots_per_frame;}
        e->bits_per_frame=8*(e->whole_slots_per_frame+e->padding);
        e->mean_bits=(int)((e->bits_per_frame-e->side_info_len)/2.0);
        mdct_sub(e);
        for(int ch=0;ch<2;ch++)for(int gr=0;gr<2;gr++){
            static int32_t ix[576]; e->xr=e->mdct_freq[ch][gr]; e->xrmax=0;
            for(int i=575;i>=0;i--){int64_t ab=e->xr[i]<0?-(int64_t)e->xr[i]:(int64_t)e->xr[i];e->xrabs[i]=(int32_t)ab;if(e->xrabs[i]>e->xrmax)e->xrmax=e->xrabs[i];}
            grinfo_t ci; memset(&ci,0,sizeof ci);
            int max_bits=e->mean_bits/2; if(max_bits>4095)max_bits=4095;
            if(!e->xrmax) continue;
            grans++; int decided=0;
            if (argc > 4) { rs = rs * 1664525u + 1013904223u; e->hide_off = (rs >> 8) % 4100; }   /* incl. offsets at / past the end */
            int next=-120,count=120; int steps[16],ns=0;
            do{int half=count/2,bit; int s=next+half; int mq=quantize(e,ix,s); steps[ns++]=s;
               if(mq>8192) bit=100000; else { bit=probe_bits(e,ix,&ci); probes++;
                  int nnz=0,nzp=0,n2p=0,last2=-1,lastnz=-1;
                  for(int p=0;p<288;p++){int a=ix[2*p],b=ix[2*p+1]; nnz+=(a!=0)+(b!=0); if(a|b){nzp++;lastnz=p;} if(a>1||b>1){n2p++;last2=p;}}
                  int lb=nnz+2*nzp+n2p;
                  int isdec=0;
                  if(lb>max_bits){ if(!(bit>max_bits)) viol++; dec_lb++; isdec=1; }   /* strict: decides both `<` and `<=` */
                  else if(mq<15){
                      /* UB: pairs with a value > 1: maxlen13(mq)+2 ; pairs with max 1: 4+2 ; zero pairs inside big region <= 3 each ; count1: <= 4*count1 + nnz_c1 (table B) */
                      (void)last2; (void)lastnz;
                      int ub=n2p*(maxlen13_le[mq]+2)+(nzp-n2p)*6+3*ci.big_values+4*ci.count1;   /* the probe's own run lengths (cheap: calc_run_len) */
                      if(ci.big_values==0&&ci.count1>0) ub=1<<30;   /* region sums pooled over stale addresses (A.E6): no bound */
                      if(ub<max_bits){ if(!(bit<max_bits)) viol++; dec_ub++; isdec=1; }
                  }
                  decided+=isdec;
               }
               if(bit<max_bits) count=half; else {next+=half;count-=half;}
            }while(count>1);
            hist[decided<9?decided:9]++;
        }
    }
    printf("granules %ld probes(bin search) %ld  decided by LB %ld  by UB %ld  violations %ld  -> %.2f of %.2f probes per granule need evaluation\n",grans,probes,dec_lb,dec_ub,viol,(double)(probes-dec_lb-dec_ub)/grans,(double)probes/grans);
    for(int i=0;i<10;i++) if(hist[i]) printf(" decided=%d:%ld",i,hist[i]); printf("\n");
    return 0;
}
